// genhancer_b200 -- Philox-4x32-10 (counter-based RNG): the LoRA-dropout mask is a pure function of (seed, offset, element
// index), shared by the plain dropout kernels (elementwise.cu) and the fused LoRA kernels (lora_fused.cu).
#pragma once
#include "common.cuh"

namespace gh {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// the 128 random bits of the 8-element block `blk` (16 bits per element, element q in the low / high half of word q / 2)
__device__ __forceinline__ uint4 dropout_bits(int64_t blk, uint2 off, uint2 key) {
  return philox4x32_10(make_uint4(static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), off.x, off.y), key);
}

}  // namespace gh
