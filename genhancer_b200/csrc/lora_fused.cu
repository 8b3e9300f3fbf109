// genhancer_b200 -- the LoRA branch with input dropout, fused around the mask.
//
// peft's lora.Linear computes lora_B(lora_A(dropout(x))) (train_SigLIP_stage2_all.py:134-142, lora_dropout 0.1): the mask
// sits between x and A, so neither u = drop(x) A^T nor dx += drop'(du A) can ride in the base GEMM.  As separate passes
// (dropout kernel -> skinny tcgen05 GEMM, skinny GEMM -> dropout-add kernel) they were 20 of the 216 ms of the SigLIP
// stage-2 step, every one of them a full trip of an [M, K] activation through HBM.  Here the mask is applied IN REGISTERS
// between the load and a warp-level mma.sync (m16n8k16, bf16 -> fp32: the operands of tcgen05.mma come from shared
// memory / TMEM, not from registers, and the rank-16 products are a rounding error next to the HBM traffic):
//
//   gh_lora_dropout_fwd   xd = drop(x)  (kept for the A-gradient GEMM)  AND  u = s xd A^T [, 1, 0...]   in one pass over x
//   gh_lora_dropout_bwd   dx += drop'(du A)                              one read-modify-write of dx, du A never exists
//
// The mask is the one of gh_dropout_fwd / gh_dropout_bwd_add (philox.cuh: one Philox block per 8 consecutive elements of
// the contiguous [M, K] tensor).  Fragment trick: a thread loads / stores 16 bytes = 8 consecutive K (or output) columns
// = one Philox block; the reduction index of an MMA is a dummy, so lane t of a quad simply DECLARES its elements 0-3 to
// be the k-indices {2t, 2t+1, 2t+8, 2t+9} of one MMA and 4-7 those of the next, and the other operand is loaded with the
// same convention -- no shuffles, no shared-memory transposes for the reduction side.
#include "common.cuh"
#include "internal.h"
#include "philox.cuh"

namespace gh {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// keep-mask of one 8-element block applied to 8 bf16 (as 4 words), scaled by 1 / (1 - p)
__device__ __forceinline__ uint4 drop8(uint4 v, uint4 r, uint32_t thresh16, float inv_keep) {
  const uint32_t rw[4] = {r.x, r.y, r.z, r.w}, vw[4] = {v.x, v.y, v.z, v.w};
  uint32_t ow[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 a = unpack_bf16x2(vw[q]);
    ow[q] = pack_bf16x2((rw[q] & 0xffffu) >= thresh16 ? a.x * inv_keep : 0.f, (rw[q] >> 16) >= thresh16 ? a.y * inv_keep : 0.f);
  }
  return make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

struct LoraDropParams {
  int M, K, R, RX;        // tokens, features, LoRA rank of the (fused) group, extra columns of u (0 or 16: [1, 0 ...])
  int64_t ldu;            // row pitch of u / du in elements
  float alpha;            // u scale (LoRA scaling)
  uint32_t thresh16;
  float inv_keep;
  uint2 key;
  unsigned long long offset;
  const unsigned long long* offset_base;
  const __nv_bfloat16* pre;   // backward only, may be NULL: dx = (dx + drop'(du A)) * act'(pre)  -- the dgrad through an MLP's
  int act;                    // activation (fc2's input gradient), which the masked term keeps out of the GEMM's epilogue
};

// ----------------------------------------------------------------------------------------------------------
// forward: CTA = 8 warps = 2 tiles of 16 token rows x 4 quarters of K (a warp per tile and quarter: 16 rows x the whole K
// per warp left ~10 warps per SM with two 16-byte loads each in flight -- 1.6 TB/s); the quarters' partial sums meet in
// shared memory.  The next iteration's loads are issued before the current one's Philox rounds and MMAs.
// NT = R / 8 output tiles of 8 LoRA columns.
// ----------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) lora_dropout_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ xd,
                                                               const bf16* __restrict__ A, bf16* __restrict__ u,
                                                               const LoraDropParams p) {
  __shared__ float red[6][32][NT * 4];   // partial sums of the K quarters 1..3 of both row tiles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int tile = warp & 1, quarter = warp >> 1;
  const int row0 = (blockIdx.x * 2 + tile) * 16;
  const unsigned long long o64 = p.offset + (p.offset_base != nullptr ? *p.offset_base : 0ull);
  const uint2 off = make_uint2(static_cast<uint32_t>(o64), static_cast<uint32_t>(o64 >> 32));
  const int ra = row0 + g, rb = row0 + g + 8;
  const bool oka = ra < p.M, okb = rb < p.M;
  const int64_t K = p.K;
  constexpr int U = 2;                   // 32-column chunks per iteration
  const int iters = (p.K + 32 * U - 1) / (32 * U);
  const int per = (iters + 3) / 4;
  const int it0 = quarter * per, it1 = min(iters, it0 + per);
  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  // The LoRA rows ride in the same software pipeline as x when they fit in registers (rank 16: NT = 2): read at their first
  // use they exposed one L2 round trip per chunk and output tile in front of every MMA pair.
  constexpr bool PREFETCH_W = NT <= 2;
  constexpr int NW = PREFETCH_W ? NT : 1;
  uint4 va[U], vb[U], na[U], nb[U], wa[U][NW], wn[U][NW];
  auto load = [&](int it, uint4 (&a)[U], uint4 (&b)[U], uint4 (&w)[U][NW]) {
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      const int c8 = (it * U + uu) * 32 + t * 8;
      const bool okc = it < it1 && c8 < p.K;
      a[uu] = (oka && okc) ? *reinterpret_cast<const uint4*>(x + ra * K + c8) : make_uint4(0u, 0u, 0u, 0u);
      b[uu] = (okb && okc) ? *reinterpret_cast<const uint4*>(x + rb * K + c8) : make_uint4(0u, 0u, 0u, 0u);
      if (PREFETCH_W) {
#pragma unroll
        for (int n = 0; n < NW; ++n)
          w[uu][n] = okc ? __ldg(reinterpret_cast<const uint4*>(A + static_cast<int64_t>(n * 8 + g) * K + c8)) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  load(it0, va, vb, wa);
  for (int it = it0; it < it1; ++it) {
    load(it + 1, na, nb, wn);
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      const int c8 = (it * U + uu) * 32 + t * 8;
      const bool okc = c8 < p.K;
      if (oka && okc) {
        va[uu] = drop8(va[uu], dropout_bits((ra * K + c8) >> 3, off, p.key), p.thresh16, p.inv_keep);
        *reinterpret_cast<uint4*>(xd + ra * K + c8) = va[uu];
      }
      if (okb && okc) {
        vb[uu] = drop8(vb[uu], dropout_bits((rb * K + c8) >> 3, off, p.key), p.thresh16, p.inv_keep);
        *reinterpret_cast<uint4*>(xd + rb * K + c8) = vb[uu];
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        // LoRA row n * 8 + g, the same 8 K columns: elements 0-3 feed the first MMA, 4-7 the second
        uint4 w;
        if (PREFETCH_W) w = wa[uu][n < NW ? n : 0];
        else w = okc ? __ldg(reinterpret_cast<const uint4*>(A + static_cast<int64_t>(n * 8 + g) * K + c8)) : make_uint4(0u, 0u, 0u, 0u);
        mma16816(acc[n], va[uu].x, vb[uu].x, va[uu].y, vb[uu].y, w.x, w.y);
        mma16816(acc[n], va[uu].z, vb[uu].z, va[uu].w, vb[uu].w, w.z, w.w);
      }
    }
#pragma unroll
    for (int uu = 0; uu < U; ++uu) {
      va[uu] = na[uu]; vb[uu] = nb[uu];
#pragma unroll
      for (int n = 0; n < NW; ++n) wa[uu][n] = wn[uu][n];
    }
  }
  if (quarter > 0) {
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) red[(quarter - 1) * 2 + tile][lane][n * 4 + i] = acc[n][i];
  }
  __syncthreads();
  if (quarter > 0 || row0 >= p.M) return;
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[n][i] += red[q * 2 + tile][lane][n * 4 + i];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    if (oka) *reinterpret_cast<uint32_t*>(u + ra * p.ldu + n * 8 + 2 * t) = pack_bf16x2(acc[n][0] * p.alpha, acc[n][1] * p.alpha);
    if (okb) *reinterpret_cast<uint32_t*>(u + rb * p.ldu + n * 8 + 2 * t) = pack_bf16x2(acc[n][2] * p.alpha, acc[n][3] * p.alpha);
  }
  if (p.RX && t < 2) {   // [1, 0, ..., 0]: the bias gradient rides in the dB GEMM as this column (tower_engine.LinGroup)
    const int r = t == 0 ? ra : rb;
    if (r < p.M) {
      *reinterpret_cast<uint4*>(u + r * p.ldu + p.R) = make_uint4(0x00003f80u, 0u, 0u, 0u);   // bf16 1.0, then zeros
      *reinterpret_cast<uint4*>(u + r * p.ldu + p.R + 8) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// backward: dx[M, K] += drop'(du[M, R] A[R, K]).  CTA = 8 warps x 16 rows, one 128-column strip of K; the strip of A sits
// transposed in shared memory ([column][rank], so a B fragment is one 4-byte load).  KS = R / 16 k-steps.
// Output columns: lane t of a quad owns the 8 consecutive columns t * 8 .. t * 8 + 7 of a 32-column step (one Philox
// block, one 16-byte read-modify-write): column 2 t' + i of MMA n-tile j is DECLARED to be column t' * 8 + j * 2 + i.
// ----------------------------------------------------------------------------------------------------------
template <int KS, bool ACT>   // ACT: also multiply by act'(pre) (its own instantiation: the activation code costs ~40 registers)
__global__ void __launch_bounds__(256) lora_dropout_bwd_kernel(const bf16* __restrict__ du, const bf16* __restrict__ A,
                                                               bf16* __restrict__ dx, const LoraDropParams p) {
  constexpr int R = KS * 16;
  constexpr int P = R + 2;     // smem row pitch in bf16 (an odd number of words: the 8 rows a warp reads spread over banks)
  constexpr int KC = 128;
  __shared__ bf16 sAT[KC * P];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int c0 = blockIdx.y * KC;
  const int64_t K = p.K;
  for (int i = threadIdx.x; i < R * KC; i += 256) {   // A[r][c0 + c] -> sAT[c][r]; coalesced along c
    const int r = i / KC, c = i - r * KC;
    sAT[c * P + r] = (c0 + c < p.K) ? A[static_cast<int64_t>(r) * K + c0 + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  const int row0 = (blockIdx.x * 8 + warp) * 16;
  if (row0 >= p.M) return;
  const unsigned long long o64 = p.offset + (p.offset_base != nullptr ? *p.offset_base : 0ull);
  const uint2 off = make_uint2(static_cast<uint32_t>(o64), static_cast<uint32_t>(o64 >> 32));
  const int ra = row0 + g, rb = row0 + g + 8;
  const bool oka = ra < p.M, okb = rb < p.M;
  uint32_t a[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    a[ks][0] = oka ? *reinterpret_cast<const uint32_t*>(du + ra * p.ldu + ks * 16 + 2 * t) : 0u;
    a[ks][1] = okb ? *reinterpret_cast<const uint32_t*>(du + rb * p.ldu + ks * 16 + 2 * t) : 0u;
    a[ks][2] = oka ? *reinterpret_cast<const uint32_t*>(du + ra * p.ldu + ks * 16 + 8 + 2 * t) : 0u;
    a[ks][3] = okb ? *reinterpret_cast<const uint32_t*>(du + rb * p.ldu + ks * 16 + 8 + 2 * t) : 0u;
  }
  // the read halves of all four read-modify-writes first (8 x 16 bytes in flight per lane), then MMAs / Philox / stores
  uint4 da[KC / 32], db[KC / 32];
#pragma unroll
  for (int st = 0; st < KC / 32; ++st) {
    const int c8 = c0 + st * 32 + t * 8;
    da[st] = (oka && c8 < p.K) ? *reinterpret_cast<const uint4*>(dx + ra * K + c8) : make_uint4(0u, 0u, 0u, 0u);
    db[st] = (okb && c8 < p.K) ? *reinterpret_cast<const uint4*>(dx + rb * K + c8) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int st = 0; st < KC / 32; ++st) {
    const int cb = st * 32;
    const int c8 = c0 + cb + t * 8;
    if (c0 + cb >= p.K) break;
    const bool okc = c8 < p.K;
    uint4 pa = make_uint4(0u, 0u, 0u, 0u), pb = pa;
    if (ACT) {   // requested before the MMAs and the Philox rounds
      if (oka && okc) pa = *reinterpret_cast<const uint4*>(p.pre + ra * K + c8);
      if (okb && okc) pb = *reinterpret_cast<const uint4*>(p.pre + rb * K + c8);
    }
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      const bf16* col = sAT + (cb + (g >> 1) * 8 + j * 2 + (g & 1)) * P;   // the column n = g of tile j stands for
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        mma16816(acc[j], a[ks][0], a[ks][1], a[ks][2], a[ks][3], *reinterpret_cast<const uint32_t*>(col + ks * 16 + 2 * t),
                 *reinterpret_cast<const uint32_t*>(col + ks * 16 + 8 + 2 * t));
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? rb : ra;
      if (!((half ? okb : oka) && okc)) continue;
      const uint4 rnd = dropout_bits((r * K + c8) >> 3, off, p.key);
      const uint4 d = half ? db[st] : da[st];
      const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w}, dw[4] = {d.x, d.y, d.z, d.w};
      float v[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 c = unpack_bf16x2(dw[j]);
        const float v0 = acc[j][2 * half], v1 = acc[j][2 * half + 1];
        v[2 * j] = c.x + ((rw[j] & 0xffffu) >= p.thresh16 ? v0 * p.inv_keep : 0.f);
        v[2 * j + 1] = c.y + ((rw[j] >> 16) >= p.thresh16 ? v1 * p.inv_keep : 0.f);
      }
      if (ACT) {
        // (as the separate pass did: the sum is rounded to bf16 before it meets act')
        const uint4 pr = half ? pb : pa;
        const uint32_t pw[4] = {pr.x, pr.y, pr.z, pr.w};
        float x[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 xx = unpack_bf16x2(pw[j]);
          x[2 * j] = xx.x; x[2 * j + 1] = xx.y;
          v[2 * j] = __bfloat162float(__float2bfloat16(v[2 * j]));
          v[2 * j + 1] = __bfloat162float(__float2bfloat16(v[2 * j + 1]));
        }
        act_bwd_mul8(p.act, x, v);
      }
      *reinterpret_cast<uint4*>(dx + r * K + c8) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

static int fill_params(LoraDropParams& q, int M, int K, int R, int RX, int64_t ldu, float alpha, float p, uint64_t seed,
                       uint64_t offset, const uint64_t* offset_base, const char* who) {
  GH_REQUIRE(M >= 0 && K > 0 && K % 8 == 0, GH_ERR_BAD_SHAPE, "%s: K=%d must be a positive multiple of 8", who, K);
  GH_REQUIRE(R == 16 || R == 32 || R == 48, GH_ERR_UNSUPPORTED, "%s: rank %d unsupported (16, 32, 48)", who, R);
  GH_REQUIRE(RX == 0 || RX == 16, GH_ERR_UNSUPPORTED, "%s: extra columns %d (0 or 16)", who, RX);
  GH_REQUIRE(ldu >= R + RX && ldu % 8 == 0, GH_ERR_ALIGN, "%s: ldu=%lld must be a multiple of 8, >= R + RX", who, (long long)ldu);
  GH_REQUIRE(p >= 0.f && p < 1.f, GH_ERR_BAD_SHAPE, "%s: p=%f outside [0, 1)", who, p);
  q.M = M; q.K = K; q.R = R; q.RX = RX; q.ldu = ldu; q.alpha = alpha;
  const double t = static_cast<double>(p) * 65536.0 + 0.5;
  q.thresh16 = t >= 65535.0 ? 65535u : static_cast<uint32_t>(t);
  q.inv_keep = 1.f / (1.f - p);
  q.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  q.offset = offset;
  q.offset_base = reinterpret_cast<const unsigned long long*>(offset_base);
  return GH_OK;
}

}  // namespace gh

using namespace gh;

extern "C" int gh_lora_dropout_fwd(const void* x_bf16, void* xd_bf16, const void* a_bf16, void* u_bf16, int32_t M, int32_t K,
                                   int32_t R, int32_t RX, int64_t ldu, float alpha, float p, uint64_t seed, uint64_t offset,
                                   const uint64_t* offset_base, void* stream) {
  GH_REQUIRE(x_bf16 && xd_bf16 && a_bf16 && u_bf16, GH_ERR_NULL, "gh_lora_dropout_fwd: NULL pointer");
  GH_REQUIRE(aligned16(x_bf16) && aligned16(xd_bf16) && aligned16(a_bf16) && aligned16(u_bf16), GH_ERR_ALIGN,
             "gh_lora_dropout_fwd: 16-byte alignment");
  LoraDropParams q{};
  if (int e = fill_params(q, M, K, R, RX, ldu, alpha, p, seed, offset, offset_base, "gh_lora_dropout_fwd")) return e;
  if (M == 0) return GH_OK;
  const int grid = (M + 31) / 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bf16 *x = static_cast<const bf16*>(x_bf16), *a = static_cast<const bf16*>(a_bf16);
  bf16 *xd = static_cast<bf16*>(xd_bf16), *u = static_cast<bf16*>(u_bf16);
  switch (R) {
    case 16: lora_dropout_fwd_kernel<2><<<grid, 256, 0, s>>>(x, xd, a, u, q); break;
    case 32: lora_dropout_fwd_kernel<4><<<grid, 256, 0, s>>>(x, xd, a, u, q); break;
    default: lora_dropout_fwd_kernel<6><<<grid, 256, 0, s>>>(x, xd, a, u, q); break;
  }
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_lora_dropout_bwd(const void* du_bf16, const void* a_bf16, void* dx_bf16, int32_t M, int32_t K, int32_t R,
                                   int64_t ldu, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_base,
                                   const void* act_pre_bf16, int32_t act, void* stream) {
  GH_REQUIRE(du_bf16 && a_bf16 && dx_bf16, GH_ERR_NULL, "gh_lora_dropout_bwd: NULL pointer");
  GH_REQUIRE(aligned16(du_bf16) && aligned16(a_bf16) && aligned16(dx_bf16), GH_ERR_ALIGN, "gh_lora_dropout_bwd: 16-byte alignment");
  LoraDropParams q{};
  if (int e = fill_params(q, M, K, R, 0, ldu, 1.f, p, seed, offset, offset_base, "gh_lora_dropout_bwd")) return e;
  GH_REQUIRE(!act_pre_bf16 || aligned16(act_pre_bf16), GH_ERR_ALIGN, "gh_lora_dropout_bwd: act_pre must be 16-byte aligned");
  q.pre = static_cast<const __nv_bfloat16*>(act_pre_bf16);
  q.act = act;
  if (M == 0) return GH_OK;
  const dim3 grid((M + 127) / 128, (K + 127) / 128);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bf16 *du = static_cast<const bf16*>(du_bf16), *a = static_cast<const bf16*>(a_bf16);
  bf16* dx = static_cast<bf16*>(dx_bf16);
  if (q.pre != nullptr) {
    switch (R) {
      case 16: lora_dropout_bwd_kernel<1, true><<<grid, 256, 0, s>>>(du, a, dx, q); break;
      case 32: lora_dropout_bwd_kernel<2, true><<<grid, 256, 0, s>>>(du, a, dx, q); break;
      default: lora_dropout_bwd_kernel<3, true><<<grid, 256, 0, s>>>(du, a, dx, q); break;
    }
  } else {
    switch (R) {
      case 16: lora_dropout_bwd_kernel<1, false><<<grid, 256, 0, s>>>(du, a, dx, q); break;
      case 32: lora_dropout_bwd_kernel<2, false><<<grid, 256, 0, s>>>(du, a, dx, q); break;
      default: lora_dropout_bwd_kernel<3, false><<<grid, 256, 0, s>>>(du, a, dx, q); break;
    }
  }
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}
