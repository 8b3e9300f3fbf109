// genhancer_b200 -- patch-embed convolution as an IMPLICIT GEMM on tcgen05 (north_star bullet 1).
//
//   tokens[m, n] = sum_k A[m, k] * W[n, k] (+ bias[n]),   m = (image b, patch row py, patch column px),
//                                                          k = c * p * p + i * p + j   (Conv2d weight [D, 3, p, p] flattened)
//   A[m, k] = (pixel[b, c, py * p + i, px * p + j] - mean[c]) / std[c]
//
// replaces HF CLIPVisionEmbeddings.patch_embedding / SiglipVisionEmbeddings.patch_embedding (Conv2d(3, D, kernel = stride
// = 14), modeling_clip.py:147-153,208-209) together with the two transforms the reference applies before it (ToTensor of
// the loader, transforms.Normalize of train_SigLIP_stage1.py:54-59).  The A operand is never materialised in HBM: the
// first version wrote it with a gather kernel (gh_patch_im2col) and read it back with the plain GEMM.
//
// Why the A tile is not a TMA box.  stride == kernel makes A a pure 6-D view of the image, but a patch row is 14
// pixels = 56 bytes of fp32 (14 bytes of uint8): TMA needs box rows and global strides in multiples of 16 bytes, so
// neither the box {14 px} nor the stride between patches of one image row (56 B) can be expressed, for either input
// format.  The tile is instead GATHERED: four producer warps (thread = patch row of the tile) read the pixels, apply
// ToTensor / Normalize, convert to bf16 and store straight into the K-major SWIZZLE_128B shared-memory layout the MMA
// reads (16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4)), then hand the stage over with
// fence.proxy.async + an mbarrier.  The weight tile arrives by TMA; accumulators live in TMEM; the same four warps
// run the epilogue (tcgen05.ld -> bias -> bf16 -> 16-byte stores).  Two stages: the gather of k block i + 1 overlaps
// the MMAs of k block i.
//
// Grid: (ceil(M / 128), ceil(D / BN)), BN = 256.  The kernel is tiny in the step (22 GFLOP of 86 TFLOP): the point is
// the dataflow (image bytes -> tokens with no intermediate), not tensor-pipe utilisation.
#include "common.cuh"
#include "internal.h"

namespace gh {

using bf16 = __nv_bfloat16;

struct PatchEmbedParams {
  const void* img;      // fp32 NCHW [B,3,S,S] in [0,1], or uint8 HWC [B,S,S,3]
  int u8;
  int B, S, p, G;       // image side, patch side, patches per side
  int M, N, K;          // B*G*G, D, 3*p*p
  int num_k_blocks;     // ceil(K / 64)
  float mean[3], istd[3];
  const float* bias;    // fp32 [N] or NULL
  bf16* out;            // [M, ldo]
  int64_t ldo;
};

constexpr int PE_BN = 256;
constexpr int PE_A_BYTES = 128 * 64 * 2;
constexpr int PE_B_BYTES = PE_BN * 64 * 2;
constexpr int PE_STAGE = PE_A_BYTES + PE_B_BYTES;
constexpr int PE_SMEM = 2 * PE_STAGE + 256 + 1024;

__global__ void __launch_bounds__(160, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmap_w, const PatchEmbedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * PE_STAGE);
  uint64_t* a_full = bars;        // [2] 4 producer warps arrived: the gathered A tile is in smem
  uint64_t* b_full = bars + 2;    // [2] TMA: the weight tile is in smem
  uint64_t* empty = bars + 4;     // [2] MMAs that read the stage have retired
  uint64_t* acc_full = bars + 6;  // accumulator complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * PE_BN;
  const int nkb = p.num_k_blocks;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 4);
      mbar_init(&b_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<PE_BN>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 4) {
    // ---------------- control warp: weight tiles by TMA, MMA issue ----------------
    constexpr uint32_t idesc = umma_idesc_bf16(128, PE_BN, false, false);
    const uint64_t desc = umma_desc_base(16u, 1024u);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1;
      const uint32_t ph = static_cast<uint32_t>(kb >> 1) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);                       // the MMAs of k block kb - 2 no longer read the stage
      if (elect_one()) {
        mbar_arrive_expect_tx(&b_full[s], PE_B_BYTES);
        tma_load_2d(smem + s * PE_STAGE + PE_A_BYTES, &tmap_w, &b_full[s], kb * 64, n0);   // rows >= N, k >= ld: zero fill
      }
      __syncwarp();
      mbar_wait(&a_full[s], ph);
      mbar_wait(&b_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_u32(smem + s * PE_STAGE), sB = sA + PE_A_BYTES;
        const int left = p.K - kb * 64;                    // k columns of this block that hold data
        const int steps = left >= 64 ? 4 : (left + 15) / 16;
        for (int k = 0; k < steps; ++k)
          umma_ss(tmem, umma_desc_at(desc, sA + k * 32), umma_desc_at(desc, sB + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
        if (kb == nkb - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------- producers: thread = one patch (one A row); then the epilogue of the same row ----------------
    const int r = threadIdx.x;                              // 0..127 == TMEM lane
    const int m = m0 + r;
    const bool live = m < p.M;
    const int gg = p.G * p.G;
    const int b = live ? m / gg : 0;
    const int rem = live ? m - b * gg : 0;
    const int py = rem / p.G, px = rem - py * p.G;
    const int pp = p.p * p.p;
    const int64_t S = p.S;
    const float* f32 = static_cast<const float*>(p.img);
    const uint8_t* u8 = static_cast<const uint8_t*>(p.img);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1;
      const uint32_t ph = static_cast<uint32_t>(kb >> 1) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* row = smem + s * PE_STAGE + r * 128;
      // (c, i, j) of the first k of this block, then walked incrementally: j fastest, then i, then c
      int k = kb * 64;
      int c = k / pp;
      int ij = k - c * pp;
      int i = ij / p.p, j = ij - i * p.p;
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {                      // 8 chunks of 8 elements = 16 bytes
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float x = 0.f;
          if (live && k < p.K) {
            const int y = py * p.p + i, xx = px * p.p + j;
            const float raw = p.u8 ? __fdiv_rn(static_cast<float>(u8[((b * S + y) * S + xx) * 3 + c]), 255.f)
                                   : f32[((static_cast<int64_t>(b) * 3 + c) * S + y) * S + xx];
            x = (raw - p.mean[c]) * p.istd[c];
          }
          v[e] = x;
          ++k;
          if (++j == p.p) { j = 0; if (++i == p.p) { i = 0; ++c; } }
        }
        uint4 w;
        w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]);
        w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(row + ((ch ^ (r & 7)) << 4)) = w;
      }
      fence_proxy_async_smem();                             // generic-proxy stores -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
    }
    // ---- epilogue: this thread's token row ----
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    bf16* orow = p.out + static_cast<int64_t>(m) * p.ldo + n0;
#pragma unroll 1
    for (int cc = 0; cc < PE_BN / 32; ++cc) {
      if (n0 + cc * 32 >= p.N) break;                       // (warp-uniform)
      uint32_t o[32];
      tmem_ld_32x32(t_lane + cc * 32, o);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = n0 + cc * 32 + q * 8;
          if (n >= p.N) break;
          float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
          }
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * q + 0]) + bb[0], __uint_as_float(o[8 * q + 1]) + bb[1]);
          w.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) + bb[2], __uint_as_float(o[8 * q + 3]) + bb[3]);
          w.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) + bb[4], __uint_as_float(o[8 * q + 5]) + bb[5]);
          w.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) + bb[6], __uint_as_float(o[8 * q + 7]) + bb[7]);
          *reinterpret_cast<uint4*>(orow + cc * 32 + q * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<PE_BN>(tmem);
}

int patch_embed_init() {
  GH_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PE_SMEM));
  return GH_OK;
}

}  // namespace gh

using namespace gh;

extern "C" int gh_patch_embed_fwd(const void* img, int32_t img_is_u8hwc, const void* w_bf16, int64_t ldw, const float* bias,
                                  void* out_bf16, int64_t ldo, int32_t B, int32_t S, int32_t patch, int32_t D,
                                  const float* mean3, const float* std3, void* stream) {
  GH_REQUIRE(img && w_bf16 && out_bf16, GH_ERR_NULL, "gh_patch_embed_fwd: NULL pointer");
  GH_REQUIRE(B > 0 && S > 0 && patch > 0 && S / patch > 0 && D > 0 && D % 8 == 0, GH_ERR_BAD_SHAPE,
             "gh_patch_embed_fwd: bad shape (D %% 8 == 0)");
  const int K = 3 * patch * patch;
  GH_REQUIRE(ldw >= K && ldw % 8 == 0 && ldo >= D && ldo % 8 == 0, GH_ERR_ALIGN,
             "gh_patch_embed_fwd: ldw >= 3*p*p, ldo >= D, both multiples of 8");
  GH_REQUIRE(aligned16(w_bf16) && aligned16(out_bf16) && (!bias || aligned16(bias)), GH_ERR_ALIGN,
             "gh_patch_embed_fwd: 16-byte alignment");
  PatchEmbedParams p{};
  p.img = img; p.u8 = img_is_u8hwc != 0;
  p.B = B; p.S = S; p.p = patch; p.G = S / patch;
  p.M = B * p.G * p.G; p.N = D; p.K = K;
  p.num_k_blocks = (K + 63) / 64;
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean3 ? mean3[c] : 0.f;                    // host pointers (3 floats)
    p.istd[c] = std3 ? 1.f / std3[c] : 1.f;
  }
  p.bias = bias;
  p.out = static_cast<bf16*>(out_bf16); p.ldo = ldo;
  CUtensorMap tw;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(ldw), static_cast<uint64_t>(D)};   // (pad columns K..ldw are zero)
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {64, PE_BN};
    if (int e = make_tmap_bf16(&tw, w_bf16, 2, dims, strides, box, nullptr)) return e;
  }
  dim3 grid((p.M + 127) / 128, (D + PE_BN - 1) / PE_BN);
  patch_embed_kernel<<<grid, 160, PE_SMEM, static_cast<cudaStream_t>(stream)>>>(tw, p);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}
