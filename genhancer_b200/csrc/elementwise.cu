// genhancer_b200 -- HBM-bound flow-matching kernels (128-bit loads, warp-shuffle reductions).
#include "common.cuh"
#include "internal.h"
#include "philox.cuh"

namespace gh {

// x_t = bf16((1-t) x1 + t x0)        algorithmic bytes / element: 4 + 4 + 2 = 10
__global__ void __launch_bounds__(256) fm_interp_kernel(const float4* __restrict__ x1, const float4* __restrict__ x0,
                                                        const float* __restrict__ t, uint2* __restrict__ xt,
                                                        int64_t n4, int64_t per_sample4) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float tb = __ldg(t + i / per_sample4);
    const float4 a = __ldcs(x1 + i);
    const float4 b = __ldcs(x0 + i);
    const float om = 1.f - tb;
    // same operation order as the reference expression (1 - t) * x_1 + t * x_0 (no fma contraction)
    uint2 o;
    o.x = pack_bf16x2(__fadd_rn(__fmul_rn(om, a.x), __fmul_rn(tb, b.x)),
                      __fadd_rn(__fmul_rn(om, a.y), __fmul_rn(tb, b.y)));
    o.y = pack_bf16x2(__fadd_rn(__fmul_rn(om, a.z), __fmul_rn(tb, b.z)),
                      __fadd_rn(__fmul_rn(om, a.w), __fmul_rn(tb, b.w)));
    xt[i] = o;
  }
}

// loss += sum((pred - (x0 - x1))^2) / numel ; dpred = bf16(scale * 2 (pred - (x0-x1)) / numel)
// algorithmic bytes / element: 2 + 4 + 4 + 2 = 12
// The grid is ONE thread-block cluster (8 CTAs of 1024 threads): every CTA reduces its share in a fixed order, the CTAs'
// partial sums meet in CTA 0's shared memory (DSMEM) and ONE thread adds their fixed-order sum to the accumulator -- the
// loss is bit-reproducible.  (The first version let every block atomicAdd its partial: the loss of two identical steps
// could differ in the last bit with the order the blocks happened to finish in, which a bit-equality test caught.)  The
// tensor is small (batch x 441 x 64 latents: ~13 MB of traffic), so eight SMs cost ~10 us per step.
constexpr int FM_MSE_CTAS = 8;
__global__ void __launch_bounds__(1024) fm_mse_kernel(const uint2* __restrict__ pred, const float4* __restrict__ x0,
                                                      const float4* __restrict__ x1, float* __restrict__ loss,
                                                      uint2* __restrict__ dpred, float gscale, float inv_numel,
                                                      int64_t n4) {
  __shared__ float part[32];
  __shared__ float cpart[FM_MSE_CTAS];
  float acc = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint2 pr = __ldcs(pred + i);
    const float4 a = __ldcs(x0 + i);
    const float4 b = __ldcs(x1 + i);
    const float2 p0 = unpack_bf16x2(pr.x), p1 = unpack_bf16x2(pr.y);
    const float d0 = p0.x - (a.x - b.x), d1 = p0.y - (a.y - b.y);
    const float d2 = p1.x - (a.z - b.z), d3 = p1.y - (a.w - b.w);
    acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    if (dpred) {
      const float g = 2.f * gscale * inv_numel;
      uint2 o;
      o.x = pack_bf16x2(g * d0, g * d1);
      o.y = pack_bf16x2(g * d2, g * d3);
      dpred[i] = o;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = part[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0)   // my partial -> slot [rank] of CTA 0
      st_shared_cluster_u32(mapa_u32(smem_u32(&cpart[cluster_ctarank()]), 0u), __float_as_uint(v));
  }
  cluster_sync_all();       // (release / acquire at cluster scope: the remote stores are visible to CTA 0)
  if (cluster_ctarank() == 0 && threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < FM_MSE_CTAS; ++r) t += cpart[r];
    atomicAdd(loss, t * inv_numel);   // ONE add per launch
  }
}

// Batched strided cast-copy: one launch moves a whole table of small matrices (LoRA A/B packs fp32 -> bf16 before the
// step, LoRA / bias gradients fp32 temp -> .grad after backward).  blockIdx.y = descriptor, blockIdx.x strides over
// its elements; consecutive threads touch consecutive columns (coalesced on both sides).
__device__ __forceinline__ int64_t grouped_row(int r, int group, int pitch) {
  return group > 0 ? static_cast<int64_t>(r / group) * pitch + r % group : r;
}
__global__ void __launch_bounds__(256) batched_copy_kernel(const gh_copy_desc* __restrict__ descs) {
  const gh_copy_desc d = descs[blockIdx.y];
  const int64_t n = static_cast<int64_t>(d.rows) * d.cols;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int r = static_cast<int>(i / d.cols), c = static_cast<int>(i - static_cast<int64_t>(r) * d.cols);
    const int64_t so = grouped_row(r, d.src_row_group, d.src_row_pitch) * d.src_ld + grouped_row(c, d.src_col_group, d.src_col_pitch);
    const int64_t dof = grouped_row(r, d.dst_row_group, d.dst_row_pitch) * d.dst_ld + grouped_row(c, d.dst_col_group, d.dst_col_pitch);
    float v = d.src_dtype == GH_F32 ? static_cast<const float*>(d.src)[so]
                                    : __bfloat162float(static_cast<const __nv_bfloat16*>(d.src)[so]);
    v *= d.scale;
    if (d.dst_dtype == GH_F32) {
      float* o = static_cast<float*>(d.dst) + dof;
      *o = d.accumulate ? *o + v : v;
    } else {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(d.dst) + dof;
      *o = __float2bfloat16(d.accumulate ? __bfloat162float(*o) + v : v);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// LoRA input dropout (peft lora_dropout = 0.1 in every stage-2 YAML; train_SigLIP_stage2_all.py:139).  Counter-based
// Philox-4x32-10 keyed by (seed, call offset): the backward regenerates the mask instead of storing it.
// Algorithmic bytes / element: fwd 2 + 2, bwd 2 + 2 + 2.
// ----------------------------------------------------------------------------------------------------------
// One thread = 8 consecutive elements (one 16-byte access) = ONE Philox block: each element takes 16 of its 128 random
// bits (keep iff bits >= p * 2^16: the drop probability is quantised to 1/65536, 0.1 -> 0.100006).  The first version
// spent a whole Philox block and an 8-byte access per 4 elements and ran at 2.9 TB/s, compute-bound on the ten
// multiply-xor rounds; this form halves the rounds per element and doubles the access width.
template <bool BWD_ADD>
__global__ void __launch_bounds__(256) dropout_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n8,
                                                      uint32_t thresh16, float inv_keep, uint2 key, unsigned long long offset,
                                                      const unsigned long long* __restrict__ offset_base) {
  // the mask is keyed by offset + *offset_base: the by-value part is the call index inside a step, the device-resident
  // part advances once per step -- a CUDA graph bakes by-value arguments in, so without it every replay would draw
  // the SAME masks
  const unsigned long long o64 = offset + (offset_base != nullptr ? *offset_base : 0ull);
  const uint2 off = make_uint2(static_cast<uint32_t>(o64), static_cast<uint32_t>(o64 >> 32));
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), off.x, off.y), key);
    const uint4 v = in[i];
    const uint32_t rw[4] = {r.x, r.y, r.z, r.w}, vw[4] = {v.x, v.y, v.z, v.w};
    uint32_t ow[4];
    uint4 d = make_uint4(0u, 0u, 0u, 0u);
    if (BWD_ADD) d = out[i];
    const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 a = unpack_bf16x2(vw[q]);
      float o0 = (rw[q] & 0xffffu) >= thresh16 ? a.x * inv_keep : 0.f;
      float o1 = (rw[q] >> 16) >= thresh16 ? a.y * inv_keep : 0.f;
      if (BWD_ADD) {
        const float2 c = unpack_bf16x2(dw[q]);
        o0 += c.x;
        o1 += c.y;
      }
      ow[q] = pack_bf16x2(o0, o1);
    }
    out[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// Euler step of the flow sampler with the classifier-free-guidance mix fused (src/flux/sampling.py:143-146):
//   v = neg ? neg + gs * (pred - neg) : pred ;  x += dt * v      (bf16 in the reference's bf16 arithmetic order:
//   every intermediate is rounded to bf16 as torch does for bf16 tensors).  4 elements per thread.
__global__ void __launch_bounds__(256) euler_step_kernel(uint2* __restrict__ x, const uint2* __restrict__ pred,
                                                         const uint2* __restrict__ neg, float dt, float gs, int64_t n4) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto r = [](float v) { return __bfloat162float(__float2bfloat16(v)); };
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint2 xv = x[i], pv = pred[i];
    float xs[4] = {unpack_bf16x2(xv.x).x, unpack_bf16x2(xv.x).y, unpack_bf16x2(xv.y).x, unpack_bf16x2(xv.y).y};
    float ps[4] = {unpack_bf16x2(pv.x).x, unpack_bf16x2(pv.x).y, unpack_bf16x2(pv.y).x, unpack_bf16x2(pv.y).y};
    if (neg) {
      const uint2 nv = neg[i];
      const float ns[4] = {unpack_bf16x2(nv.x).x, unpack_bf16x2(nv.x).y, unpack_bf16x2(nv.y).x, unpack_bf16x2(nv.y).y};
#pragma unroll
      for (int k = 0; k < 4; ++k) ps[k] = r(ns[k] + r(gs * r(ps[k] - ns[k])));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) xs[k] = r(xs[k] + r(dt * ps[k]));
    x[i] = make_uint2(pack_bf16x2(xs[0], xs[1]), pack_bf16x2(xs[2], xs[3]));
  }
}

static inline int ew_grid(int64_t n_items, int block) {
  const int64_t want = (n_items + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;  // 8 resident CTAs of 256 threads per SM
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace gh

extern "C" int gh_fm_interp_fwd(const float* x1, const float* x0, const float* t, void* xt_bf16, int64_t batch,
                                int64_t per_sample, void* stream) {
  using namespace gh;
  GH_REQUIRE(x1 && x0 && t && xt_bf16, GH_ERR_NULL, "gh_fm_interp_fwd: NULL pointer");
  GH_REQUIRE(batch >= 0 && per_sample >= 0, GH_ERR_BAD_SHAPE, "gh_fm_interp_fwd: negative size");
  if (batch == 0 || per_sample == 0) return GH_OK;
  GH_REQUIRE(per_sample % 4 == 0, GH_ERR_BAD_SHAPE, "gh_fm_interp_fwd: per_sample=%lld must be a multiple of 4",
             (long long)per_sample);
  GH_REQUIRE(aligned16(x1) && aligned16(x0) && aligned16(xt_bf16), GH_ERR_ALIGN, "gh_fm_interp_fwd: 16B alignment");
  const int64_t n4 = batch * per_sample / 4;
  fm_interp_kernel<<<ew_grid(n4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x1), reinterpret_cast<const float4*>(x0), t,
      reinterpret_cast<uint2*>(xt_bf16), n4, per_sample / 4);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_fm_mse_loss_fwdbwd(const void* pred_bf16, const float* x0, const float* x1, float* loss_accum,
                                     void* dpred_bf16, float grad_scale, int64_t numel, void* stream) {
  using namespace gh;
  GH_REQUIRE(pred_bf16 && x0 && x1 && loss_accum, GH_ERR_NULL, "gh_fm_mse_loss_fwdbwd: NULL pointer");
  GH_REQUIRE(numel > 0 && numel % 4 == 0, GH_ERR_BAD_SHAPE,
             "gh_fm_mse_loss_fwdbwd: numel=%lld must be a positive multiple of 4", (long long)numel);
  GH_REQUIRE(aligned16(x0) && aligned16(x1) && aligned16(pred_bf16) && (!dpred_bf16 || aligned16(dpred_bf16)),
             GH_ERR_ALIGN, "gh_fm_mse_loss_fwdbwd: 16B alignment");
  const int64_t n4 = numel / 4;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(FM_MSE_CTAS);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = FM_MSE_CTAS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  GH_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fm_mse_kernel, reinterpret_cast<const uint2*>(pred_bf16),
                                   reinterpret_cast<const float4*>(x0), reinterpret_cast<const float4*>(x1), loss_accum,
                                   reinterpret_cast<uint2*>(dpred_bf16), grad_scale, 1.0f / static_cast<float>(numel), n4));
  return GH_OK;
}

extern "C" int gh_batched_copy(const gh_copy_desc* descs_device, int32_t n_desc, int32_t blocks_per_desc, void* stream) {
  using namespace gh;
  GH_REQUIRE(n_desc >= 0 && blocks_per_desc > 0, GH_ERR_BAD_SHAPE, "gh_batched_copy: bad launch shape");
  if (n_desc == 0) return GH_OK;
  GH_REQUIRE(descs_device != nullptr, GH_ERR_NULL, "gh_batched_copy: NULL descriptor table");
  GH_REQUIRE(n_desc <= 65535, GH_ERR_BAD_SHAPE, "gh_batched_copy: at most 65535 descriptors per launch");
  batched_copy_kernel<<<dim3(blocks_per_desc, n_desc), 256, 0, static_cast<cudaStream_t>(stream)>>>(descs_device);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

static int launch_dropout(bool bwd_add, const void* in, void* out, int64_t numel, float p, uint64_t seed, uint64_t offset,
                          const uint64_t* offset_base, void* stream, const char* who) {
  using namespace gh;
  GH_REQUIRE(in && out, GH_ERR_NULL, "%s: NULL pointer", who);
  GH_REQUIRE(numel >= 0 && numel % 8 == 0, GH_ERR_BAD_SHAPE, "%s: numel=%lld must be a multiple of 8", who, (long long)numel);
  GH_REQUIRE(p >= 0.f && p < 1.f, GH_ERR_BAD_SHAPE, "%s: p=%f outside [0, 1)", who, p);
  GH_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, GH_ERR_ALIGN,
             "%s: 16-byte alignment", who);
  if (numel == 0) return GH_OK;
  const int64_t n8 = numel / 8;
  const double t = static_cast<double>(p) * 65536.0 + 0.5;
  const uint32_t thresh = t >= 65535.0 ? 65535u : static_cast<uint32_t>(t);
  const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const unsigned long long off = offset;
  const unsigned long long* base = reinterpret_cast<const unsigned long long*>(offset_base);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bwd_add)
    dropout_kernel<true><<<ew_grid(n8, 256), 256, 0, s>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), n8, thresh,
                                                          1.f / (1.f - p), key, off, base);
  else
    dropout_kernel<false><<<ew_grid(n8, 256), 256, 0, s>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), n8, thresh,
                                                           1.f / (1.f - p), key, off, base);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_dropout_fwd(const void* x_bf16, void* y_bf16, int64_t numel, float p, uint64_t seed, uint64_t offset,
                              const uint64_t* offset_base, void* stream) {
  return launch_dropout(false, x_bf16, y_bf16, numel, p, seed, offset, offset_base, stream, "gh_dropout_fwd");
}
extern "C" int gh_dropout_bwd_add(const void* t_bf16, void* dx_bf16, int64_t numel, float p, uint64_t seed, uint64_t offset,
                                  const uint64_t* offset_base, void* stream) {
  return launch_dropout(true, t_bf16, dx_bf16, numel, p, seed, offset, offset_base, stream, "gh_dropout_bwd_add");
}

extern "C" int gh_euler_cfg_step(void* x_bf16, const void* pred_bf16, const void* neg_pred_bf16, float dt, float true_gs,
                                 int64_t numel, void* stream) {
  using namespace gh;
  GH_REQUIRE(x_bf16 && pred_bf16, GH_ERR_NULL, "gh_euler_cfg_step: NULL pointer");
  GH_REQUIRE(numel >= 0 && numel % 4 == 0, GH_ERR_BAD_SHAPE, "gh_euler_cfg_step: numel=%lld must be a multiple of 4",
             (long long)numel);
  if (numel == 0) return GH_OK;
  const int64_t n4 = numel / 4;
  euler_step_kernel<<<ew_grid(n4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint2*>(x_bf16), static_cast<const uint2*>(pred_bf16), static_cast<const uint2*>(neg_pred_bf16), dt,
      true_gs, n4);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}
