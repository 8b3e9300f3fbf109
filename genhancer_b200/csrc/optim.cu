// genhancer_b200 -- HBM-bound optimizer kernels over FLAT parameter / gradient buffers:
//   gh_sumsq_accum : acc += sum(g^2)                (global grad-norm, accelerator.clip_grad_norm_,
//                                                     train_SigLIP_stage1.py:271-272); bit-reproducible
//   gh_adamw_step  : clip-by-global-norm + AdamW    (torch.optim.AdamW(lr, betas, eps, weight_decay),
//                                                     train_SigLIP_stage1.py:147-153,273)
// Parameters, gradients and both moment buffers share one dtype (the reference keeps bf16 Adam states for the
// bf16 DiT and fp32 states for the fp32 projectors: no fp32 master copy).  Arithmetic is fp32 with
// ONE rounding on store (the semantics of torch's fused AdamW).  The clip coefficient is computed on the device
// from the accumulated squared norm, so the optimizer step never synchronises with the host.
// algorithmic bytes / parameter: sumsq 1 read; adamw 4 reads + 3 writes  (x2 B bf16, x4 B fp32)
#include "common.cuh"
#include "internal.h"

namespace gh {

template <typename T>
struct Vec;
template <>
struct Vec<__nv_bfloat16> {  // 8 x bf16 = 16 B
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
  static __device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <>
struct Vec<float> {  // 4 x fp32 = 16 B
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ float ld1(const float* p) { return *p; }
  static __device__ __forceinline__ void st1(float* p, float v) { *p = v; }
};

// Deterministic: every block leaves its partial sum in ws[blockIdx.x]; the block that draws the last ticket adds the
// partials up in a FIXED order (independent of which block that is) and adds the total to *acc.  The same gradients
// therefore give bit-identical norms -- and, through the clip coefficient, bit-identical AdamW updates -- on every
// rank of a data-parallel job (an atomicAdd per block rounded differently from rank to rank, and the replicas drifted
// apart by an ulp per step).  ws = [gridDim.x partials | 1 ticket counter], zeroed once by the caller; the last block
// hands the counter back zeroed.
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ g, int64_t n, float* __restrict__ acc,
                                                    float* __restrict__ ws) {
  constexpr int V = Vec<T>::N;
  float s = 0.f;
  const int64_t nv = n / V;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float v[V];
    Vec<T>::load(g + i * V, v);
#pragma unroll
    for (int j = 0; j < V; ++j) s += v[j] * v[j];
  }
  if (blockIdx.x == 0 && threadIdx.x < n - nv * V) {  // scalar tail
    const float t = Vec<T>::ld1(g + nv * V + threadIdx.x);
    s += t * t;
  }
  s = warp_sum(s);
  __shared__ float part[8];
  __shared__ int last;
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      ws[blockIdx.x] = v;
      __threadfence();
      int* ticket = reinterpret_cast<int*>(ws + gridDim.x);
      last = atomicAdd(ticket, 1) == static_cast<int>(gridDim.x) - 1;
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += blockDim.x) t += __ldcg(ws + i);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i];
    *acc += tot;
    *reinterpret_cast<int*>(ws + gridDim.x) = 0;
  }
}

struct AdamWArgs {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale, max_norm;
  // Device-resident step state (NULL: `bc1` / `bc2_sqrt` above were computed by the host from the by-value step):
  // dev_state[0] = optimizer step count (>= 1), dev_state[1] = 0 -> this launch is a no-op.  Lets the update be
  // captured into the CUDA graph of the training step (a graph bakes by-value arguments in) and be skipped on
  // the replay that follows an eager flush (checkpoint / end of training).
  const int32_t* dev_state;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a, float clip) {
  g *= clip;
  p *= (1.f - a.lr * a.weight_decay);
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
}

template <typename T>
__global__ void __launch_bounds__(256) adamw_kernel(T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m,
                                                    T* __restrict__ v, int64_t n, const float* __restrict__ gnorm_sq,
                                                    AdamWArgs a) {
  constexpr int V = Vec<T>::N;
  if (a.dev_state != nullptr) {
    if (a.dev_state[1] == 0) return;
    const double st = static_cast<double>(a.dev_state[0]);   // once per thread: bias corrections from the device step
    a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(a.beta1), st));
    a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(a.beta2), st)));
  }
  // clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6)); norm of the grad_scale-d gradients
  float clip = a.grad_scale;
  if (gnorm_sq != nullptr && a.max_norm > 0.f) {
    const float norm = sqrtf(*gnorm_sq) * a.grad_scale;
    clip *= fminf(1.f, a.max_norm / (norm + 1e-6f));
  }
  const int64_t nv = n / V;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float pv[V], gv[V], mv[V], vv[V];
    Vec<T>::load(p + i * V, pv);
    Vec<T>::load(g + i * V, gv);
    Vec<T>::load(m + i * V, mv);
    Vec<T>::load(v + i * V, vv);
#pragma unroll
    for (int j = 0; j < V; ++j) adamw_one(pv[j], gv[j], mv[j], vv[j], a, clip);
    Vec<T>::store(p + i * V, pv);
    Vec<T>::store(m + i * V, mv);
    Vec<T>::store(v + i * V, vv);
  }
  if (blockIdx.x == 0 && threadIdx.x < n - nv * V) {
    const int64_t i = nv * V + threadIdx.x;
    float pp = Vec<T>::ld1(p + i), mm = Vec<T>::ld1(m + i), vv = Vec<T>::ld1(v + i);
    adamw_one(pp, Vec<T>::ld1(g + i), mm, vv, a, clip);
    Vec<T>::st1(p + i, pp); Vec<T>::st1(m + i, mm); Vec<T>::st1(v + i, vv);
  }
}

static int grid_for(int64_t n_vec) {
  const int64_t want = (n_vec + 255) / 256;
  const int64_t cap = 8L * num_sms();
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace gh

using namespace gh;

extern "C" int64_t gh_sumsq_workspace_bytes() { return (8L * num_sms() + 1) * static_cast<int64_t>(sizeof(float)); }

extern "C" int gh_sumsq_accum(const void* g, int32_t dtype, int64_t numel, float* acc, float* ws, void* stream) {
  GH_REQUIRE(g && acc && ws, GH_ERR_NULL, "gh_sumsq_accum: NULL pointer");
  GH_REQUIRE(numel > 0, GH_ERR_BAD_SHAPE, "gh_sumsq_accum: numel must be positive");
  GH_REQUIRE(aligned16(g), GH_ERR_ALIGN, "gh_sumsq_accum: buffer must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == GH_BF16)
    sumsq_kernel<__nv_bfloat16><<<grid_for(numel / 8), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(g), numel, acc, ws);
  else if (dtype == GH_F32)
    sumsq_kernel<float><<<grid_for(numel / 4), 256, 0, s>>>(static_cast<const float*>(g), numel, acc, ws);
  else
    return set_error(GH_ERR_UNSUPPORTED, "gh_sumsq_accum: dtype %d", dtype);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_adamw_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int32_t dtype, int64_t numel,
                             float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                             const float* gnorm_sq, float max_norm, float grad_scale, const int32_t* dev_state,
                             void* stream) {
  GH_REQUIRE(param && grad && exp_avg && exp_avg_sq, GH_ERR_NULL, "gh_adamw_step: NULL pointer");
  GH_REQUIRE(numel > 0 && (step >= 1 || dev_state), GH_ERR_BAD_SHAPE,
             "gh_adamw_step: numel must be positive and step >= 1 (or dev_state given)");
  if (step < 1) step = 1;
  GH_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), GH_ERR_ALIGN,
             "gh_adamw_step: buffers must be 16-byte aligned");
  AdamWArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
  a.grad_scale = grad_scale; a.max_norm = max_norm; a.dev_state = dev_state;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == GH_BF16)
    adamw_kernel<__nv_bfloat16><<<grid_for(numel / 8), 256, 0, s>>>(
        static_cast<__nv_bfloat16*>(param), static_cast<const __nv_bfloat16*>(grad), static_cast<__nv_bfloat16*>(exp_avg),
        static_cast<__nv_bfloat16*>(exp_avg_sq), numel, gnorm_sq, a);
  else if (dtype == GH_F32)
    adamw_kernel<float><<<grid_for(numel / 4), 256, 0, s>>>(static_cast<float*>(param), static_cast<const float*>(grad),
                                                          static_cast<float*>(exp_avg), static_cast<float*>(exp_avg_sq),
                                                          numel, gnorm_sq, a);
  else
    return set_error(GH_ERR_UNSUPPORTED, "gh_adamw_step: dtype %d", dtype);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}
