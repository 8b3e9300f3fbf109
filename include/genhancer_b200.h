/* genhancer_b200 -- C ABI of the B200-native GenHancer hot path.
 *
 * The reference (Jam1ezhang/GenHancer) has no FFI of its own: every "kernel" is a
 * torch/ATen library call made from Python.  The entry
 * points below are the operators a maintainer would bind (ctypes stub in
 * INTEGRATION.md) to replace those calls; each one cites the reference site(s)
 * it replaces, relative to /root/reference/Continuous unless noted.
 *
 * Conventions
 *   - every function returns 0 (GH_OK) or a negative gh_status; gh_last_error()
 *     returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch); the library
 *     never allocates, frees or synchronises.  `stream` is a cudaStream_t passed as
 *     void* (NULL = legacy default stream).
 *   - bf16 tensors are raw uint16 payloads; "ld" = leading dimension in ELEMENTS.
 *   - dtype codes: GH_BF16 = 0, GH_F32 = 1.
 */
#ifndef GENHANCER_B200_H_
#define GENHANCER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  GH_OK = 0,
  GH_ERR_BAD_SHAPE = -1,
  GH_ERR_UNSUPPORTED = -2,
  GH_ERR_CUDA = -3,
  GH_ERR_ALIGN = -4,
  GH_ERR_NULL = -5
} gh_status;

enum { GH_BF16 = 0, GH_F32 = 1 };
enum { GH_ACT_NONE = 0, GH_ACT_GELU_TANH = 1, GH_ACT_QUICK_GELU = 2, GH_ACT_GELU_ERF = 3, GH_ACT_SILU = 4 };

const char* gh_last_error(void);
int gh_version(void);
/* Resolve the driver entry points, raise the kernels' dynamic-smem limits on `device`. */
int gh_init(int device);

/* --------------------------------------------------------------------------
 * gh_gemm_bf16 -- tcgen05/TMEM GEMM fed by TMA, fused epilogue.
 *   acc[m,n] = alpha * sum_k A(m,k) * B(n,k)                    (bf16 x bf16 -> fp32)
 *   v        = acc + bias[n]
 *   aux_out[m,n] = v                                   (optional: saved pre-activation)
 *   v        = act_grad ? v * act'(aux_in[m,n]) : act(v)
 *   v        = gate[(m / rows_per_batch), n] * v               (optional AdaLN gate)
 *   D[m,n]   = v + residual[m,n]                               (optional)
 * Operand layouts: a_mn_major = 0 -> A is [M,K] row-major (lda >= K);
 *                  a_mn_major = 1 -> A is stored [K,M] row-major (lda >= M).  Same for B
 *                  with [N,K] / [K,N].  One kernel therefore serves
 *                  forward  (Y = X W^T       : A K-major,  B K-major),
 *                  dgrad    (dX = dY W       : A K-major,  B MN-major) and
 *                  wgrad    (dW = dY^T X     : A MN-major, B MN-major).
 * Replaces: every nn.Linear on the path -- HF CLIP q/k/v/out_proj/fc1/fc2
 *   (transformers modeling_clip.py:294-351), CLIP_bank.py:17-28 projectors,
 *   src/flux/modules/layers.py:52-60 (MLPEmbedder), :169-175 (Modulation.lin),
 *   :311,318,331,335 (double-block qkv/proj/mlp), :490,499 (single-block linear1/2),
 *   :568-572 (LastLayer), src/flux/model.py:154,161 (img_in/txt_in) -- and their
 *   autograd backward (train_SigLIP_stage1.py:270).
 * Requirements: N % 8 == 0, all ld % 8 == 0, pointers 16-byte aligned.
 * -------------------------------------------------------------------------- */
typedef struct {
  const void* a;
  int64_t lda;
  int32_t a_mn_major;
  const void* b;
  int64_t ldb;
  int32_t b_mn_major;
  void* d;
  int64_t ldd;
  int32_t d_dtype;
  int32_t M, N, K;
  float alpha;
  const void* bias; /* [N] or NULL */
  int32_t bias_dtype;
  int32_t act;      /* GH_ACT_* */
  int32_t act_grad; /* 1: multiply by act'(aux_in) instead of applying act */
  const void* aux_in; /* bf16 [M, ld_aux_in] */
  int64_t ld_aux_in;
  void* aux_out; /* bf16 [M, ld_aux_out] */
  int64_t ld_aux_out;
  const void* gate; /* bf16 [M / rows_per_batch, gate_ld] */
  int64_t gate_ld;
  int32_t rows_per_batch;
  const void* residual;
  int64_t ld_res;
  int32_t res_dtype;
} gh_gemm_args;
int gh_gemm_bf16(const gh_gemm_args* args, void* stream);

/* --------------------------------------------------------------------------
 * Flow-matching interpolation (train_SigLIP_stage1.py:248-250,255):
 *   x_t[b,i] = bf16( (1 - t[b]) * x1[b,i] + t[b] * x0[b,i] ),  x0/x1 fp32, t fp32.
 * t and x0 are drawn by the host (torch RNG) so that the draw order of the
 * reference is kept bit-exactly.  per_sample = L*64 elements per batch row.
 * -------------------------------------------------------------------------- */
int gh_fm_interp_fwd(const float* x1, const float* x0, const float* t, void* xt_bf16, int64_t batch,
                     int64_t per_sample, void* stream);

/* --------------------------------------------------------------------------
 * Velocity-MSE loss, forward + gradient in one pass (train_SigLIP_stage1.py:263):
 *   loss  = mean( (float(pred) - (x0 - x1))^2 )
 *   dpred = bf16( grad_scale * 2 * (float(pred) - (x0 - x1)) / numel )
 * loss_accum is ONE fp32 the caller zeroes beforehand (atomicAdd of block partials,
 * already divided by numel).  dpred may be NULL (forward only).
 * -------------------------------------------------------------------------- */
int gh_fm_mse_loss_fwdbwd(const void* pred_bf16, const float* x0, const float* x1, float* loss_accum,
                          void* dpred_bf16, float grad_scale, int64_t numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENHANCER_B200_H_ */
