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
  /* Optional second operand pair, same majors as a/b:  acc += sum over k < K2 of A2(m,k) * B2(n,k)  (before alpha).
   * This is how LoRA (peft r=16, train_SigLIP_stage2_all.py:134-142) is folded into the base GEMM: forward
   * y = x W^T + u B^T with u = (alpha/r) x A^T [M,r]; dgrad dx = dy W + du A with du = (alpha/r) dy B.
   * The rank-r slice costs one extra 16-deep MMA per tile, no second pass over y.  K2 = 0: absent. */
  const void* a2;
  int64_t lda2;
  const void* b2;
  int64_t ldb2;
  int32_t K2;
  /* Split-K for skinny outputs reduced over a long K (LoRA wgrads dA = du^T x, dB = dy^T u: 16..48 rows or columns,
   * K = all tokens of the batch): != 0 makes the call D += A B^T on an fp32 D the CALLER HAS ZEROED (or wants
   * accumulated into), computed by k_splits slices of the K loop that add their partial products with red.add;
   * < 0 picks the slice count that fills the machine.  No epilogue options, no second operand pair. */
  int32_t k_splits;
  /* Batched mode (batch > 1): `batch` independent M x N x K problems in ONE launch over FLAT 2-D operands.  Problem i
   * reads A at row offset i * a_batch_rows (K-major A: rows of the [.,K] matrix; MN-major A: rows of the [K,.] matrix,
   * i.e. the k index), B at i * b_batch_rows likewise, and writes D (and reads residual / aux_in, writes aux_out) at
   * row i * d_batch_rows + m.  Used for the AE mid-block attention (autoencoder.py:37-52: one 1764 x 1764 x 512
   * Q K^T and one P V per image) which is otherwise 64 launch-bound GEMMs.  Rows that a tile reads beyond its own
   * problem belong to the next one (finite values; their products are masked or multiply TMA-zero-filled columns).
   * No second operand pair, split-K or gate.  batch <= 1: plain GEMM. */
  int32_t batch;
  int64_t a_batch_rows, b_batch_rows, d_batch_rows;
  /* Tile schedule of THIS launch's persistent grid: 0 = static (tile = worker + i * #workers: no atomics, no start-up
   * latency), 1 = dynamic (each CTA / CTA pair pulls its next tile from a global counter through a 2-deep queue in
   * shared memory).  Data-parallel training asks for dynamic while gradient buckets are in flight: NCCL's all-reduce
   * kernels hold some SMs, and CTAs of a persistent grid that start a wave late then find the queue drained instead of
   * owing a full static share of the tiles (which doubled the GEMM's duration).  Results are
   * identical (each tile is computed by exactly one worker).  A per-call field, not library state: the policy lives in
   * the host (parallel.GradReducer), the library stays re-entrant across streams. */
  int32_t dynamic_tiles;
} gh_gemm_args;
int gh_gemm_bf16(const gh_gemm_args* args, void* stream);
/* Bring-up aid: when device_buf (int64 [8 * #SMs]) is non-NULL, every following gh_gemm_bf16 launch writes per-CTA
 * cycle counters of its TMA / MMA / epilogue pipelines there (see GemmParams::prof in csrc/umma_gemm.cuh).  NULL
 * switches it off.  Not for production: one process-wide pointer. */
int gh_debug_gemm_prof(void* device_buf);

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


/* --------------------------------------------------------------------------
 * Row views.  Token-major activations [B, L, C] are addressed as rows of C channels:
 * row (b, l) lives at base + b * batch_stride + l * row_stride (ELEMENTS).  This lets the
 * norm kernels read a slice of a concatenated txt|img sequence (Flux.forward drops the
 * txt tokens before the last layer, src/flux/model.py:225) without a copy.
 * rows_per_batch <= 0 means one flat batch.
 * -------------------------------------------------------------------------- */
typedef struct {
  int32_t rows_per_batch;
  int64_t batch_stride;
  int64_t row_stride;
} gh_rows_view;

/* LayerNorm forward, bf16 in/out, fp32 statistics (saved for backward when non-NULL):
 *   affine  (weight,bias fp32 [C])            y = n * w + b      -- HF CLIP/SigLIP layer_norm1/2, pre/post
 *                                                                   layernorm (modeling_clip.py:358-361,659-662),
 *                                                                   projector LN (CLIP_bank.py:18,24), adapter LN
 *   AdaLN   (shift,scale bf16 [B, mod_ld])    y = (1 + scale[b]) * n + shift[b]
 *                                                                -- layers.py:309-310,316-317,332,336,489,570
 *   plain   (all NULL)                        y = n
 * C % 8 == 0, C <= 4096. */
int gh_layernorm_fwd(const void* x, const gh_rows_view* xv, void* y, const gh_rows_view* yv, int32_t rows,
                     int32_t C, const float* weight, const float* bias, const void* shift, const void* scale,
                     int64_t mod_ld, float eps, float* mean_out, float* rstd_out, void* stream);
/* dx = rstd * (dn - mean(dn) - n * mean(dn * n)) [+ dres],  dn = dy * (w | 1 + scale[b] | 1). */
int gh_layernorm_bwd_dx(const void* dy, const gh_rows_view* dyv, const void* x, const gh_rows_view* xv,
                        int32_t rows, int32_t C, const float* mean, const float* rstd, const float* weight,
                        const void* scale, int64_t mod_ld, const void* dres, const gh_rows_view* drv, void* dx,
                        const gh_rows_view* dxv, void* stream);
/* dshift_acc[b,c] += sum_l dy ; dscale_acc[b,c] += sum_l dy * n   (fp32 accumulators, atomics).
 * With batches = 1 and rows_per_batch = all rows these are the affine LayerNorm's (dbias, dweight). */
int gh_layernorm_bwd_params(const void* dy, const gh_rows_view* dyv, const void* x, const gh_rows_view* xv,
                            int32_t batches, int32_t C, const float* mean, const float* rstd, float* dshift_acc,
                            float* dscale_acc, int64_t acc_ld, void* stream);
/* Backward of the gated residual  out = res + gate[b] * u  (layers.py:331-336,500):
 *   du = gate[b] * dout ;  dgate_acc[b,c] += sum_l dout * u ;
 *   dbias_acc (fp32 [C], may be NULL): dbias_acc[c] += sum_{b,l} du -- u is the output of a biased Linear, so du is that
 *   Linear's dY and its bias gradient falls out of this pass (no separate gh_colsum read of du). */
int gh_gate_bwd(const void* dout, const gh_rows_view* dov, const void* u, const gh_rows_view* uv, int32_t batches,
                int32_t C, const void* gate, int64_t gate_ld, void* du, const gh_rows_view* duv, float* dgate_acc,
                int64_t acc_ld, float* dbias_acc, void* stream);
/* acc[b,c] += sum_l dy[b,l,c]  (bias gradients). */
int gh_colsum(const void* dy, const gh_rows_view* dyv, int32_t batches, int32_t C, float* acc, int64_t acc_ld,
              void* stream);

/* RoPE cos/sin table from integer-valued ids (EmbedND + rope, layers.py:18-25, math.py:15-22):
 * ids fp32 [n_tokens, 3] -> cos_sin float2 [n_tokens, (a0+a1+a2)/2]; angles in float64. */
int gh_rope_table(const float* ids, void* cos_sin, int64_t n_tokens, int32_t axis0, int32_t axis1, int32_t axis2,
                  double theta, void* stream);
/* QKNorm (RMSNorm over head dim, layers.py:63-84) + apply_rope (math.py:25-30) + head-major scatter:
 * qkv bf16 [B, L, 3, H, D] (row pitch ld_qkv) -> q, k, v bf16 [B, H, Ltot, D] at token offset l_off
 * (the txt|img concat of layers.py:323-326 is done by writing both streams into the same buffers). D = 128. */
int gh_qk_norm_rope_fwd(const void* qkv, int64_t ld_qkv, int32_t B, int32_t L, int32_t H, int32_t D, int32_t Ltot,
                        int32_t l_off, const void* q_scale, const void* k_scale, const void* cos_sin,
                        int64_t cs_batch_stride, void* q, void* k, void* v, void* stream);
int gh_qk_norm_rope_bwd(const void* dq, const void* dk, const void* dv, const void* qkv, int64_t ld_qkv, int32_t B,
                        int32_t L, int32_t H, int32_t D, int32_t Ltot, int32_t l_off, const void* q_scale,
                        const void* k_scale, const void* cos_sin, int64_t cs_batch_stride, void* dqkv,
                        int64_t ld_dqkv, float* dscale_q_acc, float* dscale_k_acc, void* stream);

/* timestep_embedding (layers.py:28-49): t fp32 [B] -> bf16 [B, 256] = [cos | sin](1000 t f_k).
 * round_bf16 = 1 reproduces the scripts' bf16 cast of t before the multiply (train_SigLIP_stage1.py:260). */
int gh_timestep_embedding(const float* t, void* out_bf16, int32_t B, int32_t round_bf16, void* stream);
/* elementwise activation / its backward on bf16 (nn.SiLU of Modulation / LastLayer, layers.py:170,566). */
int gh_act_fwd(const void* x, void* y, int64_t numel, int32_t act, void* stream);
int gh_act_bwd(const void* dy, const void* x, void* dx, int64_t numel, int32_t act, void* stream);
/* dst (bf16|fp32) = scale * src_fp32 (+ dst): flush of the fp32 small-gradient scratch into .grad. */
int gh_accum_cast(const float* src, void* dst, int32_t dst_dtype, int64_t numel, float scale, int32_t accumulate,
                  void* stream);

/* One Euler step of the flow sampler, classifier-free-guidance mix fused (src/flux/sampling.py:129-146):
 *   v = neg_pred ? neg_pred + true_gs * (pred - neg_pred) : pred ;  x += dt * v     (dt = t_prev - t_curr)
 * bf16, in place on x; neg_pred may be NULL.  numel % 4 == 0. */
int gh_euler_cfg_step(void* x_bf16, const void* pred_bf16, const void* neg_pred_bf16, float dt, float true_gs,
                      int64_t numel, void* stream);

/* Batched strided cast-copy: dst = scale * src (+ dst) for a DEVICE-resident table of small matrices, one launch.
 * Logical row (column) i of a matrix lives at storage row (column) (i / group) * pitch + i % group when group > 0
 * (heads of 72 / 80 features stored in 128-wide slots for the attention kernels), else at i.  Used to refresh the bf16 operand
 * copies of the fp32 LoRA A / B parameters once per step and to scatter LoRA / bias gradients from the fp32 GEMM
 * outputs into .grad (peft LoRA layer + torch.autograd, train_SigLIP_stage2_all.py:134-142,290). */
typedef struct {
  const void* src;
  void* dst;
  int32_t rows, cols;
  int64_t src_ld, dst_ld;
  int32_t src_dtype, dst_dtype;
  int32_t src_row_group, src_row_pitch, src_col_group, src_col_pitch;
  int32_t dst_row_group, dst_row_pitch, dst_col_group, dst_col_pitch;
  float scale;
  int32_t accumulate;
} gh_copy_desc;
int gh_batched_copy(const gh_copy_desc* descs_device, int32_t n_desc, int32_t blocks_per_desc, void* stream);

/* LoRA input dropout (peft lora.Linear: lora_B(lora_A(dropout(x))), lora_dropout 0.1 in the stage-2 YAMLs,
 * train_SigLIP_stage2_all.py:139).  y = keep ? x / (1-p) : 0 on bf16; the keep mask is a pure function of
 * (seed, offset, element index) -- Philox-4x32-10 -- so the backward regenerates it:
 * gh_dropout_bwd_add: dx += keep ? t / (1-p) : 0.  numel % 8 == 0, contiguous, 16-byte aligned
 * (one Philox block per 8 elements, 16 random bits each: p is quantised to 1/65536).
 * offset_base (device uint64, may be NULL) is added to `offset` on the device: the caller advances it once per step
 * (a stream-ordered add), so a CUDA-graph replay -- which bakes the by-value `offset` in -- still draws fresh masks,
 * and the backward of the same step regenerates the forward's. */
/* The LoRA branch around that mask, fused (csrc/lora_fused.cu; replaces peft lora.Linear's dropout -> lora_A and the
 * backward of that pair, peft/tuners/lora/layer.py Linear.forward):
 *   gh_lora_dropout_fwd : xd = drop(x) [M,K] AND u[:, 0:R] = alpha * xd A^T (A [R,K] bf16 row-major) in one pass over x;
 *                         RX = 16 also writes u[:, R:R+16] = [1, 0, ...] (the bias gradient then rides in the dB GEMM).
 *   gh_lora_dropout_bwd : dx [M,K] += drop'(du [M,R] A): one read-modify-write of dx, the product never exists in memory;
 *                         act_pre != NULL: the result is then multiplied by act'(act_pre) (contiguous [M,K] bf16) -- the
 *                         dgrad through an MLP's activation, which the masked term keeps out of the GEMM's epilogue.
 * Same mask as gh_dropout_fwd for the same (p, seed, offset, offset_base) on the contiguous [M,K] tensor.  K % 8 == 0,
 * R in {16, 32, 48}, ldu (row pitch of u / du) % 8 == 0. */
int gh_lora_dropout_fwd(const void* x_bf16, void* xd_bf16, const void* a_bf16, void* u_bf16, int32_t M, int32_t K, int32_t R,
                        int32_t RX, int64_t ldu, float alpha, float p, uint64_t seed, uint64_t offset,
                        const uint64_t* offset_base, void* stream);
int gh_lora_dropout_bwd(const void* du_bf16, const void* a_bf16, void* dx_bf16, int32_t M, int32_t K, int32_t R, int64_t ldu,
                        float p, uint64_t seed, uint64_t offset, const uint64_t* offset_base, const void* act_pre_bf16,
                        int32_t act, void* stream);
int gh_dropout_fwd(const void* x_bf16, void* y_bf16, int64_t numel, float p, uint64_t seed, uint64_t offset,
                   const uint64_t* offset_base, void* stream);
int gh_dropout_bwd_add(const void* t_bf16, void* dx_bf16, int64_t numel, float p, uint64_t seed, uint64_t offset,
                       const uint64_t* offset_base, void* stream);

/* --------------------------------------------------------------------------
 * Flash attention (tcgen05 S/O accumulators in TMEM, TMA-staged tiles, online softmax).
 * Replaces F.scaled_dot_product_attention at HF modeling_clip.py:319-331 (ViT, D=64, scale 1/8) and
 * src/flux/math.py:9 (DiT joint attention over cat(txt,img), D=128), no mask, no dropout.
 *
 * q/k/v element [b,h,l,0:D] is contiguous bf16 at ptr + b*batch_stride + h*head_stride + l*row_stride
 * (ELEMENTS; multiples of 8) -- head-major buffers and the fused-QKV GEMM output are both zero-copy.
 * O is token-major [b, l, h*D + d] and may be split into two row segments (rows < n_split -> seg0).
 * lse2 [B,H,Lq] fp32 = log2-domain logsumexp of the scaled scores (saved for backward; may be NULL).
 * -------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;
  int64_t batch_stride, head_stride, row_stride;
} gh_attn_tensor;
typedef struct {
  void* seg0;
  int64_t seg0_batch_stride, seg0_row_stride;
  void* seg1;
  int64_t seg1_batch_stride, seg1_row_stride;
  int32_t n_split;
} gh_attn_out;
/* d_valid (multiple of 16, <= D; 0 = D): leading lanes of the head dim that hold data.  Heads narrower than the 64- /
 * 128-lane slot they are stored in (SigLIP 72, MetaCLIP-H 80: zero-padded) run only d_valid / 16 of the k-steps of the
 * MMAs that reduce over the head dim and N = d_valid in the ones that produce it; pad lanes of O / dQ / dK / dV are
 * written as exact zeros. */
int gh_flash_attn_fwd(const gh_attn_tensor* q, const gh_attn_tensor* k, const gh_attn_tensor* v, int32_t B, int32_t H,
                      int32_t Lq, int32_t Lk, int32_t D, int32_t d_valid, float scale, const gh_attn_out* o, float* lse2,
                      void* stream);
/* Backward of the above (autograd of math.py:9 / modeling_clip.py:319-331): three launches --
 * prep (delta = rowsum(dO*O), dO gathered head-major), dK/dV pass, dQ pass (S and dP recomputed; no atomics, so the
 * result is bit-reproducible).  128 x 128 score tiles; P^T / dS^T / dS reach the accumulating MMAs through TMEM; dq/dk/dv
 * leave through bulk tensor stores (csrc/attn_bwd2.cuh).
 * o / d_o are token-major with the same two-segment split as the forward output; dq/dk/dv are written at
 * [b,h,l,:] = ptr + b*batch_stride + h*head_stride + l*row_stride.
 * Workspaces: ws_do_headmajor bf16 [B,H,Lq,D], ws_delta fp32 [B,H,Lq]. */
int gh_flash_attn_bwd(const gh_attn_tensor* q, const gh_attn_tensor* k, const gh_attn_tensor* v, const gh_attn_out* o,
                      const gh_attn_out* d_o, const float* lse2, int32_t B, int32_t H, int32_t Lq, int32_t Lk,
                      int32_t D, int32_t d_valid, float scale, const gh_attn_tensor* dq, const gh_attn_tensor* dk,
                      const gh_attn_tensor* dv, void* ws_do_headmajor, float* ws_delta, void* stream);
/* Bring-up aid (like gh_debug_gemm_prof): a device buffer of 16 int64 that CTA (0,0,0) of the FIRST-FORM dK/dV kernel
 * (A/B builds with -DGH_ATTN_BWD_V1; the product kernels of attn_bwd2.cuh carry no counters) fills with cycle counters of its
 * pipeline phases; NULL (the default) switches it off.  Process-global, not for production. */
int gh_debug_attn_prof(void* device_buf);
/* bytes of the two workspaces of gh_flash_attn_bwd: [0] = ws_do_headmajor, [1] = ws_delta */
int64_t gh_flash_attn_bwd_workspace_bytes(int32_t B, int32_t H, int32_t Lq, int32_t D, int32_t which);

/* --------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05, NHWC bf16 activations (frozen FLUX AE encoder,
 * src/flux/modules/autoencoder.py:62-67 ResnetBlock convs, :85-95 Downsample, :157 conv_out).
 *   y[b,oh,ow,co] = act( sum_{kh,kw,ci} x[b, oh*stride+kh-pad, ow*stride+kw-pad, ci] * w[co,(kh*KW+kw)*Cin+ci] + bias[co] )
 *                   + residual[b,oh,ow,co]
 * Out-of-image taps read zero (TMA out-of-bounds fill): `pad` is the top/left padding; right/bottom padding
 * is implied by (Ho, Wo).  Downsample = stride 2, pad 0, Ho = H/2.  Cin % 64 == 0, Cout % 8 == 0.
 * -------------------------------------------------------------------------- */
typedef struct {
  const void* x; /* bf16 [B,H,W,Cin] */
  const void* w; /* bf16 [Cout, KH*KW*Cin] */
  void* y;       /* [B,Ho,Wo,Cout] */
  int32_t y_dtype;
  int32_t B, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo;
  const void* bias;
  int32_t bias_dtype;
  int32_t act;
  const void* residual; /* [B,Ho,Wo,Cout] or NULL */
  int32_t res_dtype;
} gh_conv_args;
int gh_conv2d_nhwc(const gh_conv_args* args, void* stream);

/* patch-embed im2col (stride == kernel): img fp32 NCHW [B,3,S,S] -> bf16 [B*(S/p)^2, ld], k = c*p*p + i*p + j,
 * optional (x-mean)/std with HOST pointers mean3/std3 (transforms.Normalize, train_SigLIP_stage1.py:58).
 * Feeds gh_gemm_bf16 with the flattened Conv2d weight: HF CLIPVisionEmbeddings, modeling_clip.py:147-153,208. */
int gh_patch_im2col(const float* img, void* out_bf16, int32_t B, int32_t S, int32_t patch, int64_t ld,
                    const float* mean3, const float* std3, void* stream);
/* Patch-embed convolution as an IMPLICIT GEMM (HF CLIPVisionEmbeddings / SiglipVisionEmbeddings.patch_embedding:
 * Conv2d(3, D, kernel = stride = patch), modeling_clip.py:147-153,208-209), with ToTensor and transforms.Normalize of the
 * reference's loader / step (dataset_cc3m.py:107-113, train_SigLIP_stage1.py:54-59) folded into the operand gather:
 *   out[(b, py, px), n] = sum_{c,i,j} ((pixel[b, c, py*p+i, px*p+j] - mean[c]) / std[c]) * w[n, c*p*p + i*p + j] + bias[n]
 * img: fp32 NCHW [B,3,S,S] in [0,1] (img_is_u8hwc = 0) or the decoded uint8 HWC [B,S,S,3] batch (1: value / 255 first).
 * w: bf16 [D, ldw] (ldw >= 3*p*p, multiple of 8, pad columns zero); bias fp32 [D] or NULL; out bf16 [B*(S/p)^2, ldo].
 * mean3 / std3: HOST pointers (3 floats) or NULL.  The A operand never exists in HBM: producer warps gather each
 * 128-patch x 64-k tile straight into the swizzled shared-memory layout the tcgen05 MMA reads (a patch row is 56 bytes
 * of fp32 / 14 of uint8 -- not a legal TMA box or stride -- so the tile cannot be a TMA view). */
int gh_patch_embed_fwd(const void* img, int32_t img_is_u8hwc, const void* w_bf16, int64_t ldw, const float* bias,
                       void* out_bf16, int64_t ldo, int32_t B, int32_t S, int32_t patch, int32_t D, const float* mean3,
                       const float* std3, void* stream);
/* The same gather straight from the DECODED image: img_u8 uint8 HWC [B,S,S,3]; every value becomes u8 / 255 (IEEE
 * division = torchvision ToTensor, image_datasets/dataset_cc3m.py:107-113) before (x-mean)/std, so the result is
 * bit-identical to gh_u8hwc_to_f32chw + gh_patch_im2col while the batch is read at 1 byte per value and no fp32 copy
 * of it exists. */
int gh_patch_im2col_u8hwc(const void* img_u8, void* out_bf16, int32_t B, int32_t S, int32_t patch, int64_t ld,
                          const float* mean3, const float* std3, void* stream);
/* 3x3/pad-1 im2col of a 3-channel image for the AE conv_in (autoencoder.py:126): -> bf16 [B*H*W, 32],
 * k = (kh*3+kw)*3 + c, columns 27..31 zero; (x-mean)/std fused (NORMALIZE_VAE, train_SigLIP_stage1.py:59). */
int gh_im2col3x3_c3(const float* img, void* out_bf16, int32_t B, int32_t H, int32_t W, float mean, float std,
                    void* stream);
/* uint8 HWC form of the above (see gh_patch_im2col_u8hwc). */
int gh_im2col3x3_c3_u8hwc(const void* img_u8, void* out_bf16, int32_t B, int32_t H, int32_t W, float mean, float std,
                          void* stream);
/* out[b,t,:] = (t < has_cls ? cls : patch[b,t-has_cls,:]) + pos[t,:]   (modeling_clip.py:210-217) */
int gh_embed_assemble(const void* patch_bf16, const float* cls, const float* pos, void* out_bf16, int32_t B, int32_t T,
                      int32_t D, int32_t has_cls, void* stream);
/* GroupNorm(32 groups) [+ swish] on NHWC bf16 (autoencoder.py:21-22,62-78).  Three launches, no atomics
 * (bit-reproducible): per-CTA partial sums -> fp64 finalize -> normalise + affine (+ swish).
 * ws: gh_groupnorm_ws_bytes(B, HW) bytes of scratch, 16-byte aligned.  C % 64 == 0, C <= 2048. */
int64_t gh_groupnorm_ws_bytes(int32_t B, int64_t HW);
int gh_groupnorm_swish_nhwc(const void* x, void* y, int32_t B, int64_t HW, int32_t C, const float* weight,
                            const float* bias, float eps, int32_t swish, void* ws, void* stream);
/* nearest-neighbour 2x upsample, NHWC bf16 [B,H,W,C] -> [B,2H,2W,C] (FLUX decoder Upsample, autoencoder.py:98-106;
 * the 3x3 conv that follows is gh_conv2d_nhwc).  C % 8 == 0. */
int gh_upsample2x_nhwc(const void* x_bf16, void* y_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* ToTensor of the reference's image pipeline (image_datasets/dataset_cc3m.py:38-44,107-113: torchvision ToTensor on
 * the cropped PIL RGB image): uint8 [B,H,W,3] -> fp32 [B,3,H,W] = value / 255 (IEEE division, bit-identical to torch).
 * Lets the decoded batch cross PCIe as bytes; NORMALIZE_VAE / NORMALIZE_CLIP stay folded into the im2col gathers. */
int gh_u8hwc_to_f32chw(const void* src_u8, float* dst, int32_t B, int32_t H, int32_t W, void* stream);
/* p[r,:n] = softmax(scale * s[r,:n]) fp32 -> bf16, pad columns zeroed (AE mid AttnBlock, autoencoder.py:37-52). */
int gh_softmax_rows(const float* s, int64_t ld_in, void* p_bf16, int64_t ld_out, int32_t rows, int32_t n, float scale,
                    void* stream);
/* DiagonalGaussian sample (noise supplied by the host RNG) + scale/shift + 2x2 patchify
 * (autoencoder.py:268-274,302-305; clip_models/sampling.py:26): -> x1 fp32 [B,(h/2)(w/2),4z]. */
int gh_ae_sample_patchify(const float* moments_nhwc, const float* noise_nchw, float* x1, int32_t B, int32_t h, int32_t w,
                          int32_t z, float scale_factor, float shift_factor, void* stream);

/* --------------------------------------------------------------------------
 * Optimizer step over FLAT buffers (one launch per dtype group instead of one per tensor).
 * gh_sumsq_accum: *acc += sum(g^2) -- the global gradient norm of accelerator.clip_grad_norm_
 *   (train_SigLIP_stage1.py:271-272); acc is ONE fp32 the caller zeroes.  Bit-reproducible: block partials go through
 *   `ws` (gh_sumsq_workspace_bytes() bytes, zeroed ONCE by the caller, reusable by calls on the same stream) and are
 *   added up in a fixed order, so every data-parallel rank derives the same clip coefficient from the same gradients.
 * gh_adamw_step: torch.optim.AdamW semantics (train_SigLIP_stage1.py:147-153,273) with the clip
 *   coefficient min(1, max_norm / (grad_scale * sqrt(*gnorm_sq) + 1e-6)) applied on the fly
 *   (gnorm_sq NULL or max_norm <= 0: no clipping).  param/grad/exp_avg/exp_avg_sq share `dtype`
 *   (bf16 DiT with bf16 states, fp32 projectors with fp32 states -- no fp32 master copy, as the
 *   reference); fp32 arithmetic, one rounding on store.  step >= 1 is the bias-correction count.
 *   dev_state (device int32[2], may be NULL): when given, the bias-correction count is read from
 *   dev_state[0] on the device instead of `step`, and the launch is a no-op while dev_state[1] == 0.
 *   That makes the update capturable into the CUDA graph of the training step (graph.PipelinedTrainStep:
 *   the update of step n runs on a side branch under step n+1's frozen AE / tower forward).
 * -------------------------------------------------------------------------- */
int64_t gh_sumsq_workspace_bytes(void);
int gh_sumsq_accum(const void* g, int32_t dtype, int64_t numel, float* acc, float* ws, void* stream);
int gh_adamw_step(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int32_t dtype, int64_t numel,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                  const float* gnorm_sq, float max_norm, float grad_scale, const int32_t* dev_state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENHANCER_B200_H_ */
