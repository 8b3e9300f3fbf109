// genhancer_b200 -- gh_gemm_bf16: host launcher of the tcgen05 GEMM (see umma_gemm.cuh).
#include <cstdlib>

#include "internal.h"
#include "umma_gemm.cuh"

namespace gh {

template <int BN, bool A_MN, bool B_MN>
static int set_attr() {
  auto* k = umma_gemm_kernel<BN, A_MN, B_MN, MODE_GEMM>;
  GH_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::SMEM_BYTES));
  return GH_OK;
}

int gemm_init() {
#define GH_SET(BN)                                   \
  if (int e = set_attr<BN, false, false>()) return e; \
  if (int e = set_attr<BN, false, true>()) return e;  \
  if (int e = set_attr<BN, true, false>()) return e;  \
  if (int e = set_attr<BN, true, true>()) return e;
  GH_SET(256)
  GH_SET(128)
  GH_SET(64)
#undef GH_SET
  return GH_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const CUtensorMap* tm, const GemmParams& p, cudaStream_t stream) {
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  umma_gemm_kernel<BN, A_MN, B_MN, MODE_GEMM>
      <<<grid, GemmCfg<BN>::THREADS, GemmCfg<BN>::SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], tm[3], p);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

template <int BN>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap* tm, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(tm, p, s);
  if (!a_mn && b_mn) return launch<BN, false, true>(tm, p, s);
  if (a_mn && !b_mn) return launch<BN, true, false>(tm, p, s);
  return launch<BN, true, true>(tm, p, s);
}

// tensor maps of one (A, B) operand pair with reduction length K
static int make_pair(CUtensorMap* ta, CUtensorMap* tb, const void* a, int64_t lda, bool a_mn, const void* b, int64_t ldb,
                     bool b_mn, int M, int N, int K, int bn) {
  {
    // A: K-major -> dims (K, M) box (64, 128); MN-major -> dims (M, K) box (64, 64)
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2];
    if (!a_mn) { dims[0] = K; dims[1] = M; box[0] = 64; box[1] = 128; }
    else       { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = 64; }
    if (int e = make_tmap_bf16(ta, a, 2, dims, strides, box, nullptr)) return e;
  }
  {
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2];
    if (!b_mn) { dims[0] = K; dims[1] = N; box[0] = 64; box[1] = static_cast<uint32_t>(bn); }
    else       { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = 64; }
    if (int e = make_tmap_bf16(tb, b, 2, dims, strides, box, nullptr)) return e;
  }
  return GH_OK;
}

static int pick_bn(int M, int N) {
  // Per-tile time ~ BN + c (MMA issue is proportional to BN; c = fixed per-tile cost), and narrow tiles re-read
  // the A operand from L2 once per BN columns: 128x64 tiles run at about a third of the 128x256 rate
  // (measured: wgrad M=21504 N=3072 K=14144 at 455 TFLOP/s with BN=64 vs 1350 with BN=256).  So: the widest tile
  // the problem fills, dropping one notch only when that removes at least a quarter of the waves' work.
  const long mb = (M + 127) / 128;
  const int sms = num_sms();
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  auto cost = [&](int bn) {
    const long tiles = mb * ((N + bn - 1) / bn);
    const long waves = (tiles + sms - 1) / sms;
    return waves * (bn + 32);
  };
  const long c256 = cost(256), c128 = cost(128);
  if (c128 * 4 <= c256 * 3) return 128;
  return 256;
}

}  // namespace gh

extern "C" int gh_gemm_bf16(const gh_gemm_args* a, void* stream) {
  using namespace gh;
  GH_REQUIRE(a != nullptr, GH_ERR_NULL, "gh_gemm_bf16: args is NULL");
  GH_REQUIRE(a->a && a->b && a->d, GH_ERR_NULL, "gh_gemm_bf16: a/b/d must be non-NULL");
  GH_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, GH_ERR_BAD_SHAPE, "gh_gemm_bf16: M,N,K must be positive (%d,%d,%d)",
             a->M, a->N, a->K);
  GH_REQUIRE(a->N % 4 == 0, GH_ERR_BAD_SHAPE, "gh_gemm_bf16: N=%d must be a multiple of 4", a->N);
  GH_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0 && a->ldd % 4 == 0, GH_ERR_ALIGN,
             "gh_gemm_bf16: lda/ldb must be multiples of 8, ldd of 4 (%lld,%lld,%lld)", (long long)a->lda,
             (long long)a->ldb, (long long)a->ldd);
  GH_REQUIRE(aligned16(a->a) && aligned16(a->b) && aligned16(a->d), GH_ERR_ALIGN,
             "gh_gemm_bf16: a/b/d must be 16-byte aligned");
  GH_REQUIRE(a->lda >= (a->a_mn_major ? a->M : a->K) && a->ldb >= (a->b_mn_major ? a->N : a->K) && a->ldd >= a->N,
             GH_ERR_BAD_SHAPE, "gh_gemm_bf16: leading dimension smaller than the row length");
  GH_REQUIRE(a->K2 >= 0 && (a->K2 == 0 || (a->a2 && a->b2)), GH_ERR_NULL, "gh_gemm_bf16: K2 > 0 needs a2 and b2");
  GH_REQUIRE(a->K2 == 0 || (a->lda2 % 8 == 0 && a->ldb2 % 8 == 0 && aligned16(a->a2) && aligned16(a->b2) &&
                            a->lda2 >= (a->a_mn_major ? a->M : a->K2) && a->ldb2 >= (a->b_mn_major ? a->N : a->K2)),
             GH_ERR_ALIGN, "gh_gemm_bf16: a2/b2 need 16-byte alignment, lda2/ldb2 multiples of 8 and >= the row length");
  GH_REQUIRE(a->act >= 0 && a->act <= 4, GH_ERR_UNSUPPORTED, "gh_gemm_bf16: unknown act %d", a->act);
  GH_REQUIRE(!a->act_grad || a->aux_in, GH_ERR_NULL, "gh_gemm_bf16: act_grad needs aux_in");
  GH_REQUIRE(!a->gate || a->rows_per_batch > 0, GH_ERR_BAD_SHAPE, "gh_gemm_bf16: gate needs rows_per_batch > 0");
  GH_REQUIRE((a->d_dtype == GH_BF16 || a->d_dtype == GH_F32), GH_ERR_UNSUPPORTED, "gh_gemm_bf16: bad d_dtype");
  GH_REQUIRE((!a->aux_in || a->ld_aux_in % 4 == 0) && (!a->aux_out || a->ld_aux_out % 4 == 0) &&
                 (!a->gate || a->gate_ld % 4 == 0) && (!a->residual || a->ld_res % 4 == 0),
             GH_ERR_ALIGN, "gh_gemm_bf16: epilogue operand leading dimensions must be multiples of 4");

  const int bn = pick_bn(a->M, a->N);
  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_m_blocks = (a->M + 127) / 128;
  p.num_n_blocks = (a->N + bn - 1) / bn;
  p.num_k_blocks = (a->K + 63) / 64;
  p.a_stage_tx_bytes = 128 * 64 * 2;
  p.mn_lbo = 8192; p.mn_sbo = 1024; p.mn_kstep = 2048;
  if (const char* dbg = getenv("GH_DEBUG_MN_DESC")) {  // bring-up aid: "lbo,sbo,kstep" in bytes
    unsigned l, sb, ks;
    if (sscanf(dbg, "%u,%u,%u", &l, &sb, &ks) == 3) { p.mn_lbo = l; p.mn_sbo = sb; p.mn_kstep = ks; }
  }
  p.ep.d = a->d; p.ep.ldd = a->ldd; p.ep.d_f32 = (a->d_dtype == GH_F32);
  p.ep.alpha = a->alpha;
  p.ep.bias = a->bias; p.ep.bias_f32 = (a->bias_dtype == GH_F32);
  p.ep.act = a->act; p.ep.act_grad = a->act_grad;
  p.ep.aux_in = static_cast<const __nv_bfloat16*>(a->aux_in); p.ep.ld_aux_in = a->ld_aux_in;
  p.ep.aux_out = static_cast<__nv_bfloat16*>(a->aux_out); p.ep.ld_aux_out = a->ld_aux_out;
  p.ep.gate = static_cast<const __nv_bfloat16*>(a->gate); p.ep.gate_ld = a->gate_ld;
  p.ep.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : 1;
  p.ep.residual = a->residual; p.ep.ld_res = a->ld_res; p.ep.res_f32 = (a->res_dtype == GH_F32);
  p.ep.vec8 = (a->N % 8 == 0) && (a->d_dtype == GH_F32 || (a->ldd % 8 == 0)) &&
              (!a->bias || a->bias_dtype == GH_F32 || aligned16(a->bias)) &&
              (!a->aux_in || (a->ld_aux_in % 8 == 0 && aligned16(a->aux_in))) &&
              (!a->aux_out || (a->ld_aux_out % 8 == 0 && aligned16(a->aux_out))) &&
              (!a->gate || (a->gate_ld % 8 == 0 && aligned16(a->gate))) &&
              (!a->residual || a->res_dtype == GH_F32 || (a->ld_res % 8 == 0 && aligned16(a->residual)));
  finalize_epilogue(p.ep);

  const bool amn = a->a_mn_major != 0, bmn = a->b_mn_major != 0;
  CUtensorMap tm[4];
  if (int e = make_pair(&tm[0], &tm[1], a->a, a->lda, amn, a->b, a->ldb, bmn, a->M, a->N, a->K, bn)) return e;
  if (a->K2 > 0) {
    if (int e = make_pair(&tm[2], &tm[3], a->a2, a->lda2, amn, a->b2, a->ldb2, bmn, a->M, a->N, a->K2, bn)) return e;
    p.num_k_blocks2 = (a->K2 + 63) / 64;
    const int tail = a->K2 - (p.num_k_blocks2 - 1) * 64;
    p.k2_last_steps = (tail + 15) / 16;
  } else {
    tm[2] = tm[0];
    tm[3] = tm[1];
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return dispatch_major<256>(amn, bmn, tm, p, s);
    case 128: return dispatch_major<128>(amn, bmn, tm, p, s);
    default: return dispatch_major<64>(amn, bmn, tm, p, s);
  }
}
