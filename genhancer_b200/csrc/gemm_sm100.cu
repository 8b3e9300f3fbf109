// genhancer_b200 -- gh_gemm_bf16: host launcher of the tcgen05 GEMM (see umma_gemm.cuh).
#include <atomic>
#include <cstdlib>

#include "internal.h"
#include "umma_gemm.cuh"

namespace gh {

// dynamic tile schedule (gh_gemm_args::dynamic_tiles): a pool of {next tile, workers done} counter pairs, one per launch in
// rotation, so launches on different streams never share a pair; a kernel hands its pair back zeroed
static constexpr unsigned TILE_CTR_SLOTS = 1024;
static int* g_tile_ctrs = nullptr;
static int g_tile_dynamic = 0;   // GH_TILE_SCHEDULER=1 in the environment: every launch dynamic (A/B runs)
static std::atomic<unsigned> g_tile_ctr_next{0};

template <int BN, bool A_MN, bool B_MN>
static int set_attr() {
  auto* k = umma_gemm_kernel<BN, A_MN, B_MN, MODE_GEMM>;
  GH_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::SMEM_BYTES));
  if constexpr (BN >= 128) {
    auto* k2 = umma_gemm_kernel<BN, A_MN, B_MN, MODE_GEMM, true>;
    GH_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN, true>::SMEM_BYTES));
  }
  return GH_OK;
}

int gemm_init() {
  if (g_tile_ctrs == nullptr) {
    GH_CHECK_CUDA(cudaMalloc(&g_tile_ctrs, TILE_CTR_SLOTS * 2 * sizeof(int)));
    GH_CHECK_CUDA(cudaMemset(g_tile_ctrs, 0, TILE_CTR_SLOTS * 2 * sizeof(int)));
  }
  if (const char* e = getenv("GH_TILE_SCHEDULER")) g_tile_dynamic = atoi(e) != 0;
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
  const long tiles = static_cast<long>(p.num_m_blocks) * p.num_n_blocks * p.k_splits * p.batch;
  const int grid = tiles < num_sms() ? static_cast<int>(tiles) : num_sms();
  umma_gemm_kernel<BN, A_MN, B_MN, MODE_GEMM>
      <<<grid, GemmCfg<BN>::THREADS, GemmCfg<BN>::SMEM_BYTES, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

// CTA pairs: clusters of 2 CTAs (one TPC), one pair per 256 x BN tile, persistent over the pair-tiles
template <int BN, bool A_MN, bool B_MN>
static int launch_pair(const CUtensorMap* tm, const GemmParams& p, cudaStream_t stream) {
  const long tiles = static_cast<long>((p.num_m_blocks + 1) / 2) * p.num_n_blocks * p.k_splits * p.batch;
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (tiles < pairs ? static_cast<int>(tiles) : pairs));
  cfg.blockDim = dim3(GemmCfg<BN, true>::THREADS);
  cfg.dynamicSmemBytes = GemmCfg<BN, true>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  GH_CHECK_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<BN, A_MN, B_MN, MODE_GEMM, true>, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p));
  return GH_OK;
}

template <int BN>
static int dispatch_major(bool pair, bool a_mn, bool b_mn, const CUtensorMap* tm, const GemmParams& p, cudaStream_t s) {
  if constexpr (BN >= 128) {
    if (pair) {
      if (!a_mn && !b_mn) return launch_pair<BN, false, false>(tm, p, s);
      if (!a_mn && b_mn) return launch_pair<BN, false, true>(tm, p, s);
      if (a_mn && !b_mn) return launch_pair<BN, true, false>(tm, p, s);
      return launch_pair<BN, true, true>(tm, p, s);
    }
  }
  if (!a_mn && !b_mn) return launch<BN, false, false>(tm, p, s);
  if (!a_mn && b_mn) return launch<BN, false, true>(tm, p, s);
  if (a_mn && !b_mn) return launch<BN, true, false>(tm, p, s);
  return launch<BN, true, true>(tm, p, s);
}

// tensor maps of one (A, B) operand pair with reduction length K
// (batched: a_extra / b_extra = (batch - 1) * batch offset, the rows the flat tensor holds beyond one problem)
static int make_pair(CUtensorMap* ta, CUtensorMap* tb, const void* a, int64_t lda, bool a_mn, const void* b, int64_t ldb,
                     bool b_mn, int M, int N, int K, int bn /* B rows per TMA box */, int64_t a_extra = 0,
                     int64_t b_extra = 0) {
  {
    // A: K-major -> dims (K, M) box (64, 128); MN-major -> dims (M, K) box (64, 64)
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2];
    if (!a_mn) { dims[0] = K; dims[1] = M + a_extra; box[0] = 64; box[1] = 128; }
    else       { dims[0] = M; dims[1] = K + a_extra; box[0] = 64; box[1] = 64; }
    if (int e = make_tmap_bf16(ta, a, 2, dims, strides, box, nullptr)) return e;
  }
  {
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2];
    if (!b_mn) { dims[0] = K; dims[1] = N + b_extra; box[0] = 64; box[1] = static_cast<uint32_t>(bn); }
    else       { dims[0] = N; dims[1] = K + b_extra; box[0] = 64; box[1] = 64; }
    if (int e = make_tmap_bf16(tb, b, 2, dims, strides, box, nullptr)) return e;
  }
  return GH_OK;
}

static long long* g_gemm_prof = nullptr;  // gh_debug_gemm_prof
long long* gemm_prof_ptr() { return g_gemm_prof; }

// Tile shape and CTA mode.  Cost model in MMA-k-step cycles per tile (measured with gh_debug_gemm_prof: the issuer
// is busy ~92 % of the time, yet a single-CTA 128 x 256 tile needs ~200 cycles per 16-deep MMA instead of 128):
//   tensor pipe      BN / 2 cycles per k-step (per SM)
//   shared memory    operand reads + TMA writes at 128 B/clk: single CTA (128 + BN) / 2, CTA pair (128 + BN/2) / 2
// plus a fixed per-tile cost; waves = tiles per worker (148 CTAs, or 74 pairs each covering 256 rows).
struct TileChoice { int bn; bool pair; };
static TileChoice pick_tile(int M, int N, int batch = 1) {
  const long mb = (M + 127) / 128;
  const int sms = num_sms();
  static const int force = [] { const char* e = getenv("GH_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  TileChoice best{64, false};
  double best_cost = 1e30;
  const int bns[3] = {256, 128, 64};
  for (int pair = 0; pair < 2; ++pair) {
    if (pair ? (mb < 2 || force == 0) : (force == 1 && mb >= 2)) continue;
    for (int bn : bns) {
      if (pair && bn < 128) continue;
      if (bn > 64 && N <= bn / 2) continue;  // a tile more than twice as wide as the problem
      const long nb = (N + bn - 1) / bn;
      const long tiles = (pair ? (mb + 1) / 2 : mb) * nb * batch;
      const long workers = pair ? sms / 2 : sms;
      const long waves = (tiles + workers - 1) / workers;
      const double t = pair ? (bn / 2 > (128 + bn / 2) / 2 ? bn / 2 : (128 + bn / 2) / 2)
                            : (bn / 2 > (128 + bn) / 2 ? bn / 2 : (128 + bn) / 2);
      const double cost = waves * (t + 24.0);
      if (cost < best_cost - 1e-9) { best_cost = cost; best = TileChoice{bn, pair != 0}; }
    }
  }
  return best;
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

  const int batch = a->batch > 1 ? a->batch : 1;
  GH_REQUIRE(batch == 1 || (a->K2 == 0 && a->k_splits == 0 && !a->gate && a->a_batch_rows >= 0 && a->b_batch_rows >= 0 &&
                            a->d_batch_rows >= a->M),
             GH_ERR_UNSUPPORTED, "gh_gemm_bf16: batched mode takes no second operand pair, split-K or gate, and needs "
                                 "d_batch_rows >= M");
  TileChoice tc = pick_tile(a->M, a->N, batch);
  if (a->k_splits != 0) {
    // split-K = a skinny output reduced over every token (LoRA wgrads): the kernel streams one big operand once and is
    // HBM-bound, so the tile that counts is the one with the WIDEST boxes of that operand (longer DRAM bursts, fewer MMA
    // instructions per byte -- an N = 64 MMA costs what an N = 128 one costs): 256-row pair tiles when the big operand is A
    // (dB = dY^T u: M = out features), 256-column tiles when it is B (dA = du^T x: N = in features).
    if (a->M >= 256 && a->N <= 128) tc = TileChoice{128, true};
    else if (a->M <= 128 && a->N >= 256) tc = TileChoice{256, false};
  }
  const int bn = tc.bn;
  GemmParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_m_blocks = (a->M + 127) / 128;
  p.num_n_blocks = (a->N + bn - 1) / bn;
  p.num_k_blocks = (a->K + 63) / 64;
  p.tile_ctr = ((a->dynamic_tiles || g_tile_dynamic) && g_tile_ctrs) ? g_tile_ctrs + 2 * (g_tile_ctr_next.fetch_add(1) % TILE_CTR_SLOTS) : nullptr;
  p.batch = batch;
  p.a_boff = static_cast<int>(a->a_batch_rows); p.b_boff = static_cast<int>(a->b_batch_rows);
  p.d_brows = static_cast<int>(a->d_batch_rows);
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
  // split-K: skinny fp32 outputs reduced over a long K (LoRA wgrads)
  p.k_splits = 1;
  if (a->k_splits != 0) {
    GH_REQUIRE(a->d_dtype == GH_F32 && a->K2 == 0 && !a->bias && a->act == 0 && !a->act_grad && !a->gate && !a->residual &&
                   !a->aux_out,
               GH_ERR_UNSUPPORTED, "gh_gemm_bf16: split-K adds bare partial products into an fp32 D (no epilogue options)");
    const int tiles_mn = (tc.pair ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks) * p.num_n_blocks;
    const int workers = tc.pair ? num_sms() / 2 : num_sms();
    int s = a->k_splits > 0 ? a->k_splits : workers / (tiles_mn > 0 ? tiles_mn : 1);   // < 0: fill the machine
    if (s > p.num_k_blocks / 4) s = p.num_k_blocks / 4;                                 // >= 4 k blocks per slice
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    p.k_splits = s;
    p.ep.atomic = 1;   // also for s == 1: the contract is "D += A B^T"
  }
  finalize_epilogue(p.ep);
  p.prof = g_gemm_prof;
  if (g_gemm_prof) { static const int dbg = [] { const char* e = getenv("GH_GEMM_DBG"); return e ? atoi(e) : 0; }(); p.dbg = dbg; }

  const bool amn = a->a_mn_major != 0, bmn = a->b_mn_major != 0;
  CUtensorMap tm[6];
  const int box_n = tc.pair ? bn / 2 : bn;
  if (int e = make_pair(&tm[0], &tm[1], a->a, a->lda, amn, a->b, a->ldb, bmn, a->M, a->N, a->K, box_n,
                        (batch - 1) * a->a_batch_rows, (batch - 1) * a->b_batch_rows))
    return e;
  if (a->K2 > 0) {
    if (int e = make_pair(&tm[2], &tm[3], a->a2, a->lda2, amn, a->b2, a->ldb2, bmn, a->M, a->N, a->K2, box_n)) return e;
    p.num_k_blocks2 = (a->K2 + 63) / 64;
    const int tail = a->K2 - (p.num_k_blocks2 - 1) * 64;
    p.k2_last_steps = (tail + 15) / 16;
  } else {
    tm[2] = tm[0];
    tm[3] = tm[1];
  }
  // lean epilogue through [32 rows x 64 columns] SWIZZLE_128B tiles: TMA store of D, TMA load of the residual
  static const int no_tma_epi = [] { const char* e = getenv("GH_GEMM_NO_TMA_EPI"); return e ? atoi(e) : 0; }();
  // act_grad (dY of an MLP's fc1 = (dH W2) * act'(pre)) rides the same path: the saved pre-activation tile arrives by
  // TMA where the residual tile would, for the two activations the path trains through (GELU-tanh, QuickGELU)
  const bool lean_actgrad = a->act_grad && (a->act == GH_ACT_GELU_TANH || a->act == GH_ACT_QUICK_GELU) && p.ep.vec8 &&
                            a->d_dtype == GH_BF16 && !a->aux_out && !a->gate && !a->residual && a->k_splits == 0;
  // aux_out (the forward of an MLP's fc1 saves the pre-activation beside the activated output): a second smem tile and
  // a second bulk store per group.  (Through the general epilogue the SigLIP tower's fc1, K = 1152, ran at 700 TFLOP/s
  // where its bias-only neighbour with the same K reaches 1450: the epilogue, not the mainloop, paced the kernel.)
  const bool lean_auxout = a->aux_out && a->act != GH_ACT_NONE && !a->act_grad && p.ep.vec8 && a->d_dtype == GH_BF16 && !a->gate &&
                           !a->residual && a->k_splits == 0;
  p.ep.tma = (p.ep.fast || lean_actgrad || lean_auxout) && !no_tma_epi && batch == 1 &&
             (!a->bias || (reinterpret_cast<uintptr_t>(a->bias) & 15u) == 0);
  tm[4] = tm[0];
  tm[5] = tm[0];
  if (p.ep.tma) {
    uint64_t dims[2] = {static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->M)};
    uint32_t box[2] = {64, 32};
    uint64_t sd[1] = {static_cast<uint64_t>(a->ldd) * 2};
    if (int e = make_tmap_bf16(&tm[4], a->d, 2, dims, sd, box, nullptr)) return e;
    if (a->act_grad) {
      uint64_t sr[1] = {static_cast<uint64_t>(a->ld_aux_in) * 2};
      if (int e = make_tmap_bf16(&tm[5], a->aux_in, 2, dims, sr, box, nullptr)) return e;
    } else if (a->residual) {
      uint64_t sr[1] = {static_cast<uint64_t>(a->ld_res) * 2};
      if (int e = make_tmap_bf16(&tm[5], a->residual, 2, dims, sr, box, nullptr)) return e;
    } else if (a->aux_out) {
      uint64_t sr[1] = {static_cast<uint64_t>(a->ld_aux_out) * 2};
      if (int e = make_tmap_bf16(&tm[5], a->aux_out, 2, dims, sr, box, nullptr)) return e;
    }
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return dispatch_major<256>(tc.pair, amn, bmn, tm, p, s);
    case 128: return dispatch_major<128>(tc.pair, amn, bmn, tm, p, s);
    default: return dispatch_major<64>(false, amn, bmn, tm, p, s);
  }
}

extern "C" int gh_debug_gemm_prof(void* device_buf) {
  gh::g_gemm_prof = static_cast<long long*>(device_buf);
  return GH_OK;
}
