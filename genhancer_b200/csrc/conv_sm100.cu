// genhancer_b200 -- implicit-GEMM convolution on tcgen05 (NHWC activations), host launcher.
//
// The device code is umma_gemm_kernel<.., MODE_CONV>: every K-step (filter tap, 64-channel chunk) pulls a
// shifted [TH x TW x 64ch] box of the NHWC input with ONE 4-D TMA load; out-of-image taps are zero-filled by
// the TMA unit, so padding (including the FLUX Downsample's asymmetric right/bottom pad,
// src/flux/modules/autoencoder.py:85-95) costs nothing, and stride-2 convs use the tensor map's element strides.
// Replaces nn.Conv2d at autoencoder.py:62-67 (ResnetBlock 3x3), :89 (Downsample), :157 (conv_out).
#include <cstdlib>

#include "conv3x3_res.cuh"
#include "internal.h"
#include "umma_gemm.cuh"

namespace gh {

template <int BN>
static int conv_set_attr() {
  auto* k = umma_gemm_kernel<BN, false, false, MODE_CONV>;
  GH_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::SMEM_BYTES));
  if constexpr (BN >= 128) {
    auto* k2 = umma_gemm_kernel<BN, false, false, MODE_CONV, true>;
    GH_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN, true>::SMEM_BYTES));
  }
  return GH_OK;
}

int conv_init() {
  GH_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Conv3x3ResCfg::SMEM_BYTES));
  if (int e = conv_set_attr<256>()) return e;
  if (int e = conv_set_attr<128>()) return e;
  if (int e = conv_set_attr<64>()) return e;
  return GH_OK;
}

template <int BN>
static int conv_launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  umma_gemm_kernel<BN, false, false, MODE_CONV><<<grid, GemmCfg<BN>::THREADS, GemmCfg<BN>::SMEM_BYTES, s>>>(ta, tb, ta, tb, ta, tb, p);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

// CTA pairs: two adjacent output patches (256 GEMM rows) share one weight tile, half of it staged per CTA
template <int BN>
static int conv_launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  const int tiles = ((p.num_m_blocks + 1) / 2) * p.num_n_blocks;
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  cfg.blockDim = dim3(GemmCfg<BN, true>::THREADS);
  cfg.dynamicSmemBytes = GemmCfg<BN, true>::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  GH_CHECK_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<BN, false, false, MODE_CONV, true>, ta, tb, ta, tb, ta, tb, p));
  return GH_OK;
}

// 3x3 / stride 1 / pad 1, Cin = Cout = 128: weights resident in shared memory, one input box per column shift
// (conv3x3_res.cuh).  Returns GH_OK after launching, or 1 when the shape is not this kernel's.
static int conv3x3_res_try(const gh_conv_args* a, cudaStream_t s) {
  using Cfg = Conv3x3ResCfg;
  static const int enabled = [] { const char* e = getenv("GH_CONV_RES"); return e ? atoi(e) : 1; }();
  if (!enabled || a->Cin != 128 || a->Cout != 128 || a->KH != 3 || a->KW != 3 || a->stride != 1 || a->pad != 1 ||
      a->Ho != a->H || a->Wo != a->W || a->act != 0 || a->y_dtype != GH_BF16 || (a->residual && a->res_dtype != GH_BF16))
    return 1;
  GemmParams p{};
  static const int prefetch = [] { const char* e = getenv("GH_CONV_PREFETCH"); return e ? atoi(e) : 1; }();
  p.dbg = prefetch ? 0 : 2;
  p.cv.B = a->B; p.cv.Ho = a->Ho; p.cv.Wo = a->Wo; p.cv.TW = Cfg::TW; p.cv.TH = Cfg::TH;
  p.cv.tiles_w = (a->Wo + Cfg::TW - 1) / Cfg::TW;
  p.cv.tiles_h = (a->Ho + Cfg::TH - 1) / Cfg::TH;
  p.num_m_blocks = a->B * p.cv.tiles_w * p.cv.tiles_h;
  if (p.num_m_blocks < 2 * num_sms()) return 1;           // too few patches to amortise the 144 KB weight load per CTA
  p.M = a->B * a->Ho * a->Wo; p.N = 128; p.K = 9 * 128;
  p.ep.d = a->y; p.ep.ldd = a->Cout; p.ep.d_f32 = 0;
  p.ep.alpha = 1.f;
  p.ep.bias = a->bias; p.ep.bias_f32 = (a->bias_dtype == GH_F32);
  p.ep.rows_per_batch = 1;
  p.ep.residual = a->residual; p.ep.ld_res = a->Cout; p.ep.res_f32 = 0;
  p.ep.vec8 = 1;
  finalize_epilogue(p.ep);
  CUtensorMap tx, tw;
  {
    const uint64_t dims[4] = {128, static_cast<uint64_t>(a->W), static_cast<uint64_t>(a->H), static_cast<uint64_t>(a->B)};
    const uint64_t strides[3] = {128 * 2, static_cast<uint64_t>(a->W) * 128 * 2, static_cast<uint64_t>(a->H) * a->W * 128 * 2};
    const uint32_t box[4] = {64, Cfg::TW, Cfg::TH + 2, 1};
    if (int e = make_tmap_bf16(&tx, a->x, 4, dims, strides, box, nullptr)) return e;
  }
  {
    const uint64_t dims[2] = {9 * 128, 128};
    const uint64_t strides[1] = {9 * 128 * 2};
    const uint32_t box[2] = {64, 64};
    if (int e = make_tmap_bf16(&tw, a->w, 2, dims, strides, box, nullptr)) return e;
  }
  const int tiles = (p.num_m_blocks + 1) / 2;
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  GH_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_res_kernel, tx, tw, p));
  return GH_OK;
}

// output patch (TW x TH <= 128 pixels) that wastes the fewest MMA rows
static void pick_patch(int Wo, int Ho, int* tw, int* th) {
  double best = -1;
  for (int w = 1; w <= 128 && w <= Wo; ++w) {
    int h = 128 / w;
    if (h > Ho) h = Ho;
    if (h < 1) continue;
    const long tiles = static_cast<long>((Wo + w - 1) / w) * ((Ho + h - 1) / h);
    const double eff = static_cast<double>(Wo) * Ho / (tiles * 128.0);
    if (eff > best + 1e-9) { best = eff; *tw = w; *th = h; }
  }
}

}  // namespace gh

using namespace gh;

extern "C" int gh_conv2d_nhwc(const gh_conv_args* a, void* stream) {
  GH_REQUIRE(a && a->x && a->w && a->y, GH_ERR_NULL, "gh_conv2d_nhwc: NULL pointer");
  GH_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, GH_ERR_BAD_SHAPE, "gh_conv2d_nhwc: bad shape");
  GH_REQUIRE(a->Cin % 64 == 0, GH_ERR_UNSUPPORTED, "gh_conv2d_nhwc: Cin=%d must be a multiple of 64", a->Cin);
  GH_REQUIRE(a->Cout % 8 == 0, GH_ERR_UNSUPPORTED, "gh_conv2d_nhwc: Cout=%d must be a multiple of 8", a->Cout);
  GH_REQUIRE(a->KH >= 1 && a->KW >= 1 && a->KH <= 7 && a->KW <= 7 && (a->stride == 1 || a->stride == 2) && a->pad >= 0,
             GH_ERR_UNSUPPORTED, "gh_conv2d_nhwc: unsupported filter geometry");
  GH_REQUIRE(a->Ho > 0 && a->Wo > 0, GH_ERR_BAD_SHAPE, "gh_conv2d_nhwc: bad output extent");
  GH_REQUIRE(aligned16(a->x) && aligned16(a->w) && aligned16(a->y), GH_ERR_ALIGN, "gh_conv2d_nhwc: 16B alignment");
  GH_REQUIRE(a->act >= 0 && a->act <= 4, GH_ERR_UNSUPPORTED, "gh_conv2d_nhwc: unknown act");

  if (aligned16(a->bias) && aligned16(a->residual)) {
    const int r = conv3x3_res_try(a, static_cast<cudaStream_t>(stream));
    if (r <= 0) return r;            // launched (GH_OK) or failed; 1 = not this kernel's shape
  }
  int TW = 1, TH = 1;
  pick_patch(a->Wo, a->Ho, &TW, &TH);
  const int K = a->KH * a->KW * a->Cin;
  int bn = a->Cout >= 256 ? 256 : (a->Cout > 64 ? 128 : 64);
  GemmParams p{};
  p.k_splits = 1;
  p.prof = gemm_prof_ptr();
  p.cv.B = a->B; p.cv.Ho = a->Ho; p.cv.Wo = a->Wo; p.cv.TW = TW; p.cv.TH = TH;
  p.cv.tiles_w = (a->Wo + TW - 1) / TW;
  p.cv.tiles_h = (a->Ho + TH - 1) / TH;
  p.cv.KW = a->KW; p.cv.cin_chunks = a->Cin / 64; p.cv.stride = a->stride; p.cv.pad = a->pad;
  p.M = a->B * a->Ho * a->Wo; p.N = a->Cout; p.K = K;
  p.num_m_blocks = a->B * p.cv.tiles_w * p.cv.tiles_h;
  // fewer, fatter tiles waste the tail wave: drop to a narrower N tile if that fills the machine better
  while (bn > 64 && static_cast<long>(p.num_m_blocks) * ((a->Cout + bn - 1) / bn) < num_sms()) bn /= 2;
  p.num_n_blocks = (a->Cout + bn - 1) / bn;
  static const int conv_pair = [] { const char* e = getenv("GH_CONV_PAIR"); return e ? atoi(e) : 1; }();
  const bool pair = conv_pair && bn >= 128 && p.num_m_blocks >= 2 * num_sms();
  p.num_k_blocks = K / 64;
  p.a_stage_tx_bytes = static_cast<uint32_t>(TW * TH * 128);
  p.mn_lbo = 8192; p.mn_sbo = 1024; p.mn_kstep = 2048;
  p.ep.d = a->y; p.ep.ldd = a->Cout; p.ep.d_f32 = (a->y_dtype == GH_F32);
  p.ep.alpha = 1.f;
  p.ep.bias = a->bias; p.ep.bias_f32 = (a->bias_dtype == GH_F32);
  p.ep.act = a->act;
  p.ep.rows_per_batch = 1;
  p.ep.residual = a->residual; p.ep.ld_res = a->Cout; p.ep.res_f32 = (a->res_dtype == GH_F32);
  p.ep.vec8 = (!a->bias || a->bias_dtype == GH_F32 || aligned16(a->bias)) && (!a->residual || aligned16(a->residual));
  finalize_epilogue(p.ep);

  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(a->Cin), static_cast<uint64_t>(a->W), static_cast<uint64_t>(a->H),
                              static_cast<uint64_t>(a->B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(a->Cin) * 2, static_cast<uint64_t>(a->W) * a->Cin * 2,
                                 static_cast<uint64_t>(a->H) * a->W * a->Cin * 2};
    const uint32_t box[4] = {64, static_cast<uint32_t>(TW * a->stride), static_cast<uint32_t>(TH * a->stride), 1};
    const uint32_t est[4] = {1, static_cast<uint32_t>(a->stride), static_cast<uint32_t>(a->stride), 1};
    if (int e = make_tmap_bf16(&ta, a->x, 4, dims, strides, box, est)) return e;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(a->Cout)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(pair ? bn / 2 : bn)};
    if (int e = make_tmap_bf16(&tb, a->w, 2, dims, strides, box, nullptr)) return e;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pair) return bn == 256 ? conv_launch_pair<256>(ta, tb, p, s) : conv_launch_pair<128>(ta, tb, p, s);
  switch (bn) {
    case 256: return conv_launch<256>(ta, tb, p, s);
    case 128: return conv_launch<128>(ta, tb, p, s);
    default: return conv_launch<64>(ta, tb, p, s);
  }
}
