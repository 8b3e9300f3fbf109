// genhancer_b200 -- C-ABI plumbing: error reporting, init, TMA descriptor encoding.
#include <cstdlib>
#include <mutex>
#include <string>

#include "internal.h"

namespace gh {

static thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

static PFN_encodeTiled g_encode = nullptr;
static int g_num_sms = 0;
static std::once_flag g_driver_once;

PFN_encodeTiled get_encode_tiled() {
  std::call_once(g_driver_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  });
  return g_encode;
}

static int g_sm_budget = 0;   // GH_SM_BUDGET in the environment (an experiment knob, DESIGN.md section 6: negative result)

int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
    if (const char* e = getenv("GH_SM_BUDGET")) {
      const int b = atoi(e);
      if (b > 0) g_sm_budget = b;
    }
  }
  if (g_sm_budget > 0 && g_sm_budget < g_num_sms) return g_sm_budget & ~1;   // even: CTA pairs need whole TPCs
  return g_num_sms;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled enc = get_encode_tiled();
  GH_REQUIRE(enc != nullptr, GH_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5];
  cuuint32_t gest[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    gest[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                   gdims, gstr, gbox, gest, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GH_REQUIRE(r == CUDA_SUCCESS, GH_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims[0..1]=%llu,%llu stride0=%llu box=%u,%u",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
  return GH_OK;
}

}  // namespace gh

extern "C" const char* gh_last_error(void) { return gh::g_last_error.c_str(); }
extern "C" int gh_version(void) { return 100; }

extern "C" int gh_init(int device) {
  using namespace gh;
  GH_CHECK_CUDA(cudaSetDevice(device));
  int major = 0, minor = 0;
  GH_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  GH_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  GH_REQUIRE(major == 10, GH_ERR_UNSUPPORTED, "genhancer_b200 needs an sm_100a device, found sm_%d%d", major, minor);
  GH_REQUIRE(get_encode_tiled() != nullptr, GH_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  (void)num_sms();
  if (int e = gemm_init()) return e;
  if (int e = attn_init()) return e;
  if (int e = conv_init()) return e;
  if (int e = patch_embed_init()) return e;
  return GH_OK;
}
