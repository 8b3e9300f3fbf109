// genhancer_b200 -- host-side internals shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/genhancer_b200.h"

namespace gh {

// thread-local error string behind gh_last_error()
int set_error(int code, const char* fmt, ...);

#define GH_CHECK_CUDA(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::gh::set_error(GH_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                \
  } while (0)

#define GH_REQUIRE(cond, code, ...)                        \
  do {                                                     \
    if (!(cond)) return ::gh::set_error(code, __VA_ARGS__); \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// bf16 tiled tensor map, SWIZZLE_128B, zero OOB fill. dims/box innermost-first (elements),
// strides_bytes has rank-1 entries (dims 1..rank-1). elem_strides may be NULL (all 1).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

int num_sms();
long long* gemm_prof_ptr();  // gh_debug_gemm_prof buffer or NULL

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// per-file one-time setup hooks called from gh_init()
int gemm_init();
int attn_init();
int conv_init();
int patch_embed_init();

}  // namespace gh
