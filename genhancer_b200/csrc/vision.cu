// genhancer_b200 -- HBM-bound kernels around the vision tower and the FLUX autoencoder encoder:
// im2col gathers feeding the tcgen05 GEMM, GroupNorm(32)+swish on NHWC, row softmax for the AE's
// single-head mid attention, CLIP token assembly, DiagonalGaussian sampling + 2x2 patchify.
#include "common.cuh"
#include "internal.h"

namespace gh {

using bf16 = __nv_bfloat16;

// ---------------------------------------------------------------------------
// patch-embed im2col (stride == kernel, so it is a pure gather): image fp32 NCHW [B,3,S,S] ->
// bf16 [B*G*G, ld] with k = c*p*p + i*p + j (the flattening of Conv2d weight [D,3,p,p]); optional
// per-channel normalisation (x - mean) / std fused in.  One CTA per (b, patch-row).
// HF CLIPVisionEmbeddings.patch_embedding, modeling_clip.py:147-153,208-209.
// ---------------------------------------------------------------------------
// Pixel (b, c, y, x) of the batch as the fp32 value in [0, 1] the reference's loader hands over: either an fp32 NCHW
// tensor, or -- U8 -- the decoded uint8 HWC image itself, value / 255 with IEEE division (torchvision's ToTensor,
// image_datasets/dataset_cc3m.py:107-113).  With U8 the batch crosses PCIe AND is read from HBM at 1 byte per value
// and the 43 MB fp32 copy of a 32 x 336 x 336 batch never exists.
template <bool U8>
__device__ __forceinline__ float load_px(const void* img, int b, int c, int y, int x, int H, int W) {
  if (U8) {
    const uint8_t* s = static_cast<const uint8_t*>(img);
    return __fdiv_rn(static_cast<float>(s[((static_cast<int64_t>(b) * H + y) * W + x) * 3 + c]), 255.f);
  }
  return static_cast<const float*>(img)[((static_cast<int64_t>(b) * 3 + c) * H + y) * W + x];
}

template <bool U8>
__global__ void patch_im2col_kernel(const void* __restrict__ img, bf16* __restrict__ out, int S, int p, int G,
                                    int64_t ld, float m0, float m1, float m2, float is0, float is1, float is2) {
  const int b = blockIdx.y, py = blockIdx.x;
  const float mean[3] = {m0, m1, m2}, istd[3] = {is0, is1, is2};
  const int row_elems = G * p;  // pixels of one image row that belong to patches
  for (int ci = 0; ci < 3 * p; ++ci) {
    const int c = ci / p, i = ci % p;
    for (int x = threadIdx.x; x < row_elems; x += blockDim.x) {
      const int px = x / p, j = x - px * p;
      const float v = (load_px<U8>(img, b, c, py * p + i, x, S, S) - mean[c]) * istd[c];
      out[(static_cast<int64_t>(b) * G * G + py * G + px) * ld + c * p * p + i * p + j] = __float2bfloat16_rn(v);
    }
  }
}

// 3x3 / pad 1 im2col for the AE's conv_in (Cin = 3): image fp32 NCHW -> bf16 [B*H*W, 32], k = (kh*3+kw)*3 + c,
// columns 27..31 zero; normalisation fused.  autoencoder.py:126.
template <bool U8>
__global__ void im2col3x3_c3_kernel(const void* __restrict__ img, bf16* __restrict__ out, int B, int H, int W,
                                    float mean, float istd) {
  const int64_t n = static_cast<int64_t>(B) * H * W;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = static_cast<int>(i % W);
  const int y = static_cast<int>((i / W) % H);
  const int b = static_cast<int>(i / (static_cast<int64_t>(W) * H));
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int yy = y + kh - 1, xx = x + kw - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[(kh * 3 + kw) * 3 + c] = (load_px<U8>(img, b, c, yy, xx, H, W) - mean) * istd;
      }
    }
  uint4* dst = reinterpret_cast<uint4*>(out + i * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]); u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]); u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
    dst[q] = u;
  }
}

// CLIP / SigLIP token assembly: out[b, t] = (t < has_cls ? cls : patch[b, t - has_cls]) + pos[t]
// (modeling_clip.py:210-217).  8 channels per thread.
__global__ void embed_assemble_kernel(const bf16* __restrict__ patch, const float* __restrict__ cls,
                                      const float* __restrict__ pos, bf16* __restrict__ out, int B, int T, int D,
                                      int has_cls) {
  const int64_t n8 = static_cast<int64_t>(B) * T * D / 8;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int d = static_cast<int>((i * 8) % D);
  const int t = static_cast<int>((i * 8 / D) % T);
  const int b = static_cast<int>(i * 8 / (static_cast<int64_t>(D) * T));
  float v[8];
  if (t < has_cls) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = cls[d + j];
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(patch + (static_cast<int64_t>(b) * (T - has_cls) + t - has_cls) * D + d);
    const float2 a = unpack_bf16x2(u.x), c2 = unpack_bf16x2(u.y), e = unpack_bf16x2(u.z), g = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = c2.x; v[3] = c2.y; v[4] = e.x; v[5] = e.y; v[6] = g.x; v[7] = g.y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] += pos[static_cast<int64_t>(t) * D + d + j];
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(out)[i] = o;
}

// ---------------------------------------------------------------------------
// GroupNorm(32 groups, eps) on NHWC bf16: pass 1 = per-(sample, group) sum / sum-of-squares (fp32 in-block,
// fp64 across blocks), pass 2 = normalise + affine (+ swish).  autoencoder.py:21-22,62-78,156,177-178.
// algorithmic bytes / element: 2 (stats read) + 2 (apply read) + 2 (write)
// ---------------------------------------------------------------------------
// Deterministic: per-thread partials -> fixed-order in-block reduction -> per-CTA partial in the workspace ->
// gn_finalize sums the CTA partials in fp64 in fixed order.  No atomics anywhere (run-to-run bit-identical).
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ partial,
                                                       int64_t HW, int C, int pix_per_cta) {
  __shared__ float sh[256][8];  // [thread][slot*2 + {sum, sumsq}]
  const int b = blockIdx.y;
  const int c8 = C / 8;
  const int cpg = C / 32;
  const int tpp = c8;                       // threads per pixel
  const int ppi = blockDim.x / tpp;         // pixels per iteration (C <= 2048)
  const int my_c = (threadIdx.x % tpp) * 8;
  const int my_p = threadIdx.x / tpp;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_cta;
  const int64_t p1 = min(p0 + pix_per_cta, HW);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};  // one slot per channel pair (cpg >= 2)
  if (my_p < ppi) {
    const bf16* xb = x + static_cast<int64_t>(b) * HW * C + my_c;
    for (int64_t pix = p0 + my_p; pix < p1; pix += 4 * ppi) {
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t pp = pix + static_cast<int64_t>(q) * ppi;
        u[q] = pp < p1 ? *reinterpret_cast<const uint4*>(xb + pp * C) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 a = unpack_bf16x2(u[q].x), c2 = unpack_bf16x2(u[q].y), e = unpack_bf16x2(u[q].z), g = unpack_bf16x2(u[q].w);
        s[0] += a.x + a.y;   ss[0] += a.x * a.x + a.y * a.y;
        s[1] += c2.x + c2.y; ss[1] += c2.x * c2.x + c2.y * c2.y;
        s[2] += e.x + e.y;   ss[2] += e.x * e.x + e.y * e.y;
        s[3] += g.x + g.y;   ss[3] += g.x * g.x + g.y * g.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[threadIdx.x][2 * j] = s[j]; sh[threadIdx.x][2 * j + 1] = ss[j]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int grp = threadIdx.x >> 1, st = threadIdx.x & 1;
    float acc = 0.f;
    for (int p = 0; p < ppi; ++p)
      for (int q = 0; q < cpg / 2; ++q) {
        const int c = grp * cpg + 2 * q;
        acc += sh[p * tpp + c / 8][((c % 8) / 2) * 2 + st];
      }
    partial[((static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * 32 + grp) * 2 + st] = acc;
  }
}

// one thread per (sample, group): fixed-order fp64 sum of the CTA partials -> (mean, rstd)
__global__ void gn_finalize_kernel(const float* __restrict__ partial, float2* __restrict__ mr, int B, int nblk,
                                   double cnt, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 32) return;
  const int b = i / 32, grp = i % 32;
  double sm = 0.0, sq = 0.0;
  for (int k = 0; k < nblk; ++k) {
    const float* p = partial + ((static_cast<int64_t>(b) * nblk + k) * 32 + grp) * 2;
    sm += static_cast<double>(p[0]);
    sq += static_cast<double>(p[1]);
  }
  const double mean = sm / cnt;
  const float var = fmaxf(static_cast<float>(sq / cnt - mean * mean), 0.f);
  mr[i] = make_float2(static_cast<float>(mean), rsqrtf(var + eps));
}

// Each thread owns 8 fixed channels (its scale/shift live in registers for the whole kernel) and walks over
// pixels, 4 independent 16-byte loads in flight per thread.  y = x * (rstd*w) + (b - mean*rstd*w), then swish.
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, const float2* __restrict__ mr,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       bf16* __restrict__ y, int64_t HW, int C, int swish,
                                                       int pix_per_cta) {
  const int b = blockIdx.y;
  const int tpp = C / 8;
  const int ppi = blockDim.x / tpp;
  const int my_c = (threadIdx.x % tpp) * 8;
  const int my_p = threadIdx.x / tpp;
  if (my_p >= ppi) return;
  const int cpg = C / 32;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 st = __ldg(mr + b * 32 + (my_c + k) / cpg);
    const float wk = __ldg(w + my_c + k), bk = __ldg(bias + my_c + k);
    sc[k] = st.y * wk;
    sh[k] = fmaf(-st.x, sc[k], bk);
  }
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_cta;
  const int64_t p1 = min(p0 + pix_per_cta, HW);
  const bf16* xb = x + static_cast<int64_t>(b) * HW * C + my_c;
  bf16* yb = y + static_cast<int64_t>(b) * HW * C + my_c;
  for (int64_t pix = p0 + my_p; pix < p1; pix += 4 * ppi) {
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t pp = pix + static_cast<int64_t>(q) * ppi;
      if (pp < p1) u[q] = __ldcs(reinterpret_cast<const uint4*>(xb + pp * C));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t pp = pix + static_cast<int64_t>(q) * ppi;
      if (pp >= p1) continue;
      const float2 a = unpack_bf16x2(u[q].x), c2 = unpack_bf16x2(u[q].y), e = unpack_bf16x2(u[q].z), g = unpack_bf16x2(u[q].w);
      float v[8] = {a.x, a.y, c2.x, c2.y, e.x, e.y, g.x, g.y};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = fmaf(v[k], sc[k], sh[k]);
        if (swish) t = t * sigmoidf_(t);
        v[k] = t;
      }
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(yb + pp * C) = o;
    }
  }
}

// row softmax: p[r, :n] = softmax(scale * s[r, :n]) (fp32 in, bf16 out), one warp per row; pad columns [n, ld_out) = 0
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, int64_t ld_in, bf16* __restrict__ p,
                                                           int64_t ld_out, int rows, int n, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const float* sr = s + static_cast<int64_t>(r) * ld_in;
  float mx = -INFINITY;
  for (int c = lane; c < n; c += 32) mx = fmaxf(mx, sr[c] * scale);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < n; c += 32) sum += __expf(sr[c] * scale - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  bf16* pr = p + static_cast<int64_t>(r) * ld_out;
  for (int c = lane; c < ld_out; c += 32) pr[c] = __float2bfloat16_rn(c < n ? __expf(sr[c] * scale - mx) * inv : 0.f);
}

// DiagonalGaussian sample + scale/shift + 2x2 patchify (autoencoder.py:268-274,302-305; sampling.py:26):
// moments fp32 NHWC [B,h,w,2z] , noise fp32 NCHW [B,z,h,w] -> x1 fp32 [B,(h/2)(w/2), z*4], inner index c*4+ph*2+pw
__global__ void ae_sample_patchify_kernel(const float* __restrict__ mom, const float* __restrict__ noise,
                                          float* __restrict__ x1, int B, int h, int w, int z, float scale_factor,
                                          float shift_factor) {
  const int64_t n = static_cast<int64_t>(B) * z * h * w;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // i indexes the OUTPUT (coalesced writes)
  const int inner = static_cast<int>(i % (z * 4));
  const int64_t tok = i / (z * 4);
  const int w2 = w / 2, h2 = h / 2;
  const int px = static_cast<int>(tok % w2), py = static_cast<int>((tok / w2) % h2);
  const int b = static_cast<int>(tok / (static_cast<int64_t>(w2) * h2));
  const int c = inner / 4, ph = (inner / 2) & 1, pw = inner & 1;
  const int yy = py * 2 + ph, xx = px * 2 + pw;
  const float* m = mom + ((static_cast<int64_t>(b) * h + yy) * w + xx) * (2 * z);
  const float mean = m[c], logvar = m[z + c];
  const float eps = noise[((static_cast<int64_t>(b) * z + c) * h + yy) * w + xx];
  x1[i] = scale_factor * (mean + __expf(0.5f * logvar) * eps - shift_factor);
}

static inline int grid1d(int64_t n, int block) { return static_cast<int>((n + block - 1) / block); }

}  // namespace gh

using namespace gh;

static int patch_im2col_any(bool u8, const void* img, void* out_bf16, int32_t B, int32_t S, int32_t patch, int64_t ld,
                            const float* mean3, const float* std3, void* stream) {
  using namespace gh;
  GH_REQUIRE(img && out_bf16, GH_ERR_NULL, "gh_patch_im2col: NULL pointer");
  GH_REQUIRE(B > 0 && S > 0 && patch > 0 && S / patch > 0 && ld >= 3 * patch * patch, GH_ERR_BAD_SHAPE,
             "gh_patch_im2col: bad shape");
  const int G = S / patch;
  float m[3] = {0, 0, 0}, is[3] = {1, 1, 1};
  if (mean3 && std3)
    for (int c = 0; c < 3; ++c) { m[c] = mean3[c]; is[c] = 1.f / std3[c]; }  // host pointers (3 floats)
  if (u8)
    patch_im2col_kernel<true><<<dim3(G, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        img, static_cast<bf16*>(out_bf16), S, patch, G, ld, m[0], m[1], m[2], is[0], is[1], is[2]);
  else
    patch_im2col_kernel<false><<<dim3(G, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        img, static_cast<bf16*>(out_bf16), S, patch, G, ld, m[0], m[1], m[2], is[0], is[1], is[2]);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_patch_im2col(const float* img, void* out_bf16, int32_t B, int32_t S, int32_t patch, int64_t ld,
                               const float* mean3, const float* std3, void* stream) {
  return patch_im2col_any(false, img, out_bf16, B, S, patch, ld, mean3, std3, stream);
}
extern "C" int gh_patch_im2col_u8hwc(const void* img_u8, void* out_bf16, int32_t B, int32_t S, int32_t patch, int64_t ld,
                                     const float* mean3, const float* std3, void* stream) {
  return patch_im2col_any(true, img_u8, out_bf16, B, S, patch, ld, mean3, std3, stream);
}

static int im2col3x3_c3_any(bool u8, const void* img, void* out_bf16, int32_t B, int32_t H, int32_t W, float mean,
                            float std, void* stream) {
  using namespace gh;
  GH_REQUIRE(img && out_bf16 && aligned16(out_bf16), GH_ERR_NULL, "gh_im2col3x3_c3: NULL / misaligned pointer");
  GH_REQUIRE(B > 0 && H > 0 && W > 0 && std != 0.f, GH_ERR_BAD_SHAPE, "gh_im2col3x3_c3: bad shape");
  const int64_t n = static_cast<int64_t>(B) * H * W;
  if (u8)
    im2col3x3_c3_kernel<true><<<grid1d(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        img, static_cast<bf16*>(out_bf16), B, H, W, mean, 1.f / std);
  else
    im2col3x3_c3_kernel<false><<<grid1d(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        img, static_cast<bf16*>(out_bf16), B, H, W, mean, 1.f / std);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_im2col3x3_c3(const float* img, void* out_bf16, int32_t B, int32_t H, int32_t W, float mean,
                               float std, void* stream) {
  return im2col3x3_c3_any(false, img, out_bf16, B, H, W, mean, std, stream);
}
extern "C" int gh_im2col3x3_c3_u8hwc(const void* img_u8, void* out_bf16, int32_t B, int32_t H, int32_t W, float mean,
                                     float std, void* stream) {
  return im2col3x3_c3_any(true, img_u8, out_bf16, B, H, W, mean, std, stream);
}

extern "C" int gh_embed_assemble(const void* patch_bf16, const float* cls, const float* pos, void* out_bf16, int32_t B,
                                 int32_t T, int32_t D, int32_t has_cls, void* stream) {
  GH_REQUIRE(patch_bf16 && pos && out_bf16 && (!has_cls || cls), GH_ERR_NULL, "gh_embed_assemble: NULL pointer");
  GH_REQUIRE(B > 0 && T > has_cls && D % 8 == 0 && (has_cls == 0 || has_cls == 1), GH_ERR_BAD_SHAPE,
             "gh_embed_assemble: bad shape");
  const int64_t n8 = static_cast<int64_t>(B) * T * D / 8;
  embed_assemble_kernel<<<grid1d(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(patch_bf16), cls, pos, static_cast<bf16*>(out_bf16), B, T, D, has_cls);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

static void gn_plan(int B, int64_t HW, int* pix_per_cta, int* nblk) {
  int ppc = 256;
  while ((HW + ppc - 1) / ppc * B > 16L * num_sms() && ppc < (1 << 20)) ppc *= 2;
  *pix_per_cta = ppc;
  *nblk = static_cast<int>((HW + ppc - 1) / ppc);
}

extern "C" int64_t gh_groupnorm_ws_bytes(int32_t B, int64_t HW) {
  if (B <= 0 || HW <= 0) return 0;
  int ppc, nblk;
  gn_plan(B, HW, &ppc, &nblk);
  return static_cast<int64_t>(B) * nblk * 64 * sizeof(float) + static_cast<int64_t>(B) * 32 * sizeof(float2);
}

extern "C" int gh_groupnorm_swish_nhwc(const void* x, void* y, int32_t B, int64_t HW, int32_t C, const float* weight,
                                       const float* bias, float eps, int32_t swish, void* ws, void* stream) {
  GH_REQUIRE(x && y && weight && bias && ws, GH_ERR_NULL, "gh_groupnorm_swish_nhwc: NULL pointer");
  GH_REQUIRE(B > 0 && HW > 0 && C % 64 == 0 && C <= 2048, GH_ERR_BAD_SHAPE,
             "gh_groupnorm_swish_nhwc: C=%d must be a multiple of 64 (32 groups of >= 2 channels), <= 2048", C);
  GH_REQUIRE(aligned16(x) && aligned16(y) && aligned16(ws) && aligned16(weight) && aligned16(bias), GH_ERR_ALIGN,
             "gh_groupnorm_swish_nhwc: 16B alignment");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int pix_per_cta, nblk;
  gn_plan(B, HW, &pix_per_cta, &nblk);
  float* partial = static_cast<float*>(ws);
  float2* mr = reinterpret_cast<float2*>(partial + static_cast<int64_t>(B) * nblk * 64);
  gn_stats_kernel<<<dim3(nblk, B), 256, 0, s>>>(static_cast<const bf16*>(x), partial, HW, C, pix_per_cta);
  GH_CHECK_CUDA(cudaGetLastError());
  gn_finalize_kernel<<<(B * 32 + 127) / 128, 128, 0, s>>>(partial, mr, B, nblk, static_cast<double>(HW) * (C / 32), eps);
  GH_CHECK_CUDA(cudaGetLastError());
  gn_apply_kernel<<<dim3(nblk, B), 256, 0, s>>>(static_cast<const bf16*>(x), mr, weight, bias, static_cast<bf16*>(y), HW, C,
                                                swish, pix_per_cta);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

namespace gh {
// nearest-neighbour 2x upsample on NHWC bf16 (FLUX decoder Upsample, autoencoder.py:98-106): one thread = 8 channels
// of one OUTPUT pixel; reads hit L1/L2 (each input vector is read by 4 outputs).  algorithmic bytes / output elem: 0.5 + 2
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H,
                                                         int W, int C8) {
  const int64_t n = static_cast<int64_t>(B) * 2 * H * 2 * W * C8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = static_cast<int>(i % C8);
    int64_t t = i / C8;
    const int ow = static_cast<int>(t % (2 * W));
    t /= 2 * W;
    const int oh = static_cast<int>(t % (2 * H));
    const int b = static_cast<int>(t / (2 * H));
    y[i] = __ldg(x + ((static_cast<int64_t>(b) * H + (oh >> 1)) * W + (ow >> 1)) * C8 + c);
  }
}
// ToTensor of the reference's image pipeline (image_datasets/dataset_cc3m.py:38-44, torchvision ToTensor on a PIL RGB
// image): uint8 HWC -> fp32 CHW in [0,1], value / 255 with IEEE division (bit-identical to torch's .div(255)).
// The decoded image crosses PCIe as 1 byte per value instead of 4; thread = pixel: 96 contiguous bytes per warp in,
// three coalesced 128-byte rows out.
__global__ void __launch_bounds__(256) u8hwc_to_f32chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                              int64_t hw, int64_t total_pix) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_pix; i += stride) {
    const int64_t b = i / hw, p = i - b * hw;
    const uint8_t* s = src + i * 3;
    float* d = dst + b * 3 * hw + p;
    d[0] = __fdiv_rn(static_cast<float>(s[0]), 255.f);
    d[hw] = __fdiv_rn(static_cast<float>(s[1]), 255.f);
    d[2 * hw] = __fdiv_rn(static_cast<float>(s[2]), 255.f);
  }
}

}  // namespace gh

extern "C" int gh_u8hwc_to_f32chw(const void* src_u8, float* dst, int32_t B, int32_t H, int32_t W, void* stream) {
  using namespace gh;
  GH_REQUIRE(B >= 0 && H > 0 && W > 0, GH_ERR_BAD_SHAPE, "gh_u8hwc_to_f32chw: bad shape");
  if (B == 0) return GH_OK;   // (an empty batch has no storage: checked before the pointers)
  GH_REQUIRE(src_u8 && dst, GH_ERR_NULL, "gh_u8hwc_to_f32chw: NULL pointer");
  const int64_t hw = static_cast<int64_t>(H) * W, total = hw * B;
  const int64_t want = (total + 255) / 256, cap = static_cast<int64_t>(num_sms()) * 16;
  u8hwc_to_f32chw_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src_u8), dst, hw, total);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_upsample2x_nhwc(const void* x_bf16, void* y_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  using namespace gh;
  GH_REQUIRE(x_bf16 && y_bf16, GH_ERR_NULL, "gh_upsample2x_nhwc: NULL pointer");
  GH_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, GH_ERR_BAD_SHAPE, "gh_upsample2x_nhwc: bad shape (C %% 8 == 0)");
  GH_REQUIRE(aligned16(x_bf16) && aligned16(y_bf16), GH_ERR_ALIGN, "gh_upsample2x_nhwc: 16B alignment");
  if (B == 0) return GH_OK;
  const int64_t n = static_cast<int64_t>(B) * 4 * H * W * (C / 8);
  const int64_t want = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  upsample2x_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x_bf16), static_cast<uint4*>(y_bf16), B, H, W, C / 8);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_softmax_rows(const float* s, int64_t ld_in, void* p_bf16, int64_t ld_out, int32_t rows, int32_t n,
                               float scale, void* stream) {
  GH_REQUIRE(s && p_bf16, GH_ERR_NULL, "gh_softmax_rows: NULL pointer");
  GH_REQUIRE(rows > 0 && n > 0 && ld_in >= n && ld_out >= n, GH_ERR_BAD_SHAPE, "gh_softmax_rows: bad shape");
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(s, ld_in, static_cast<bf16*>(p_bf16),
                                                                                    ld_out, rows, n, scale);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_ae_sample_patchify(const float* moments_nhwc, const float* noise_nchw, float* x1, int32_t B, int32_t h,
                                     int32_t w, int32_t z, float scale_factor, float shift_factor, void* stream) {
  GH_REQUIRE(moments_nhwc && noise_nchw && x1, GH_ERR_NULL, "gh_ae_sample_patchify: NULL pointer");
  GH_REQUIRE(B > 0 && h > 0 && w > 0 && z > 0 && h % 2 == 0 && w % 2 == 0, GH_ERR_BAD_SHAPE,
             "gh_ae_sample_patchify: latent extent must be even (got %dx%d)", h, w);
  const int64_t n = static_cast<int64_t>(B) * z * h * w;
  ae_sample_patchify_kernel<<<grid1d(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      moments_nhwc, noise_nchw, x1, B, h, w, z, scale_factor, shift_factor);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}
