// genhancer_b200 -- HBM-bound normalisation / modulation / RoPE kernels and their backwards.
// All of them: 128-bit vector loads, fp32 statistics, warp-shuffle reductions, no smem tiling
// (no reuse to exploit).  Row length C is the channel dim; "rpb" = rows (tokens) per sample.
#include "common.cuh"
#include "internal.h"

namespace gh {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

struct RowMap {  // row r = b * rpb + l  ->  element offset b * batch_stride + l * row_stride
  int rpb;
  int64_t batch_stride;
  int64_t row_stride;
  __device__ __forceinline__ int64_t off(int r) const {
    const int b = r / rpb;
    return static_cast<int64_t>(b) * batch_stride + static_cast<int64_t>(r - b * rpb) * row_stride;
  }
};

// Sum over the WPR warps that share one row (WPR = 1: plain warp reduction).  Wide rows (C > 2048) are split
// over two warps so that the row cache stays at 8 x 16 B per lane: with one warp per 3072-wide row the kernels
// needed 144 / 180 registers -> ONE 256-thread block per SM -> 2.1-2.4 TB/s.  The pair meets at its own named
// barrier (64 threads), so blocks whose last pair is past the end need no block-wide sync.
template <int WPR>
__device__ __forceinline__ float row_sum(float v, int warp, int lane, float* red) {
  v = warp_sum(v);
  if (WPR == 1) return v;
  const int grp = warp / WPR;
  const int bar = 1 + grp;          // one named barrier per group of WPR warps (32 * WPR threads)
  if (lane == 0) red[warp] = v;
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(32 * WPR) : "memory");
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < WPR; ++k) t += red[grp * WPR + k];
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(32 * WPR) : "memory");  // red[] may be rewritten by the next reduction
  return t;
}

// ============================================================================
// LayerNorm forward: one warp (or a pair of warps) per row, row cached in registers.
//   y = n * w + b                      (affine, fp32 params)            or
//   y = (1 + scale[b]) * n + shift[b]  (AdaLN, bf16 per-sample vectors) or  y = n
// algorithmic bytes / element: 2 (read) + 2 (write)
// ============================================================================
template <int VMAX, int WPR>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, RowMap xm, bf16* __restrict__ y,
                                                     RowMap ym, int rows, int C, const float* __restrict__ w,
                                                     const float* __restrict__ bias, const bf16* __restrict__ shift,
                                                     const bf16* __restrict__ scale, int64_t mod_ld, float eps,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (8 / WPR) + warp / WPR;
  if (r >= rows) return;
  const int lane_g = (warp % WPR) * 32 + lane;   // position among the 32 * WPR lanes of this row
  const bf16* xr = x + xm.off(r);
  uint4 v[VMAX];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int c = (i * 32 * WPR + lane_g) * 8;
    if (c < C) {
      v[i] = *reinterpret_cast<const uint4*>(xr + c);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
  }
  const float mean = row_sum<WPR>(s, warp, lane, red) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int c = (i * 32 * WPR + lane_g) * 8;
    if (c < C) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(row_sum<WPR>(q, warp, lane, red) / C + eps);
  if (lane_g == 0) {
    if (mean_out) mean_out[r] = mean;
    if (rstd_out) rstd_out[r] = rstd;
  }
  const int b = r / xm.rpb;
  bf16* yr = y + ym.off(r);
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int c = (i * 32 * WPR + lane_g) * 8;
    if (c < C) {
      float f[8], o[8];
      unpack8(v[i], f);
      if (scale) {
        float sc[8], sh[8];
        unpack8(*reinterpret_cast<const uint4*>(scale + static_cast<int64_t>(b) * mod_ld + c), sc);
        unpack8(*reinterpret_cast<const uint4*>(shift + static_cast<int64_t>(b) * mod_ld + c), sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (1.f + sc[j]) * ((f[j] - mean) * rstd) + sh[j];
      } else if (w) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + c), w1 = *reinterpret_cast<const float4*>(w + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c), b1 = *reinterpret_cast<const float4*>(bias + c + 4);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd * ww[j] + bb[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd;
      }
      *reinterpret_cast<uint4*>(yr + c) = pack8(o);
    }
  }
}

// ============================================================================
// LayerNorm backward w.r.t. x (one warp per row):
//   dn = dy * g,  g = w | (1 + scale[b]) | 1 ;  dx = rstd * (dn - mean(dn) - n * mean(dn * n)) [+ dres]
// ============================================================================
template <int VMAX, int WPR>
__global__ void __launch_bounds__(256) ln_bwd_dx_kernel(const bf16* __restrict__ dy, RowMap dym,
                                                        const bf16* __restrict__ x, RowMap xm, int rows, int C,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ w, const bf16* __restrict__ scale,
                                                        int64_t mod_ld, const bf16* __restrict__ dres, RowMap drm,
                                                        bf16* __restrict__ dx, RowMap dxm) {
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (8 / WPR) + warp / WPR;
  if (r >= rows) return;
  const int lane_g = (warp % WPR) * 32 + lane;
  const bf16* xr = x + xm.off(r);
  const bf16* dyr = dy + dym.off(r);
  const int b = r / xm.rpb;
  const float mu = mean[r], rs = rstd[r];
  const bf16* drr = dres ? dres + drm.off(r) : nullptr;
  uint4 vn[VMAX], vd[VMAX], vr[VMAX];  // raw x, dy and (requested with them, used after the reductions) dres
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int c = (i * 32 * WPR + lane_g) * 8;
    if (c < C) {
      vn[i] = *reinterpret_cast<const uint4*>(xr + c);
      vd[i] = *reinterpret_cast<const uint4*>(dyr + c);
      if (drr) vr[i] = *reinterpret_cast<const uint4*>(drr + c);
      float fx[8], fd[8];
      unpack8(vn[i], fx);
      unpack8(vd[i], fd);
      float g[8];
      if (scale) {
        unpack8(*reinterpret_cast<const uint4*>(scale + static_cast<int64_t>(b) * mod_ld + c), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += 1.f;
      } else if (w) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + c), w1 = *reinterpret_cast<const float4*>(w + c + 4);
        g[0] = w0.x; g[1] = w0.y; g[2] = w0.z; g[3] = w0.w; g[4] = w1.x; g[5] = w1.y; g[6] = w1.z; g[7] = w1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 1.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        fd[j] *= g[j];
        s1 += fd[j];
        s2 += fd[j] * ((fx[j] - mu) * rs);
      }
      // keep dn in fp32 precision would need 2x registers; re-derive from dy*g below instead
    }
  }
  const float m1 = row_sum<WPR>(s1, warp, lane, red) / C, m2 = row_sum<WPR>(s2, warp, lane, red) / C;
  bf16* dxr = dx + dxm.off(r);
#pragma unroll
  for (int i = 0; i < VMAX; ++i) {
    const int c = (i * 32 * WPR + lane_g) * 8;
    if (c < C) {
      float fx[8], fd[8], g[8], o[8];
      unpack8(vn[i], fx);
      unpack8(vd[i], fd);
      if (scale) {
        unpack8(*reinterpret_cast<const uint4*>(scale + static_cast<int64_t>(b) * mod_ld + c), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += 1.f;
      } else if (w) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + c), w1 = *reinterpret_cast<const float4*>(w + c + 4);
        g[0] = w0.x; g[1] = w0.y; g[2] = w0.z; g[3] = w0.w; g[4] = w1.x; g[5] = w1.y; g[6] = w1.z; g[7] = w1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 1.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float n = (fx[j] - mu) * rs;
        o[j] = rs * (fd[j] * g[j] - m1 - n * m2);
      }
      if (drr) {
        float rr[8];
        unpack8(vr[i], rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rr[j];
      }
      *reinterpret_cast<uint4*>(dxr + c) = pack8(o);
    }
  }
}

// ============================================================================
// Column reductions (thread owns 8 columns, loops over a chunk of rows of ONE sample):
//   MODE 0 (LN params):  acc0[b,c] += sum_l dy            acc1[b,c] += sum_l dy * n(x)
//   MODE 1 (gate):       du = gate[b] * dy (written)      acc0[b,c] += sum_l dy * u     acc1[c] += sum_l du  (optional:
//                        the bias gradient of the linear whose output u is -- du is that linear's dY, so its column
//                        sum comes out of this pass instead of a separate read of du)
//   MODE 2 (bias):       acc0[b,c] += sum_l dy
// grid = (ceil(rpb / ROWS), B, ceil(C / (8 * blockDim.x)))
// ============================================================================
template <int MODE>
__global__ void __launch_bounds__(512) col_reduce_kernel(const bf16* __restrict__ dy, RowMap dym,
                                                         const bf16* __restrict__ x, RowMap xm, int C,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const bf16* __restrict__ gate, int64_t gate_ld,
                                                         bf16* __restrict__ du, RowMap dum, float* __restrict__ acc0,
                                                         float* __restrict__ acc1, int64_t acc_ld, int64_t acc1_ld,
                                                         int rows_per_cta) {
  // blockDim = (C / 8 column groups, ny row lanes): thread (tx, ty) owns 8 columns and the rows l0 + ty + k * ny of
  // its CTA's chunk; the ny partial sums meet in shared memory, so one CTA issues ONE atomic per column however
  // many rows it covers (same-address atomics serialise in L2: fewer, fatter CTAs beat many small ones).
  extern __shared__ float cr_smem[];
  const int ny = blockDim.y, ty = threadIdx.y;
  const int c = (blockIdx.z * blockDim.x + threadIdx.x) * 8;
  const bool live = c < C;
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * rows_per_cta;
  const int l1 = min(l0 + rows_per_cta, dym.rpb);
  float a0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live && l0 + ty < l1) {
    float g[8];
    if (MODE == 1) unpack8(*reinterpret_cast<const uint4*>(gate + static_cast<int64_t>(b) * gate_ld + c), g);
    // the chunk lies inside ONE sample: resolve the row maps once, then walk by the row strides, four rows per trip
    // with all loads issued before the math (a per-row `off()` division serialised the address stream)
    const int r0 = b * dym.rpb + l0 + ty;
    const int64_t sd = dym.row_stride * ny, sx = xm.row_stride * ny, su = dum.row_stride * ny;
    const bf16* pd = dy + dym.off(r0) + c;
    const bf16* px = (MODE != 2) ? x + xm.off(r0) + c : nullptr;
    bf16* pu = (MODE == 1) ? du + dum.off(r0) + c : nullptr;
    auto one = [&](const uint4& vd, const uint4& vx, int r, bf16* out) {
      float fd[8];
      unpack8(vd, fd);
      if (MODE == 0) {
        float fx[8];
        unpack8(vx, fx);
        const float mu = mean[r], rs = rstd[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += fd[j]; a1[j] += fd[j] * ((fx[j] - mu) * rs); }
      } else if (MODE == 1) {
        float fu[8], o[8];
        unpack8(vx, fu);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += fd[j] * fu[j]; o[j] = g[j] * fd[j]; a1[j] += o[j]; }
        *reinterpret_cast<uint4*>(out) = pack8(o);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a0[j] += fd[j];
      }
    };
    int l = l0 + ty, r = r0;
    constexpr int U = MODE == 2 ? 8 : 4;   // rows in flight per thread (the plain column sum has registers to spare)
    for (; l + (U - 1) * ny < l1; l += U * ny, r += U * ny) {
      uint4 vd[U], vx[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        vd[u] = *reinterpret_cast<const uint4*>(pd + u * sd);
        if (MODE != 2) vx[u] = *reinterpret_cast<const uint4*>(px + u * sx);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) one(vd[u], vx[u], r + u * ny, MODE == 1 ? pu + u * su : nullptr);
      pd += U * sd;
      if (MODE != 2) px += U * sx;
      if (MODE == 1) pu += U * su;
    }
    for (; l < l1; l += ny, r += ny) {
      const uint4 vd = *reinterpret_cast<const uint4*>(pd);
      uint4 vx = make_uint4(0, 0, 0, 0);
      if (MODE != 2) vx = *reinterpret_cast<const uint4*>(px);
      one(vd, vx, r, pu);
      pd += sd;
      if (MODE != 2) px += sx;
      if (MODE == 1) pu += su;
    }
  }
  if (ny > 1) {  // fold the row lanes: [ny][blockDim.x][16] floats
    float* mine = cr_smem + (static_cast<size_t>(ty) * blockDim.x + threadIdx.x) * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mine[j] = a0[j]; mine[8 + j] = a1[j]; }
    __syncthreads();
    if (ty != 0) return;
    for (int y = 1; y < ny; ++y) {
      const float* o = cr_smem + (static_cast<size_t>(y) * blockDim.x + threadIdx.x) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { a0[j] += o[j]; a1[j] += o[8 + j]; }
    }
  }
  if (!live) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(acc0 + static_cast<int64_t>(b) * acc_ld + c + j, a0[j]);
    if (MODE != 2 && acc1) atomicAdd(acc1 + static_cast<int64_t>(b) * acc1_ld + c + j, a1[j]);
  }
}

// ============================================================================
// RoPE table: ids [B, L, 3] fp32 -> cs [B, L, D/2] float2 (cos, sin), angles in float64
// (reference builds omega and the angle in float64, src/flux/math.py:15-22)
// ============================================================================
__global__ void rope_table_kernel(const float* __restrict__ ids, float2* __restrict__ cs, int64_t n_tok, int a0,
                                  int a1, int a2, double theta) {
  const int half = (a0 + a1 + a2) / 2;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_tok * half) return;
  const int64_t tok = i / half;
  int j = static_cast<int>(i - tok * half);
  int axis, dim;
  if (j < a0 / 2) { axis = 0; dim = a0; }
  else if (j < (a0 + a1) / 2) { axis = 1; dim = a1; j -= a0 / 2; }
  else { axis = 2; dim = a2; j -= (a0 + a1) / 2; }
  const double omega = 1.0 / pow(theta, static_cast<double>(2 * j) / dim);
  const double ang = static_cast<double>(ids[tok * 3 + axis]) * omega;
  cs[i] = make_float2(static_cast<float>(cos(ang)), static_cast<float>(sin(ang)));
}

// ============================================================================
// QK RMSNorm + RoPE + head-major scatter, D = 128, one warp per (token, head), lane owns 4 dims.
//   qkv [B, L, 3, H, D] (row pitch ld_qkv) -> q, k, v at [B, H, Ltot, D], token offset l_off.
// Roundings follow the reference: bf16(x * rrms) -> bf16(. * scale) -> fp32 rope -> bf16.
// ============================================================================
__global__ void __launch_bounds__(256) qk_norm_rope_fwd_kernel(const bf16* __restrict__ qkv, int64_t ld_qkv, int B,
                                                               int L, int H, int Ltot, int l_off,
                                                               const bf16* __restrict__ q_scale,
                                                               const bf16* __restrict__ k_scale,
                                                               const float2* __restrict__ cs, int64_t cs_batch_stride,
                                                               bf16* __restrict__ q, bf16* __restrict__ k,
                                                               bf16* __restrict__ v) {
  constexpr int D = 128;
  const int lane = threadIdx.x & 31;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (wid >= static_cast<int64_t>(B) * L * H) return;
  const int h = static_cast<int>(wid % H);
  const int64_t tok = wid / H;
  const int l = static_cast<int>(tok % L);
  const int b = static_cast<int>(tok / L);
  const bf16* src = qkv + tok * ld_qkv + h * D + lane * 4;
  const int64_t dst = ((static_cast<int64_t>(b) * H + h) * Ltot + l_off + l) * D + lane * 4;
  const float2* csr = cs + b * cs_batch_stride + static_cast<int64_t>(l_off + l) * (D / 2) + lane * 2;
  const float2 cs0 = csr[0], cs1 = csr[1];
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const uint2 raw = *reinterpret_cast<const uint2*>(src + which * H * D);
    const float2 p0 = unpack_bf16x2(raw.x), p1 = unpack_bf16x2(raw.y);
    const float ss = warp_sum(p0.x * p0.x + p0.y * p0.y + p1.x * p1.x + p1.y * p1.y);
    const float rr = rsqrtf(ss / D + 1e-6f);
    const uint2 sraw = *reinterpret_cast<const uint2*>((which ? k_scale : q_scale) + lane * 4);
    const float2 s0 = unpack_bf16x2(sraw.x), s1 = unpack_bf16x2(sraw.y);
    auto rnd = [](float f) { return __bfloat162float(__float2bfloat16_rn(f)); };
    const float x0 = rnd(rnd(p0.x * rr) * s0.x), x1 = rnd(rnd(p0.y * rr) * s0.y);
    const float x2 = rnd(rnd(p1.x * rr) * s1.x), x3 = rnd(rnd(p1.y * rr) * s1.y);
    uint2 o;
    o.x = pack_bf16x2(cs0.x * x0 - cs0.y * x1, cs0.y * x0 + cs0.x * x1);
    o.y = pack_bf16x2(cs1.x * x2 - cs1.y * x3, cs1.y * x2 + cs1.x * x3);
    *reinterpret_cast<uint2*>((which ? k : q) + dst) = o;
  }
  *reinterpret_cast<uint2*>(v + dst) = *reinterpret_cast<const uint2*>(src + 2 * H * D);
}

// backward: dq, dk, dv [B,H,Ltot,D] -> dqkv [B,L,3,H,D]; dscale_q/k [D] fp32 accumulated.
__global__ void __launch_bounds__(256, 3) qk_norm_rope_bwd_kernel(const bf16* __restrict__ dq, const bf16* __restrict__ dk,
                                                               const bf16* __restrict__ dv,
                                                               const bf16* __restrict__ qkv, int64_t ld_qkv, int B,
                                                               int L, int H, int Ltot, int l_off,
                                                               const bf16* __restrict__ q_scale,
                                                               const bf16* __restrict__ k_scale,
                                                               const float2* __restrict__ cs, int64_t cs_batch_stride,
                                                               bf16* __restrict__ dqkv, int64_t ld_dqkv,
                                                               float* __restrict__ dscale_q,
                                                               float* __restrict__ dscale_k, int tokens_per_cta) {
  // A warp handles TWO heads of one token at a time: 16 lanes per 128-wide head, 8 elements (16 bytes) per lane, so
  // every load / store is a full 16-byte vector and one 4-step shuffle reduction serves both heads (the first version
  // used 8-byte accesses, one head per warp and 5-step reductions: issue-bound at 3.2 TB/s).
  constexpr int D = 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, sl = lane & 15;
  __shared__ float red[2][16][D];
  float acc[2][8];
#pragma unroll
  for (int w = 0; w < 2; ++w)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[w][j] = 0.f;
  auto half_sum = [](float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  const int HP = (H + 1) / 2;   // head pairs per token
  const int64_t n_tok = static_cast<int64_t>(B) * L;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * tokens_per_cta;
  const int64_t t1 = min(t0 + tokens_per_cta, n_tok);
  float qs[8], ks[8];
  unpack8(*reinterpret_cast<const uint4*>(q_scale + sl * 8), qs);
  unpack8(*reinterpret_cast<const uint4*>(k_scale + sl * 8), ks);
  for (int64_t it = t0 * HP + warp; it < t1 * HP; it += 8) {
    const int h = 2 * static_cast<int>(it % HP) + half;
    const bool live = h < H;            // odd H: the upper half of the last pair idles (but joins the shuffles)
    const int hh = live ? h : H - 1;
    const int64_t tok = it / HP;
    const int l = static_cast<int>(tok % L);
    const int b = static_cast<int>(tok / L);
    const bf16* src = qkv + tok * ld_qkv + hh * D + sl * 8;
    bf16* dstp = dqkv + tok * ld_dqkv + hh * D + sl * 8;
    const int64_t hm = ((static_cast<int64_t>(b) * H + hh) * Ltot + l_off + l) * D + sl * 8;
    const float4* csr = reinterpret_cast<const float4*>(cs + b * cs_batch_stride +
                                                        static_cast<int64_t>(l_off + l) * (D / 2) + sl * 4);
    // all loads of the iteration first
    const float4 c01 = csr[0], c23 = csr[1];   // (cos, sin) of rotation pairs 4 sl .. 4 sl + 3
    const uint4 gq = *reinterpret_cast<const uint4*>(dq + hm), gk = *reinterpret_cast<const uint4*>(dk + hm);
    const uint4 gv = *reinterpret_cast<const uint4*>(dv + hm);
    const uint4 xq = *reinterpret_cast<const uint4*>(src), xk = *reinterpret_cast<const uint4*>(src + H * D);
    const float cc[4] = {c01.x, c01.z, c23.x, c23.z}, sn[4] = {c01.y, c01.w, c23.y, c23.w};
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float g[8], x[8];
      unpack8(which ? gk : gq, g);
      unpack8(which ? xk : xq, x);
      float dy[8], ss = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // inverse rotation
        dy[2 * j] = cc[j] * g[2 * j] + sn[j] * g[2 * j + 1];
        dy[2 * j + 1] = -sn[j] * g[2 * j] + cc[j] * g[2 * j + 1];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += x[j] * x[j];
      ss = half_sum(ss);
      const float rr = rsqrtf(ss / D + 1e-6f);
      float dxn[8], dot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xn = x[j] * rr;
        x[j] = xn;
        if (live) acc[which][j] += dy[j] * xn;
        dxn[j] = dy[j] * (which ? ks[j] : qs[j]);
        dot += dxn[j] * xn;
      }
      dot = half_sum(dot) / D;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rr * (dxn[j] - x[j] * dot);
      if (live) *reinterpret_cast<uint4*>(dstp + which * H * D) = pack8(o);
    }
    if (live) *reinterpret_cast<uint4*>(dstp + 2 * H * D) = gv;
  }
#pragma unroll
  for (int which = 0; which < 2; ++which)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[which][warp * 2 + half][sl * 8 + j] = acc[which][j];
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    const int which = i / D, d = i % D;
    float s = 0.f;
#pragma unroll
    for (int w16 = 0; w16 < 16; ++w16) s += red[which][w16][d];
    atomicAdd((which ? dscale_k : dscale_q) + d, s);
  }
}

// ============================================================================
// small elementwise kernels
// ============================================================================
// timestep_embedding (layers.py:28-49): out[b, 0:128] = cos(tt * f_k), out[b, 128:256] = sin(tt * f_k),
// tt = bf16(1000 * bf16(t)) when round_bf16 (the scripts hand the DiT a bf16 t), else 1000 * t.
__global__ void timestep_embed_kernel(const float* __restrict__ t, bf16* __restrict__ out, int B, int round_bf16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 128) return;
  const int b = i / 128, k = i % 128;
  float tv = t[b];
  if (round_bf16) {
    tv = __bfloat162float(__float2bfloat16_rn(tv));
    tv = __bfloat162float(__float2bfloat16_rn(1000.f * tv));
  } else {
    tv = 1000.f * tv;
  }
  const float f = expf(-9.210340371976184f * static_cast<float>(k) / 128.f);  // ln(10000)
  const float a = tv * f;
  out[b * 256 + k] = __float2bfloat16_rn(cosf(a));
  out[b * 256 + 128 + k] = __float2bfloat16_rn(sinf(a));
}

__global__ void act_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t n8, int act) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(x)[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_fwd(act, f[j]);
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
}
// dx = dy * act'(x)
__global__ void act_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx,
                               int64_t n8, int act) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float f[8], g[8];
    unpack8(reinterpret_cast<const uint4*>(x)[i], f);
    unpack8(reinterpret_cast<const uint4*>(dy)[i], g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= act_bwd(act, f[j]);
    reinterpret_cast<uint4*>(dx)[i] = pack8(g);
  }
}

// dst(bf16|fp32) (+)= scale * src(fp32)   -- flushes the fp32 small-gradient scratch into .grad
template <typename T>
__global__ void accum_cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n, float scale,
                                  int accumulate) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = scale * src[i];
    if (accumulate) v += static_cast<float>(dst[i]);
    dst[i] = static_cast<T>(v);
  }
}

static inline int grid_for(int64_t n, int block, int cap_mult = 8) {
  const int64_t want = (n + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * cap_mult;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

static RowMap mk(int rpb, int64_t bs, int64_t rs) {
  RowMap m;
  m.rpb = rpb > 0 ? rpb : 0x7fffffff;  // <= 0: one flat batch
  m.batch_stride = bs;
  m.row_stride = rs;
  return m;
}

}  // namespace gh

using namespace gh;

// A "rows view": rows = B * rpb rows of C channels, row (b, l) at base + b * batch_stride + l * row_stride.
#define GH_VIEW_OK(ptr, v) ((ptr) != nullptr && aligned16(ptr) && (v).row_stride % 8 == 0 && (v).batch_stride % 8 == 0)

extern "C" int gh_layernorm_fwd(const void* x, const gh_rows_view* xv, void* y, const gh_rows_view* yv, int32_t rows,
                                int32_t C, const float* weight, const float* bias, const void* shift,
                                const void* scale, int64_t mod_ld, float eps, float* mean_out, float* rstd_out,
                                void* stream) {
  GH_REQUIRE(x && y && xv && yv, GH_ERR_NULL, "gh_layernorm_fwd: NULL pointer");
  if (rows == 0) return GH_OK;
  GH_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 4096, GH_ERR_BAD_SHAPE,
             "gh_layernorm_fwd: need rows>0 and C a multiple of 8 <= 4096 (rows=%d C=%d)", rows, C);
  GH_REQUIRE(GH_VIEW_OK(x, *xv) && GH_VIEW_OK(y, *yv), GH_ERR_ALIGN, "gh_layernorm_fwd: misaligned view");
  GH_REQUIRE((shift == nullptr) == (scale == nullptr), GH_ERR_NULL, "gh_layernorm_fwd: shift and scale go together");
  GH_REQUIRE(!(weight && scale), GH_ERR_UNSUPPORTED, "gh_layernorm_fwd: affine and AdaLN are exclusive");
  GH_REQUIRE(!weight || bias, GH_ERR_NULL, "gh_layernorm_fwd: affine needs bias");
  GH_REQUIRE(!scale || mod_ld % 8 == 0, GH_ERR_ALIGN, "gh_layernorm_fwd: mod_ld must be a multiple of 8");
  const RowMap xm = mk(xv->rows_per_batch, xv->batch_stride, xv->row_stride);
  const RowMap ym = mk(yv->rows_per_batch, yv->batch_stride, yv->row_stride);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define GH_LN(V, W)                                                                                                  \
  ln_fwd_kernel<V, W><<<(rows + 8 / W - 1) / (8 / W), 256, 0, s>>>(                                                   \
      static_cast<const bf16*>(x), xm, static_cast<bf16*>(y), ym, rows, C, weight, bias,                            \
      static_cast<const bf16*>(shift), static_cast<const bf16*>(scale), mod_ld, eps, mean_out, rstd_out)
  if (C <= 1024) GH_LN(4, 1);
  else if (C <= 2048) GH_LN(8, 1);
  else if (C <= 3072) GH_LN(3, 4);   // wide rows: four warps per row (see row_sum)
  else GH_LN(4, 4);
#undef GH_LN
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_layernorm_bwd_dx(const void* dy, const gh_rows_view* dyv, const void* x, const gh_rows_view* xv,
                                   int32_t rows, int32_t C, const float* mean, const float* rstd, const float* weight,
                                   const void* scale, int64_t mod_ld, const void* dres, const gh_rows_view* drv,
                                   void* dx, const gh_rows_view* dxv, void* stream) {
  GH_REQUIRE(dy && x && dx && mean && rstd && dyv && xv && dxv, GH_ERR_NULL, "gh_layernorm_bwd_dx: NULL pointer");
  if (rows == 0) return GH_OK;
  GH_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 4096, GH_ERR_BAD_SHAPE, "gh_layernorm_bwd_dx: bad shape");
  GH_REQUIRE(GH_VIEW_OK(dy, *dyv) && GH_VIEW_OK(x, *xv) && GH_VIEW_OK(dx, *dxv) && (!dres || (drv && GH_VIEW_OK(dres, *drv))),
             GH_ERR_ALIGN, "gh_layernorm_bwd_dx: misaligned view");
  const RowMap dym = mk(dyv->rows_per_batch, dyv->batch_stride, dyv->row_stride);
  const RowMap xm = mk(xv->rows_per_batch, xv->batch_stride, xv->row_stride);
  const RowMap dxm = mk(dxv->rows_per_batch, dxv->batch_stride, dxv->row_stride);
  const RowMap drm = dres ? mk(drv->rows_per_batch, drv->batch_stride, drv->row_stride) : dxm;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define GH_LNB(V, W)                                                                                               \
  ln_bwd_dx_kernel<V, W><<<(rows + 8 / W - 1) / (8 / W), 256, 0, s>>>(                                              \
      static_cast<const bf16*>(dy), dym, static_cast<const bf16*>(x), xm, rows, C, mean, rstd, weight,             \
      static_cast<const bf16*>(scale), mod_ld, static_cast<const bf16*>(dres), drm, static_cast<bf16*>(dx), dxm)
  if (C <= 1024) GH_LNB(4, 1);
  else if (C <= 2048) GH_LNB(8, 1);
  else if (C <= 3072) GH_LNB(3, 4);   // four warps per row: 3 x 16 B per lane and tensor, ~3 blocks per SM
  else GH_LNB(4, 4);
#undef GH_LNB
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

static int launch_col_reduce(int mode, const void* dy, const gh_rows_view* dyv, const void* x, const gh_rows_view* xv,
                             int32_t batches, int32_t C, const float* mean, const float* rstd, const void* gate,
                             int64_t gate_ld, void* du, const gh_rows_view* duv, float* acc0, float* acc1,
                             int64_t acc_ld, int64_t acc1_ld, void* stream, const char* who) {
  GH_REQUIRE(dy && dyv && acc0, GH_ERR_NULL, "%s: NULL pointer", who);
  if (batches == 0 || dyv->rows_per_batch == 0) return GH_OK;
  GH_REQUIRE(C > 0 && C % 8 == 0, GH_ERR_BAD_SHAPE, "%s: C=%d must be a multiple of 8", who, C);
  GH_REQUIRE(GH_VIEW_OK(dy, *dyv) && (!x || (xv && GH_VIEW_OK(x, *xv))) && (!du || (duv && GH_VIEW_OK(du, *duv))),
             GH_ERR_ALIGN, "%s: misaligned view", who);
  const RowMap dym = mk(dyv->rows_per_batch, dyv->batch_stride, dyv->row_stride);
  const RowMap xm = x ? mk(xv->rows_per_batch, xv->batch_stride, xv->row_stride) : dym;
  const RowMap dum = du ? mk(duv->rows_per_batch, duv->batch_stride, duv->row_stride) : dym;
  const int threads = (C / 8) < 512 ? ((C / 8 + 31) / 32 * 32) : 512;
  const int gz = (C / 8 + threads - 1) / threads;
  int ny = 512 / threads;                 // row lanes per CTA (narrow rows: several rows side by side)
  if (ny < 1) ny = 1;
  // about two 512-thread CTAs per SM, each walking >= 16 rows per lane: one atomic per column and CTA at the end
  const int64_t want_ctas = 2L * num_sms();
  int rows_per_cta = 16 * ny;
  while (static_cast<int64_t>((dym.rpb + rows_per_cta - 1) / rows_per_cta) * batches * gz > want_ctas && rows_per_cta < 8192)
    rows_per_cta *= 2;
  dim3 grid((dym.rpb + rows_per_cta - 1) / rows_per_cta, batches, gz);
  dim3 block(threads, ny);
  const size_t smem = ny > 1 ? static_cast<size_t>(ny) * threads * 16 * sizeof(float) : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define GH_CR(M)                                                                                                   \
  col_reduce_kernel<M><<<grid, block, smem, s>>>(static_cast<const bf16*>(dy), dym, static_cast<const bf16*>(x), xm, C, \
                                                mean, rstd, static_cast<const bf16*>(gate), gate_ld,                  \
                                                static_cast<bf16*>(du), dum, acc0, acc1, acc_ld, acc1_ld, rows_per_cta)
  if (mode == 0) GH_CR(0);
  else if (mode == 1) GH_CR(1);
  else GH_CR(2);
#undef GH_CR
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_layernorm_bwd_params(const void* dy, const gh_rows_view* dyv, const void* x, const gh_rows_view* xv,
                                       int32_t batches, int32_t C, const float* mean, const float* rstd,
                                       float* dshift_acc, float* dscale_acc, int64_t acc_ld, void* stream) {
  GH_REQUIRE(x && mean && rstd, GH_ERR_NULL, "gh_layernorm_bwd_params: NULL pointer");
  return launch_col_reduce(0, dy, dyv, x, xv, batches, C, mean, rstd, nullptr, 0, nullptr, nullptr, dshift_acc,
                           dscale_acc, acc_ld, acc_ld, stream, "gh_layernorm_bwd_params");
}

extern "C" int gh_gate_bwd(const void* dout, const gh_rows_view* dov, const void* u, const gh_rows_view* uv,
                           int32_t batches, int32_t C, const void* gate, int64_t gate_ld, void* du,
                           const gh_rows_view* duv, float* dgate_acc, int64_t acc_ld, float* dbias_acc, void* stream) {
  GH_REQUIRE(u && gate && du, GH_ERR_NULL, "gh_gate_bwd: NULL pointer");
  GH_REQUIRE(gate_ld % 8 == 0, GH_ERR_ALIGN, "gh_gate_bwd: gate_ld must be a multiple of 8");
  return launch_col_reduce(1, dout, dov, u, uv, batches, C, nullptr, nullptr, gate, gate_ld, du, duv, dgate_acc,
                           dbias_acc, acc_ld, 0, stream, "gh_gate_bwd");
}

extern "C" int gh_colsum(const void* dy, const gh_rows_view* dyv, int32_t batches, int32_t C, float* acc,
                         int64_t acc_ld, void* stream) {
  return launch_col_reduce(2, dy, dyv, nullptr, nullptr, batches, C, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
                           acc, nullptr, acc_ld, 0, stream, "gh_colsum");
}

extern "C" int gh_rope_table(const float* ids, void* cos_sin, int64_t n_tokens, int32_t axis0, int32_t axis1,
                             int32_t axis2, double theta, void* stream) {
  GH_REQUIRE(ids && cos_sin, GH_ERR_NULL, "gh_rope_table: NULL pointer");
  if (n_tokens == 0) return GH_OK;
  GH_REQUIRE(axis0 % 2 == 0 && axis1 % 2 == 0 && axis2 % 2 == 0 && axis0 + axis1 + axis2 > 0, GH_ERR_BAD_SHAPE,
             "gh_rope_table: axes dims must be even");
  const int64_t n = n_tokens * ((axis0 + axis1 + axis2) / 2);
  rope_table_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, static_cast<float2*>(cos_sin), n_tokens, axis0, axis1, axis2, theta);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_qk_norm_rope_fwd(const void* qkv, int64_t ld_qkv, int32_t B, int32_t L, int32_t H, int32_t D,
                                   int32_t Ltot, int32_t l_off, const void* q_scale, const void* k_scale,
                                   const void* cos_sin, int64_t cs_batch_stride, void* q, void* k, void* v,
                                   void* stream) {
  GH_REQUIRE(qkv && q_scale && k_scale && cos_sin && q && k && v, GH_ERR_NULL, "gh_qk_norm_rope_fwd: NULL pointer");
  GH_REQUIRE(D == 128, GH_ERR_UNSUPPORTED, "gh_qk_norm_rope_fwd: head dim %d unsupported (DiT uses 128)", D);
  GH_REQUIRE(B >= 0 && L >= 0 && H > 0 && l_off >= 0 && l_off + L <= Ltot, GH_ERR_BAD_SHAPE,
             "gh_qk_norm_rope_fwd: bad shape");
  GH_REQUIRE(ld_qkv % 4 == 0 && ld_qkv >= 3LL * H * D, GH_ERR_ALIGN, "gh_qk_norm_rope_fwd: bad ld_qkv");
  const int64_t n = static_cast<int64_t>(B) * L * H;
  if (n == 0) return GH_OK;
  qk_norm_rope_fwd_kernel<<<static_cast<int>((n + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(qkv), ld_qkv, B, L, H, Ltot, l_off, static_cast<const bf16*>(q_scale),
      static_cast<const bf16*>(k_scale), static_cast<const float2*>(cos_sin), cs_batch_stride, static_cast<bf16*>(q),
      static_cast<bf16*>(k), static_cast<bf16*>(v));
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_qk_norm_rope_bwd(const void* dq, const void* dk, const void* dv, const void* qkv, int64_t ld_qkv,
                                   int32_t B, int32_t L, int32_t H, int32_t D, int32_t Ltot, int32_t l_off,
                                   const void* q_scale, const void* k_scale, const void* cos_sin,
                                   int64_t cs_batch_stride, void* dqkv, int64_t ld_dqkv, float* dscale_q_acc,
                                   float* dscale_k_acc, void* stream) {
  GH_REQUIRE(dq && dk && dv && qkv && q_scale && k_scale && cos_sin && dqkv && dscale_q_acc && dscale_k_acc,
             GH_ERR_NULL, "gh_qk_norm_rope_bwd: NULL pointer");
  GH_REQUIRE(D == 128, GH_ERR_UNSUPPORTED, "gh_qk_norm_rope_bwd: head dim %d unsupported", D);
  GH_REQUIRE(B >= 0 && L >= 0 && H > 0 && l_off >= 0 && l_off + L <= Ltot, GH_ERR_BAD_SHAPE,
             "gh_qk_norm_rope_bwd: bad shape");
  GH_REQUIRE(ld_qkv % 8 == 0 && ld_dqkv % 8 == 0 && cs_batch_stride % 2 == 0 && aligned16(dq) && aligned16(dk) &&
                 aligned16(dv) && aligned16(qkv) && aligned16(dqkv) && aligned16(cos_sin) && aligned16(q_scale) &&
                 aligned16(k_scale),
             GH_ERR_ALIGN, "gh_qk_norm_rope_bwd: 16-byte vectors need ld % 8 == 0 and 16-byte aligned pointers");
  const int64_t n_tok = static_cast<int64_t>(B) * L;
  if (n_tok == 0) return GH_OK;
  int tokens_per_cta = 8;
  while ((n_tok + tokens_per_cta - 1) / tokens_per_cta > 8L * num_sms() && tokens_per_cta < 256) tokens_per_cta *= 2;
  const int grid = static_cast<int>((n_tok + tokens_per_cta - 1) / tokens_per_cta);
  qk_norm_rope_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(dq), static_cast<const bf16*>(dk), static_cast<const bf16*>(dv),
      static_cast<const bf16*>(qkv), ld_qkv, B, L, H, Ltot, l_off, static_cast<const bf16*>(q_scale),
      static_cast<const bf16*>(k_scale), static_cast<const float2*>(cos_sin), cs_batch_stride,
      static_cast<bf16*>(dqkv), ld_dqkv, dscale_q_acc, dscale_k_acc, tokens_per_cta);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_timestep_embedding(const float* t, void* out_bf16, int32_t B, int32_t round_bf16, void* stream) {
  GH_REQUIRE(t && out_bf16, GH_ERR_NULL, "gh_timestep_embedding: NULL pointer");
  if (B <= 0) return B == 0 ? GH_OK : set_error(GH_ERR_BAD_SHAPE, "gh_timestep_embedding: B<0");
  timestep_embed_kernel<<<(B * 128 + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      t, static_cast<bf16*>(out_bf16), B, round_bf16);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_act_fwd(const void* x, void* y, int64_t numel, int32_t act, void* stream) {
  GH_REQUIRE(x && y, GH_ERR_NULL, "gh_act_fwd: NULL pointer");
  if (numel == 0) return GH_OK;
  GH_REQUIRE(numel > 0 && numel % 8 == 0 && aligned16(x) && aligned16(y), GH_ERR_ALIGN,
             "gh_act_fwd: numel must be a multiple of 8 and pointers 16B aligned");
  act_kernel<<<grid_for(numel / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), numel / 8, act);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_act_bwd(const void* dy, const void* x, void* dx, int64_t numel, int32_t act, void* stream) {
  GH_REQUIRE(dy && x && dx, GH_ERR_NULL, "gh_act_bwd: NULL pointer");
  if (numel == 0) return GH_OK;
  GH_REQUIRE(numel > 0 && numel % 8 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx), GH_ERR_ALIGN,
             "gh_act_bwd: numel must be a multiple of 8 and pointers 16B aligned");
  act_bwd_kernel<<<grid_for(numel / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(dy), static_cast<const bf16*>(x), static_cast<bf16*>(dx), numel / 8, act);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int gh_accum_cast(const float* src, void* dst, int32_t dst_dtype, int64_t numel, float scale,
                             int32_t accumulate, void* stream) {
  GH_REQUIRE(src && dst, GH_ERR_NULL, "gh_accum_cast: NULL pointer");
  if (numel == 0) return GH_OK;
  GH_REQUIRE(numel > 0, GH_ERR_BAD_SHAPE, "gh_accum_cast: negative numel");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dst_dtype == GH_BF16)
    accum_cast_kernel<bf16><<<grid_for(numel, 256), 256, 0, s>>>(src, static_cast<bf16*>(dst), numel, scale, accumulate);
  else if (dst_dtype == GH_F32)
    accum_cast_kernel<float><<<grid_for(numel, 256), 256, 0, s>>>(src, static_cast<float*>(dst), numel, scale, accumulate);
  else
    return set_error(GH_ERR_UNSUPPORTED, "gh_accum_cast: bad dtype");
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}
