// genhancer_b200 -- persistent, warp-specialised tcgen05 GEMM core for sm_100a.
//
//   warp 0      : TMA producer   (one lane issues cp.async.bulk.tensor into a smem ring)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma, accumulators in TMEM)
//   warp 2      : TMEM allocator / deallocator
//   warps 4..7  : epilogue       (tcgen05.ld -> fp32 smem transpose -> fused math ->
//                                 coalesced 8/16-byte global stores)
//
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue,
// two accumulator stages so the epilogue of tile i overlaps the mainloop of tile i+1),
// and a static persistent tile schedule (grid = #SMs, M-grouped rasterisation for L2 reuse).
//
// The same kernel body serves
//   MODE_GEMM : D[M,N] = A * B^T with A/B either K-major or MN-major (fwd, dgrad, wgrad)
//   MODE_CONV : implicit-GEMM convolution over an NHWC activation; the A tile of each
//               (tap, channel-chunk) K-step is a shifted 4-D TMA box whose out-of-bounds
//               part is zero-filled by the hardware (padding costs nothing).
#pragma once
#include "common.cuh"

namespace gh {

enum { MODE_GEMM = 0, MODE_CONV = 1 };

struct EpilogueParams {
  void* d;
  int64_t ldd;
  int d_f32;
  float alpha;
  const void* bias;
  int bias_f32;
  int act;
  int act_grad;
  const __nv_bfloat16* aux_in;
  int64_t ld_aux_in;
  __nv_bfloat16* aux_out;
  int64_t ld_aux_out;
  const __nv_bfloat16* gate;
  int64_t gate_ld;
  int rows_per_batch;
  const void* residual;
  int64_t ld_res;
  int res_f32;
  int vec8;  // every epilogue operand allows 16-byte bf16 vectors (N % 8 == 0, leading dimensions % 8 == 0)
  int atomic;  // fp32 D accumulated with red.global.add (split-K)
  int fast;  // bf16 out = act(alpha*acc + bias) [+ bf16 residual], vec8: the lean path (set by finalize_epilogue)
  int tma;   // lean path through smem tiles + TMA store (tmap_d [, tmap_r] are valid); GEMM mode only
};

inline void finalize_epilogue(EpilogueParams& ep) {
  ep.fast = ep.vec8 && !ep.d_f32 && !ep.aux_out && !ep.act_grad && !ep.gate && (!ep.residual || !ep.res_f32);
}

struct ConvGeom {  // MODE_CONV only
  int B, Ho, Wo;   // output extent
  int TW, TH;      // output patch per M tile (TW*TH <= 128)
  int tiles_w, tiles_h;
  int KW;          // filter width (taps = KH*KW)
  int cin_chunks;  // Cin / 64
  int stride, pad; // input coord = out*stride + tap - pad
};

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  // Second operand pair (LoRA branch folded into the base GEMM): D = A*B^T + A2*B2^T.  The K loop simply runs
  // num_k_blocks2 more 64-wide steps that pull from (tmap_a2, tmap_b2); the last of them issues only
  // k2_last_steps of the four 16-deep MMAs (rank 16 -> one MMA, the rest of the box is TMA zero fill).
  int num_k_blocks2, k2_last_steps;
  // Split-K for skinny outputs (LoRA wgrads: 16..48 x K outputs reduced over all tokens): the tile grid gets a third
  // axis of k_splits slices of the K loop; every slice ADDS its partial result into the fp32 output with red.add
  // (the caller zeroes D).  k_splits = 1: off.  Not combined with the second operand pair.
  int k_splits;
  // Batched GEMM over FLAT 2-D operands (the AE mid-block attention: 32 per-image Q K^T / P V products in one
  // launch): problem b reads A at row (or, MN-major, k) offset b * a_boff of the same tensor map, B at b * b_boff,
  // and writes D rows b * d_brows + m.  Tiles that overrun a problem's M / N read the neighbour's rows (or TMA
  // zero fill at the end of the tensor) and are masked by the epilogue; an overrun along K of a K-major operand is
  // zero-filled because the tensor map's K extent is the per-problem K.  batch = 1: plain GEMM.
  int batch, a_boff, b_boff, d_brows;
  // Dynamic tile schedule (GEMM mode; NULL = the static schedule tile = worker + i * #workers).  tile_ctr[0] hands out
  // tile indices (atomicAdd), tile_ctr[1] counts workers that ran dry; the last one zeroes both for the next launch.
  // For data-parallel training: NCCL's all-reduce kernels hold some SMs while the backward GEMMs run, so a few CTAs
  // of a persistent grid start a whole wave late -- with the static schedule they still owe their full share of
  // tiles and the GEMM takes twice as long; with the dynamic one they find the queue (nearly) empty and leave.
  int* tile_ctr;
  uint32_t a_stage_tx_bytes;  // bytes TMA deposits for the A tile of one stage
  uint32_t mn_lbo, mn_sbo, mn_kstep;  // MN-major descriptor geometry (bytes); see common.cuh
  EpilogueParams ep;
  ConvGeom cv;
  // bring-up aid (gh_debug_gemm_prof): per-CTA cycle counters of the three pipelines, NULL in production.
  //   [0] issuer total  [1] issuer waiting for smem data (TMA starvation)  [2] issuer waiting for a free accumulator
  //   [3] tiles  [4] epilogue warp total  [5] epilogue waiting for an accumulator  [6] producer waiting for a free slot
  long long* prof;
  int dbg;  // bring-up: bit 0 = no TMA (MMAs run on whatever is in smem: isolates the tensor-pipe rate)
};

// CTA2 = the kernel runs as CTA pairs (cluster of 2, tcgen05 cta_group::2): one tile is 256 x BN, each CTA stages
// its own 128 rows of A and BN/2 rows of B (see common.cuh).
template <int BN, bool CTA2 = false>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (CTA2 ? BN / 2 : BN) * BK * 2;   // per CTA
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 8;   // two per TMEM lane quadrant: they take alternate 16-column slices
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int EPI_LD = 20;     // floats per staged row (16 + 4 pad: conflict-free float4 access)
  static constexpr int EPI_BIAS_FLOATS = BN / 2;  // per epilogue warp: fp32 bias of the columns that warp owns
  // per epilogue warp: two 4 KB tiles (32 rows x 128 B, SWIZZLE_128B) that the TMA-store epilogue double-buffers;
  // the register/STG epilogues use the first as fp32 transposition scratch and the second for the staged bias
  static constexpr int EPI_WARP_BYTES = 8192;
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_LIMIT = 227 * 1024;
  static constexpr int STAGES_FIT = (SMEM_LIMIT - 1024 - BAR_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // + align slack
};

__device__ __forceinline__ void decode_tile(int tile, int num_m_blocks, int num_n_blocks, int& m_blk, int& n_blk) {
  constexpr int G = 8;  // M-blocks per raster group
  const int tiles_per_group = G * num_n_blocks;
  const int group = tile / tiles_per_group;
  const int first_m = group * G;
  const int gsize = min(G, num_m_blocks - first_m);
  const int in_group = tile - group * tiles_per_group;
  m_blk = first_m + in_group % gsize;
  n_blk = in_group / gsize;
}

// ----------------------------------------------------------------------------------------------------------
// Fused epilogue of one 32-row x 32-column accumulator chunk held by one warp (fp32, staged row-major in smem).
// Lane l owns the 8 columns col8..col8+7 (col8 = 8*(l&3)) of rows it*8 + (l>>2), it = 0..3: one 16-byte bf16
// vector per row and operand.  All global LOADS of the chunk (bias once; residual / aux_in / gate for the 4 rows)
// are issued up front so that their latencies overlap; only then the math and the stores run.  (A load placed
// after a store cannot be hoisted by the compiler -- possible aliasing -- which exposed one L2 round trip per row
// in the first version of this epilogue: 195 TFLOP/s on the ViT's K=1024 GEMMs.)
// ep.vec8 == 0 (N or a leading dimension not a multiple of 8): 8-byte accesses, upper half masked by `hi`.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld8_bf16(const __nv_bfloat16* p, bool vec8, bool hi) {
  if (vec8) return *reinterpret_cast<const uint4*>(p);
  const uint2 a = *reinterpret_cast<const uint2*>(p);
  uint2 b = make_uint2(0u, 0u);
  if (hi) b = *reinterpret_cast<const uint2*>(p + 4);
  return make_uint4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&v)[8], bool vec8, bool hi) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  if (vec8) {
    *reinterpret_cast<uint4*>(p) = o;
  } else {
    *reinterpret_cast<uint2*>(p) = make_uint2(o.x, o.y);
    if (hi) *reinterpret_cast<uint2*>(p + 4) = make_uint2(o.z, o.w);
  }
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// Operands of one 32-row x 16-column slice, as loaded from global memory (lane = 8 columns of 2 rows).
struct SliceLoads {
  float b8[8];
  uint4 res16[2], aux[2], gt[2];  // (an fp32 residual -- wgrad accumulation -- is loaded at its use instead)
};

__device__ __forceinline__ void slice_issue_loads(const EpilogueParams& ep, const int (&rows)[2], uint32_t okmask,
                                                  bool hi, int gn, SliceLoads& L) {
  const bool vec8 = ep.vec8 != 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) L.b8[k] = 0.f;
  if (ep.bias && okmask) {
    if (ep.bias_f32) {
      const float* bp = static_cast<const float*>(ep.bias) + gn;
      const float4 x = __ldg(reinterpret_cast<const float4*>(bp));
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hi) y = __ldg(reinterpret_cast<const float4*>(bp + 4));
      L.b8[0] = x.x; L.b8[1] = x.y; L.b8[2] = x.z; L.b8[3] = x.w;
      L.b8[4] = y.x; L.b8[5] = y.y; L.b8[6] = y.z; L.b8[7] = y.w;
    } else {
      unpack8(ld8_bf16(static_cast<const __nv_bfloat16*>(ep.bias) + gn, vec8, hi), L.b8);
    }
  }
  if (ep.residual && !ep.res_f32) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (okmask >> i & 1u)
        L.res16[i] = ld8_bf16(static_cast<const __nv_bfloat16*>(ep.residual) +
                              static_cast<int64_t>(rows[i]) * ep.ld_res + gn, vec8, hi);
  }
  if (ep.act_grad) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (okmask >> i & 1u) L.aux[i] = ld8_bf16(ep.aux_in + static_cast<int64_t>(rows[i]) * ep.ld_aux_in + gn, vec8, hi);
  }
  if (ep.gate) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (okmask >> i & 1u)
        L.gt[i] = ld8_bf16(ep.gate + static_cast<int64_t>(rows[i] / ep.rows_per_batch) * ep.gate_ld + gn, vec8, hi);
  }
}

__device__ __forceinline__ void slice_finish(const EpilogueParams& ep, const float* __restrict__ st, int ld_st,
                                             const int (&rows)[2], uint32_t okmask, bool hi, int sub_row, int col8,
                                             int gn, const SliceLoads& L) {
  const bool vec8 = ep.vec8 != 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (!(okmask >> i & 1u)) continue;
    const int64_t row = rows[i];
    const float* sp = st + (i * 16 + sub_row) * ld_st + col8;
    const float4 a0 = *reinterpret_cast<const float4*>(sp), a1 = *reinterpret_cast<const float4*>(sp + 4);
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = v[k] * ep.alpha + L.b8[k];
    if (ep.aux_out) st8_bf16(ep.aux_out + row * ep.ld_aux_out + gn, v, vec8, hi);
    if (ep.act_grad) {
      float x[8];
      unpack8(L.aux[i], x);
      act_bwd_mul8(ep.act, x, v);
    } else if (ep.act != ACT_NONE) {
      act_fwd8(ep.act, v);
    }
    if (ep.gate) {
      float g[8];
      unpack8(L.gt[i], g);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= g[k];
    }
    if (ep.residual) {
      if (ep.res_f32) {
        const float* rp = static_cast<const float*>(ep.residual) + row * ep.ld_res + gn;
        const float4 f0 = *reinterpret_cast<const float4*>(rp);
        const float4 f1 = hi ? *reinterpret_cast<const float4*>(rp + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[0] += f0.x; v[1] += f0.y; v[2] += f0.z; v[3] += f0.w;
        v[4] += f1.x; v[5] += f1.y; v[6] += f1.z; v[7] += f1.w;
      } else {
        float r[8];
        unpack8(L.res16[i], r);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += r[k];
      }
    }
    if (ep.d_f32 && ep.atomic) {
      float* dp = static_cast<float*>(ep.d) + row * ep.ldd + gn;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < 4 || hi) atomicAdd(dp + k, v[k]);
    } else if (ep.d_f32) {
      float* dp = static_cast<float*>(ep.d) + row * ep.ldd + gn;
      *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
      if (hi) *reinterpret_cast<float4*>(dp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      st8_bf16(static_cast<__nv_bfloat16*>(ep.d) + row * ep.ldd + gn, v, vec8, hi);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// Lean epilogue of one warp for one full-width tile:  bf16 out = act(alpha*acc + bias) [+ bf16 residual].
// Compile-time activation / residual so the slice loop has no option branches; TMEM -> registers runs one slice
// ahead of the math; the residual of slice j+1 is requested before slice j is computed.
// ----------------------------------------------------------------------------------------------------------
template <int BN, int ACT, bool HAS_RES>
__device__ __forceinline__ void lean_tile(const EpilogueParams& ep, float* __restrict__ st, const float* __restrict__ bias_s,
                                          uint32_t t_row, int half, int lane, int sub_row, int col8, int n0,
                                          const int (&rows)[2], uint32_t rowmask) {
  constexpr int LD = GemmCfg<BN>::EPI_LD;
  constexpr int NJ = BN / 32;  // slices per warp
  const bool ok0 = rowmask & 1u, ok1 = (rowmask >> 1) & 1u;
  __nv_bfloat16* d0 = static_cast<__nv_bfloat16*>(ep.d) + static_cast<int64_t>(rows[0]) * ep.ldd + n0 + col8 + half * 16;
  __nv_bfloat16* d1 = static_cast<__nv_bfloat16*>(ep.d) + static_cast<int64_t>(rows[1]) * ep.ldd + n0 + col8 + half * 16;
  const __nv_bfloat16* r0 = nullptr;
  const __nv_bfloat16* r1 = nullptr;
  uint4 res0 = make_uint4(0, 0, 0, 0), res1 = res0;
  if (HAS_RES) {
    r0 = static_cast<const __nv_bfloat16*>(ep.residual) + static_cast<int64_t>(rows[0]) * ep.ld_res + n0 + col8 + half * 16;
    r1 = static_cast<const __nv_bfloat16*>(ep.residual) + static_cast<int64_t>(rows[1]) * ep.ld_res + n0 + col8 + half * 16;
    if (ok0) res0 = *reinterpret_cast<const uint4*>(r0);
    if (ok1) res1 = *reinterpret_cast<const uint4*>(r1);
  }
  const float alpha = ep.alpha;
  const float* sp0 = st + sub_row * LD + col8;
  const float* sp1 = sp0 + 16 * LD;
  float4* dst = reinterpret_cast<float4*>(st + lane * LD);
  uint32_t treg[16];
  tmem_ld_32x16(t_row + half * 16, treg);
  tmem_ld_wait();
#pragma unroll 1
  for (int j = 0; j < NJ; ++j) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_float4(__uint_as_float(treg[4 * q]), __uint_as_float(treg[4 * q + 1]),
                           __uint_as_float(treg[4 * q + 2]), __uint_as_float(treg[4 * q + 3]));
    __syncwarp();
    if (j + 1 < NJ) tmem_ld_32x16(t_row + half * 16 + (j + 1) * 32, treg);
    const float4 bb0 = *reinterpret_cast<const float4*>(bias_s + j * 16 + col8);
    const float4 bb1 = *reinterpret_cast<const float4*>(bias_s + j * 16 + col8 + 4);
    float v0[8], v1[8];
    {
      const float4 a0 = *reinterpret_cast<const float4*>(sp0), a1 = *reinterpret_cast<const float4*>(sp0 + 4);
      const float4 c0 = *reinterpret_cast<const float4*>(sp1), c1 = *reinterpret_cast<const float4*>(sp1 + 4);
      v0[0] = fmaf(a0.x, alpha, bb0.x); v0[1] = fmaf(a0.y, alpha, bb0.y); v0[2] = fmaf(a0.z, alpha, bb0.z); v0[3] = fmaf(a0.w, alpha, bb0.w);
      v0[4] = fmaf(a1.x, alpha, bb1.x); v0[5] = fmaf(a1.y, alpha, bb1.y); v0[6] = fmaf(a1.z, alpha, bb1.z); v0[7] = fmaf(a1.w, alpha, bb1.w);
      v1[0] = fmaf(c0.x, alpha, bb0.x); v1[1] = fmaf(c0.y, alpha, bb0.y); v1[2] = fmaf(c0.z, alpha, bb0.z); v1[3] = fmaf(c0.w, alpha, bb0.w);
      v1[4] = fmaf(c1.x, alpha, bb1.x); v1[5] = fmaf(c1.y, alpha, bb1.y); v1[6] = fmaf(c1.z, alpha, bb1.z); v1[7] = fmaf(c1.w, alpha, bb1.w);
    }
    if (ACT != ACT_NONE) {
      act_fwd8(ACT, v0);
      act_fwd8(ACT, v1);
    }
    if (HAS_RES) {
      float r[8];
      unpack8(res0, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v0[k] += r[k];
      unpack8(res1, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v1[k] += r[k];
      if (j + 1 < NJ) {  // next slice's residual: in flight during the stores and the next staging
        if (ok0) res0 = *reinterpret_cast<const uint4*>(r0 + (j + 1) * 32);
        if (ok1) res1 = *reinterpret_cast<const uint4*>(r1 + (j + 1) * 32);
      }
    }
    if (ok0) {
      uint4 o;
      o.x = pack_bf16x2(v0[0], v0[1]); o.y = pack_bf16x2(v0[2], v0[3]);
      o.z = pack_bf16x2(v0[4], v0[5]); o.w = pack_bf16x2(v0[6], v0[7]);
      *reinterpret_cast<uint4*>(d0 + j * 32) = o;
    }
    if (ok1) {
      uint4 o;
      o.x = pack_bf16x2(v1[0], v1[1]); o.y = pack_bf16x2(v1[2], v1[3]);
      o.z = pack_bf16x2(v1[4], v1[5]); o.w = pack_bf16x2(v1[6], v1[7]);
      *reinterpret_cast<uint4*>(d1 + j * 32) = o;
    }
    __syncwarp();
    if (j + 1 < NJ) tmem_ld_wait();
  }
}

// ----------------------------------------------------------------------------------------------------------
// Lean epilogue through TMA:  bf16 out = act(alpha*acc + bias) [+ bf16 residual].
// tcgen05.ld hands every lane one ROW of the accumulator (16 consecutive columns per load), so no transposition is
// needed: the lane converts its 16 values, writes two 16-byte chunks into a [32 rows x 128 B] SWIZZLE_128B tile
// (conflict-free: chunk index ^ (row & 7)), and one elected lane stores the tile with ONE bulk tensor store --
// full 128-byte lines, clipped by the hardware at M and N.  The residual tile arrives the same way (TMA load into
// the same buffer, read-modify-write in place).  Replaces ~32 partial-line STG wavefronts per 16-column slice.
// Warp (quad, half) owns rows 32*quad.. and the 64-column groups half, half+2, ...
// ----------------------------------------------------------------------------------------------------------
// RM = 0: no second tile; 1: + bf16 residual; 2: act_grad -- the tile is aux_in (the saved pre-activation) and the
// result is (alpha*acc + bias) * ACT'(aux_in): the dgrad through a GELU / QuickGELU MLP (fc1's dY) on the lean path.
// Request the second tile (residual / saved pre-activation) of 64-column group `g` into staging buffer `b`.  The bulk
// store that last read that buffer must be done with it: `pending_ok` = how many of this thread's committed store groups
// may still be reading (1 when the most recent store used the OTHER buffer, 0 when it used this one).
template <int PENDING_OK>
__device__ __forceinline__ void lean_res_request(const CUtensorMap* tmap_r, uint8_t* stg, uint64_t* res_bar2, int b, int gn, int row0) {
  if (elect_one()) {
    bulk_wait_group_read<PENDING_OK>();
    mbar_arrive_expect_tx(&res_bar2[b], 4096u);
    tma_load_2d(stg + b * 4096, tmap_r, &res_bar2[b], gn, row0);
  }
  __syncwarp();
}

// RM != 0: the second tile of a group is requested BEFORE the group is worked on -- the first group of an output tile by
// the caller, ahead of its wait for the accumulator (a whole mainloop of cover), every later group at the start of the
// group before it -- so its L2 / DRAM latency is off the epilogue's critical path (it used to be requested at the top
// of its own group and waited for ~100 instructions later: with K = 1024 the epilogue paces the kernel).
template <int BN, int ACT, int RM>
__device__ __forceinline__ void lean_tile_tma(const EpilogueParams& ep, const CUtensorMap* tmap_d, const CUtensorMap* tmap_r,
                                              uint8_t* stg, uint64_t* res_bar2, uint32_t& res_phase, int& buf,
                                              uint32_t t_row, int half, int lane, int row0, int n0, int N,
                                              const float4 (&breg)[(BN + 255) / 256]) {
  constexpr int NG = BN / 64;
  const float alpha = ep.alpha;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
#pragma unroll 1
  for (int g = half; g < NG; g += 2) {
    const int gn = n0 + g * 64;
    if (gn >= N) break;
    uint8_t* tile = stg + buf * 4096;
    uint8_t* rowp = tile + lane * 128;
    if (RM == 1 || RM == 2) {
      // next group's second tile -> the other buffer (last read by the store committed just before this group)
      if (g + 2 < NG && gn + 128 < N) lean_res_request<0>(tmap_r, stg, res_bar2, buf ^ 1, gn + 128, row0);
    } else if (RM == 3) {
      // RM == 3 (aux_out): BOTH buffers are written below -- the activated tile and the pre-activation tile an MLP's
      // backward needs -- and leave by two bulk stores; the pair committed for the previous group must be done reading
      if (elect_one()) bulk_wait_group_read<0>();
      __syncwarp();
    } else {
      // the bulk store that last read this buffer must be done with it (one other group may still be in flight)
      if (elect_one()) bulk_wait_group_read<1>();
      __syncwarp();
    }
    uint32_t treg[16];
    tmem_ld_32x16(t_row + g * 64, treg);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // bias of these 16 columns: held 4 per lane by the lanes (gi % 2) * 16 + 4c .. + 3 of register set gi / 2
      float b[16];
      {
        const int gi = (g - half) >> 1;
        const float4 src = breg[gi >> 1];
        const int l0 = (gi & 1) * 16 + c * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          b[4 * q] = __shfl_sync(0xffffffffu, src.x, l0 + q);
          b[4 * q + 1] = __shfl_sync(0xffffffffu, src.y, l0 + q);
          b[4 * q + 2] = __shfl_sync(0xffffffffu, src.z, l0 + q);
          b[4 * q + 3] = __shfl_sync(0xffffffffu, src.w, l0 + q);
        }
      }
      tmem_ld_wait();
      float v0[8], v1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v0[k] = fmaf(__uint_as_float(treg[k]), alpha, b[k]);
        v1[k] = fmaf(__uint_as_float(treg[8 + k]), alpha, b[8 + k]);
      }
      if (c + 1 < 4) tmem_ld_32x16(t_row + g * 64 + (c + 1) * 16, treg);   // next chunk in flight during the math
      if (RM == 3) {
        uint8_t* rowa = stg + (buf ^ 1) * 4096 + lane * 128;
        uint4 o;
        o.x = pack_bf16x2(v0[0], v0[1]); o.y = pack_bf16x2(v0[2], v0[3]);
        o.z = pack_bf16x2(v0[4], v0[5]); o.w = pack_bf16x2(v0[6], v0[7]);
        *reinterpret_cast<uint4*>(rowa + (((2 * c) ^ sw) << 4)) = o;
        o.x = pack_bf16x2(v1[0], v1[1]); o.y = pack_bf16x2(v1[2], v1[3]);
        o.z = pack_bf16x2(v1[4], v1[5]); o.w = pack_bf16x2(v1[6], v1[7]);
        *reinterpret_cast<uint4*>(rowa + (((2 * c + 1) ^ sw) << 4)) = o;
      }
      if (ACT != ACT_NONE && RM != 2) {
        act_fwd8(ACT, v0);
        act_fwd8(ACT, v1);
      }
      uint4* p0 = reinterpret_cast<uint4*>(rowp + (((2 * c) ^ sw) << 4));
      uint4* p1 = reinterpret_cast<uint4*>(rowp + (((2 * c + 1) ^ sw) << 4));
      if (RM == 1 || RM == 2) {
        if (c == 0) {
          mbar_wait(&res_bar2[buf], (res_phase >> buf) & 1u);
          res_phase ^= 1u << buf;
        }
        float r[8];
        unpack8(*p0, r);
        if (RM == 2) {
          act_bwd_mul8(ACT, r, v0);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) v0[k] += r[k];
        }
        unpack8(*p1, r);
        if (RM == 2) {
          act_bwd_mul8(ACT, r, v1);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) v1[k] += r[k];
        }
      }
      uint4 o;
      o.x = pack_bf16x2(v0[0], v0[1]); o.y = pack_bf16x2(v0[2], v0[3]);
      o.z = pack_bf16x2(v0[4], v0[5]); o.w = pack_bf16x2(v0[6], v0[7]);
      *p0 = o;
      o.x = pack_bf16x2(v1[0], v1[1]); o.y = pack_bf16x2(v1[2], v1[3]);
      o.z = pack_bf16x2(v1[4], v1[5]); o.w = pack_bf16x2(v1[6], v1[7]);
      *p1 = o;
    }
    fence_proxy_async_smem();   // generic-proxy writes of the tile -> visible to the bulk store
    __syncwarp();
    if (elect_one()) {
      tma_store_2d(tmap_d, tile, gn, row0);
      if (RM == 3) tma_store_2d(tmap_r, stg + (buf ^ 1) * 4096, gn, row0);
      bulk_commit_group();
    }
    buf ^= 1;
  }
}

template <int BN, bool A_MN, bool B_MN, int MODE, bool CTA2 = false>
__global__ void __launch_bounds__(GemmCfg<BN, CTA2>::THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b2,
                 const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, CTA2>;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(!(MODE == MODE_CONV && A_MN), "conv A operand is K-major (NHWC channels)");
  static_assert(!CTA2 || BN >= 128, "a CTA pair splits the B tile: BN/2 must be a multiple of 64");
  // CTA pair: rank 0 (leader) issues the MMAs for both; tiles are indexed per PAIR (256 rows)
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int worker = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int m_units = CTA2 ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks;   // rows of the tile grid

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment of the (shared-window) address
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* stage_base = smem;
  float* epi_buf = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
  uint64_t* res_bars = bars + 2 * STAGES + 4;   // [2 * EPI_WARPS] residual tiles of the TMA epilogue (one per staging buffer)
  uint64_t* tq_full = bars + 2 * STAGES + 4 + 2 * Cfg::EPI_WARPS;   // [2] dynamic tile queue: entry published
  uint64_t* tq_empty = tq_full + 2;                             // [2] entry read by every consumer warp (leader's copy)
  int* tile_q = reinterpret_cast<int*>(tq_empty + 2);           // [2] tile index or -1 (no more tiles)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tile_q + 2);
  static_assert((2 * STAGES + 4 + 2 * Cfg::EPI_WARPS + 4) * 8 + 8 + 4 <= Cfg::BAR_BYTES, "barrier block too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_mn = m_units * p.num_n_blocks;
  const int tiles_pb = tiles_mn * p.k_splits;    // tiles of one problem: index = split * tiles_mn + (m, n) index
  const int num_tiles = tiles_pb * (MODE == MODE_GEMM ? p.batch : 1);   // batch-major
  const int nkb1 = p.num_k_blocks;
  const int nkb = nkb1 + (MODE == MODE_GEMM ? p.num_k_blocks2 : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (MODE == MODE_GEMM && p.num_k_blocks2 > 0) {
      tma_prefetch_desc(&tmap_a2);
      tma_prefetch_desc(&tmap_b2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], Cfg::EPI_WARPS * (CTA2 ? 2 : 1));  // one arrive per epilogue warp (of both CTAs)
    mbar_init(&tempty_bar[1], Cfg::EPI_WARPS * (CTA2 ? 2 : 1));
    for (int w = 0; w < 2 * Cfg::EPI_WARPS; ++w) mbar_init(&res_bars[w], 1);
    // consumers of a tile-queue entry: producer + issuer + epilogue warps of the leader, producer + epilogue warps of the peer
    mbar_init(&tq_full[0], 1);
    mbar_init(&tq_full[1], 1);
    mbar_init(&tq_empty[0], CTA2 ? 2 * Cfg::EPI_WARPS + 3 : Cfg::EPI_WARPS + 2);
    mbar_init(&tq_empty[1], CTA2 ? 2 * Cfg::EPI_WARPS + 3 : Cfg::EPI_WARPS + 2);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CTA2) tmem_alloc2<Cfg::TMEM_COLS>(tmem_ptr);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // the it-th tile of this worker (CTA or CTA pair), or -1: static stride schedule, or the next entry of the dynamic
  // queue that warp 3 of the (leader) CTA fills from the global counter
#ifdef GH_NO_DYNAMIC_TILES   // A/B builds: compile the dynamic schedule out
  constexpr bool dyn = false;
#else
  const bool dyn = MODE == MODE_GEMM && p.tile_ctr != nullptr;
#endif
  const uint32_t tq_empty_leader[2] = {CTA2 ? mapa_u32(smem_u32(&tq_empty[0]), 0u) : 0u,
                                       CTA2 ? mapa_u32(smem_u32(&tq_empty[1]), 0u) : 0u};
  auto fetch_tile = [&](int it) -> int {
    if (!dyn) {
      const int t = worker + it * num_workers;
      return t < num_tiles ? t : -1;
    }
    const int slot = it & 1;
    const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
    if (CTA2) mbar_wait_cluster(&tq_full[slot], ph);
    else mbar_wait(&tq_full[slot], ph);
    const int t = *reinterpret_cast<volatile int*>(&tile_q[slot]);
    __syncwarp();
    if (lane == 0) {
      if (CTA2) mbar_arrive_cluster(tq_empty_leader[slot]);
      else mbar_arrive(&tq_empty[slot]);
    }
    return t;
  };

  if (warp == 3 && rank == 0 && dyn) {
    // ===================== tile scheduler (dynamic mode) =====================
    for (int it = 0;; ++it) {
      const int slot = it & 1;
      const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
      if (CTA2) mbar_wait_cluster(&tq_empty[slot], ph ^ 1u);   // every consumer has read the entry of round it - 2
      else mbar_wait(&tq_empty[slot], ph ^ 1u);
      int t = 0;
      if (lane == 0) {
        t = atomicAdd(p.tile_ctr, 1);
        if (t >= num_tiles) t = -1;
        *reinterpret_cast<volatile int*>(&tile_q[slot]) = t;
        if (CTA2) {
          st_shared_cluster_u32(mapa_u32(smem_u32(&tile_q[slot]), 1u), static_cast<uint32_t>(t));
          mbar_arrive_cluster(mapa_u32(smem_u32(&tq_full[slot]), 1u));   // release.cluster: orders the store above
        }
        mbar_arrive(&tq_full[slot]);
      }
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t < 0) {
        if (lane == 0) {
          const int done = atomicAdd(p.tile_ctr + 1, 1);
          if (done == num_workers - 1) {   // every worker has made its last fetch: hand the counters back zeroed
            p.tile_ctr[0] = 0;
            p.tile_ctr[1] = 0;
          }
        }
        break;
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    long long w_slot = 0;
    const uint32_t tx_bytes = (p.a_stage_tx_bytes + Cfg::B_BYTES) * (CTA2 ? 2u : 1u);
    constexpr int BNL = CTA2 ? BN / 2 : BN;            // B rows this CTA loads
    for (int it = 0;; ++it) {
      const int tile = fetch_tile(it);
      if (tile < 0) break;
      int m_blk, n_blk;
      const int bi = (MODE == MODE_GEMM && p.batch > 1) ? tile / tiles_pb : 0;
      const int ptile = tile - bi * tiles_pb;
      const int split = ptile / tiles_mn;
      decode_tile(ptile - split * tiles_mn, m_units, p.num_n_blocks, m_blk, n_blk);
      const int kb0 = (split * nkb) / p.k_splits, kb1 = ((split + 1) * nkb) / p.k_splits;
      if (CTA2) m_blk = m_blk * 2 + static_cast<int>(rank);
      const int nrow0 = n_blk * BN + static_cast<int>(rank) * BNL;
      const int a_off = bi * p.a_boff, b_off = bi * p.b_boff;
      int cb = 0, ch0 = 0, cw0 = 0;
      if (MODE == MODE_CONV) {
        const int tiles_per_img = p.cv.tiles_w * p.cv.tiles_h;
        cb = m_blk / tiles_per_img;
        const int r = m_blk - cb * tiles_per_img;
        ch0 = (r / p.cv.tiles_w) * p.cv.TH;
        cw0 = (r % p.cv.tiles_w) * p.cv.TW;
      }
      int c_cc = 0, c_kw = 0, c_kh = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (p.prof) {
          const long long t0 = clock64();
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          w_slot += clock64() - t0;
        } else {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
        }
        if (elect_one()) {
          uint8_t* sA = stage_base + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          // CTA pair: both CTAs' loads count on the LEADER's full barrier (it expects the bytes of both)
          const uint32_t fb = CTA2 ? mapa_u32(smem_u32(&full_bar[stage]), 0u) : 0u;
          if (p.dbg & 1) {
            if (!CTA2 || rank == 0) mbar_arrive(&full_bar[stage]);
            goto next_stage;
          }
          if (!CTA2 || rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          auto ld2 = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if (CTA2) tma2_load_2d(dst, m, fb, c0, c1);
            else tma_load_2d(dst, m, &full_bar[stage], c0, c1);
          };
          if (MODE == MODE_CONV) {
            const int cc = c_cc, kh = c_kh, kw = c_kw;   // (tap, 64-channel chunk) of this k block, kept incrementally
            if (CTA2)
              tma2_load_4d(sA, &tmap_a, fb, cc * 64, cw0 * p.cv.stride + kw - p.cv.pad, ch0 * p.cv.stride + kh - p.cv.pad, cb);
            else
              tma_load_4d(sA, &tmap_a, &full_bar[stage], cc * 64, cw0 * p.cv.stride + kw - p.cv.pad,
                          ch0 * p.cv.stride + kh - p.cv.pad, cb);
          } else {
            const bool second = kb >= nkb1;
            const CUtensorMap* ma = second ? &tmap_a2 : &tmap_a;
            const int kc = (second ? kb - nkb1 : kb) * 64;
            if (!A_MN) {
              ld2(sA, ma, kc, m_blk * 128 + a_off);
            } else {
              ld2(sA, ma, m_blk * 128, kc + a_off);
              ld2(sA + 8192, ma, m_blk * 128 + 64, kc + a_off);
            }
          }
          if (MODE == MODE_CONV) {
            ld2(sB, &tmap_b, kb * 64, nrow0);
          } else {
            const bool second = kb >= nkb1;
            const CUtensorMap* mb = second ? &tmap_b2 : &tmap_b;
            const int kc = (second ? kb - nkb1 : kb) * 64;
            if (!B_MN) {
              ld2(sB, mb, kc, nrow0 + b_off);
            } else {
#pragma unroll
              for (int i = 0; i < BNL / 64; ++i) ld2(sB + i * 8192, mb, nrow0 + i * 64, kc + b_off);
            }
          }
        }
      next_stage:
        __syncwarp();
        if (MODE == MODE_CONV) {
          if (++c_cc == p.cv.cin_chunks) {
            c_cc = 0;
            if (++c_kw == p.cv.KW) { c_kw = 0; ++c_kh; }
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (p.prof && lane == 0) p.prof[blockIdx.x * 8 + 6] = w_slot;
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA of a pair) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(CTA2 ? 256 : 128, BN, A_MN, B_MN);
    const uint64_t a_desc_base = A_MN ? umma_desc_base(p.mn_lbo, p.mn_sbo) : umma_desc_base(16u, 1024u);
    const uint64_t b_desc_base = B_MN ? umma_desc_base(p.mn_lbo, p.mn_sbo) : umma_desc_base(16u, 1024u);
    const uint32_t A_KSTEP = A_MN ? p.mn_kstep : 32u;
    const uint32_t B_KSTEP = B_MN ? p.mn_kstep : 32u;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_full = 0, w_acc = 0, n_tiles = 0;
    const long long t_begin = p.prof ? clock64() : 0;
    for (int it = 0;; ++it) {
      const int tile = fetch_tile(it);
      if (tile < 0) break;
      if (p.prof) {
        const long long t0 = clock64();
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        w_acc += clock64() - t0;
        ++n_tiles;
      } else {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      const int split = (tile % tiles_pb) / tiles_mn;
      const int kb0 = (split * nkb) / p.k_splits, kb1 = ((split + 1) * nkb) / p.k_splits;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (p.prof) {
          const long long t0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          w_full += clock64() - t0;
        } else {
          mbar_wait(&full_bar[stage], phase);
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + Cfg::A_BYTES;
          auto mma = [&](int k) {
            if (CTA2)
              umma2_ss(d_tmem, umma_desc_at(a_desc_base, sA + k * A_KSTEP), umma_desc_at(b_desc_base, sB + k * B_KSTEP),
                       idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            else
              umma_ss(d_tmem, umma_desc_at(a_desc_base, sA + k * A_KSTEP), umma_desc_at(b_desc_base, sB + k * B_KSTEP),
                      idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          };
          if (kb == nkb - 1 && kb >= nkb1) {  // last k block of a folded LoRA pair: only the steps that hold data
            for (int k = 0; k < p.k2_last_steps; ++k) mma(k);
          } else {
            mma(0); mma(1); mma(2); mma(3);
          }
          if (CTA2) {
            umma2_commit(&empty_bar[stage]);                   // the slot is free in BOTH CTAs once these MMAs retire
            if (kb == kb1 - 1) umma2_commit(&tfull_bar[acc]);  // accumulator complete (both CTAs' epilogues)
          } else {
            umma_commit(&empty_bar[stage]);              // smem slot free once these MMAs retire
            if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (p.prof && lane == 0) {
      p.prof[blockIdx.x * 8 + 0] = clock64() - t_begin;
      p.prof[blockIdx.x * 8 + 1] = w_full;
      p.prof[blockIdx.x * 8 + 2] = w_acc;
      p.prof[blockIdx.x * 8 + 3] = n_tiles;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // 8 warps: warp (4 + q) and (8 + q) both own TMEM lane quadrant q (rows 32q..32q+31 of the tile) and take
    // alternate 16-column slices.  Per slice: tcgen05.ld -> fp32 staging in smem (row-major) -> each lane handles
    // 8 consecutive columns of 2 rows with 16-byte global accesses.  The global operands of slice s+2 are
    // requested before slice s is processed (software prefetch), so L2 latency overlaps the math.
    const int ew = warp - 4;
    const int quad = ew & 3;
    const int half = ew >> 2;
    uint8_t* stg = reinterpret_cast<uint8_t*>(epi_buf) + ew * Cfg::EPI_WARP_BYTES;   // 1024-byte aligned
    float* st = reinterpret_cast<float*>(stg);
    float* bias_s = reinterpret_cast<float*>(stg + 4096);
    uint32_t res_phase = 0;
    int stg_buf = 0;   // which of the two tiles the next 64-column group uses (alternates across tiles too)
    const bool tma_epi = MODE == MODE_GEMM && p.ep.tma != 0;
    if (tma_epi && lane == 0) {
      tma_prefetch_desc(&tmap_d);
      if (p.ep.residual || p.ep.act_grad) tma_prefetch_desc(&tmap_r);
    }
    const int sub_row = lane & 15;        // 8 consecutive lanes read 8 consecutive staged rows: conflict-free
    const int col8 = (lane >> 4) * 8;
    const EpilogueParams& ep = p.ep;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_tfull = 0, w_lean = 0;
    const long long e_begin = p.prof ? clock64() : 0;
    const uint32_t tempty_leader[2] = {CTA2 ? mapa_u32(smem_u32(&tempty_bar[0]), 0u) : 0u,
                                       CTA2 ? mapa_u32(smem_u32(&tempty_bar[1]), 0u) : 0u};
    for (int it = 0;; ++it) {
      const int tile = fetch_tile(it);
      if (tile < 0) break;
      int m_blk, n_blk;
      decode_tile(tile % tiles_mn, m_units, p.num_n_blocks, m_blk, n_blk);
      if (CTA2) m_blk = m_blk * 2 + static_cast<int>(rank);
      const int d_row_off = (MODE == MODE_GEMM && p.batch > 1) ? (tile / tiles_pb) * p.d_brows : 0;
      // the 2 output rows this lane touches in every slice of the tile (row index == logical GEMM row)
      int rows[2];
      uint32_t rowmask = 0;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int tr = quad * 32 + it * 16 + sub_row;  // row within the 128-row tile
        bool ok;
        if (MODE == MODE_CONV) {
          const int tiles_per_img = p.cv.tiles_w * p.cv.tiles_h;
          const int cb = m_blk / tiles_per_img;
          const int rem = m_blk - cb * tiles_per_img;
          const int oh = (rem / p.cv.tiles_w) * p.cv.TH + tr / p.cv.TW;
          const int ow = (rem % p.cv.tiles_w) * p.cv.TW + tr % p.cv.TW;
          ok = (tr < p.cv.TW * p.cv.TH) && (oh < p.cv.Ho) && (ow < p.cv.Wo) && (cb < p.cv.B);
          rows[it] = (cb * p.cv.Ho + oh) * p.cv.Wo + ow;
        } else {
          ok = m_blk * 128 + tr < p.M;
          rows[it] = m_blk * 128 + tr + d_row_off;   // row of the flat output (batched: problem b starts at b * d_brows)
        }
        rowmask |= static_cast<uint32_t>(ok) << it;
      }
      const int n0 = n_blk * BN;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
      if (tma_epi) {
        // ---------------- lean path through smem tiles + TMA store (any tile, clipped at M / N by the hardware) ----------
        // this warp's bias slice, 4 columns per lane, requested while the accumulator is still being produced:
        // lane l of register set s holds columns n0 + 64 * (half + 2 * (2 s + l / 16)) + 4 * (l % 16) ..
        float4 breg[(BN + 255) / 256];
#pragma unroll
        for (int sidx = 0; sidx < (BN + 255) / 256; ++sidx) {
          const int col = n0 + 64 * (half + 2 * (2 * sidx + (lane >> 4))) + 4 * (lane & 15);
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias && col < p.N && col < n0 + BN) {
            if (ep.bias_f32) {
              bv = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(ep.bias) + col));
            } else {
              const uint2 u = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(ep.bias) + col));
              const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
              bv = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
          }
          breg[sidx] = bv;
        }
        const int row0 = m_blk * 128 + quad * 32;
        // second tile (residual / pre-activation) of this warp's FIRST group: requested before the accumulator is ready;
        // the buffer it lands in was last read by the store before the most recent one (at most one may still be reading)
        if ((ep.residual || ep.act_grad) && row0 < p.M && half < BN / 64 && n0 + half * 64 < p.N)   // (iff the loop below runs)
          lean_res_request<1>(&tmap_r, stg, &res_bars[2 * ew], stg_buf, n0 + half * 64, row0);
        if (p.prof) {
          const long long t0 = clock64();
          mbar_wait(&tfull_bar[acc], acc_phase);
          w_tfull += clock64() - t0;
        } else {
          mbar_wait(&tfull_bar[acc], acc_phase);
        }
        tc_fence_after();
        const long long t_lean0 = p.prof ? clock64() : 0;
#define GH_LEANT(A, R) \
  lean_tile_tma<BN, A, R>(ep, &tmap_d, &tmap_r, stg, &res_bars[2 * ew], res_phase, stg_buf, t_row, half, lane, row0, n0, p.N, breg)
        if (row0 < p.M) {
          if (ep.act_grad) {   // (the host only routes GELU-tanh and QuickGELU here: the DiT's and the ViT's MLPs)
            if (ep.act == ACT_GELU_TANH) GH_LEANT(ACT_GELU_TANH, 2);
            else GH_LEANT(ACT_QUICK_GELU, 2);
          } else if (ep.aux_out) {   // forward of an MLP's fc1: act(pre) AND pre leave through the lean path
            switch (ep.act) {
              case ACT_GELU_TANH: GH_LEANT(ACT_GELU_TANH, 3); break;
              case ACT_QUICK_GELU: GH_LEANT(ACT_QUICK_GELU, 3); break;
              case ACT_SILU: GH_LEANT(ACT_SILU, 3); break;
              default: GH_LEANT(ACT_GELU_ERF, 3); break;
            }
          } else if (ep.residual) {
            switch (ep.act) {
              case ACT_GELU_TANH: GH_LEANT(ACT_GELU_TANH, 1); break;
              case ACT_QUICK_GELU: GH_LEANT(ACT_QUICK_GELU, 1); break;
              case ACT_SILU: GH_LEANT(ACT_SILU, 1); break;
              case ACT_GELU_ERF: GH_LEANT(ACT_GELU_ERF, 1); break;
              default: GH_LEANT(ACT_NONE, 1); break;
            }
          } else {
            switch (ep.act) {
              case ACT_GELU_TANH: GH_LEANT(ACT_GELU_TANH, 0); break;
              case ACT_QUICK_GELU: GH_LEANT(ACT_QUICK_GELU, 0); break;
              case ACT_SILU: GH_LEANT(ACT_SILU, 0); break;
              case ACT_GELU_ERF: GH_LEANT(ACT_GELU_ERF, 0); break;
              default: GH_LEANT(ACT_NONE, 0); break;
            }
          }
        }
#undef GH_LEANT
        if (p.prof) w_lean += clock64() - t_lean0;
      } else if (ep.fast && n0 + BN <= p.N) {
        // ---------------- lean path: bf16 out = act(alpha*acc + bias) [+ bf16 residual], full-width tile ----------------
        // bias of this warp's columns -> smem (fp32) while the accumulator is still being produced
        {  // all loads first, then the smem stores (a store between two loads pins their order: 4 L2 round trips)
          constexpr int NB = Cfg::EPI_BIAS_FLOATS / 32;
          float bv[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int idx = lane + 32 * i;
            const int col = n0 + (half + 2 * (idx >> 4)) * 16 + (idx & 15);
            bv[i] = 0.f;
            if (ep.bias)
              bv[i] = ep.bias_f32 ? __ldg(static_cast<const float*>(ep.bias) + col)
                                  : __bfloat162float(static_cast<const __nv_bfloat16*>(ep.bias)[col]);
          }
#pragma unroll
          for (int i = 0; i < NB; ++i) bias_s[lane + 32 * i] = bv[i];
        }
        __syncwarp();
        if (p.prof) {
          const long long t0 = clock64();
          mbar_wait(&tfull_bar[acc], acc_phase);
          w_tfull += clock64() - t0;
        } else {
          mbar_wait(&tfull_bar[acc], acc_phase);
        }
        tc_fence_after();
#define GH_LEAN(A, R) lean_tile<BN, A, R>(ep, st, bias_s, t_row, half, lane, sub_row, col8, n0, rows, rowmask)
        const long long t_lean0 = p.prof ? clock64() : 0;
        if (ep.residual) {
          switch (ep.act) {
            case ACT_GELU_TANH: GH_LEAN(ACT_GELU_TANH, true); break;
            case ACT_QUICK_GELU: GH_LEAN(ACT_QUICK_GELU, true); break;
            case ACT_SILU: GH_LEAN(ACT_SILU, true); break;
            case ACT_GELU_ERF: GH_LEAN(ACT_GELU_ERF, true); break;
            default: GH_LEAN(ACT_NONE, true); break;
          }
        } else {
          switch (ep.act) {
            case ACT_GELU_TANH: GH_LEAN(ACT_GELU_TANH, false); break;
            case ACT_QUICK_GELU: GH_LEAN(ACT_QUICK_GELU, false); break;
            case ACT_SILU: GH_LEAN(ACT_SILU, false); break;
            case ACT_GELU_ERF: GH_LEAN(ACT_GELU_ERF, false); break;
            default: GH_LEAN(ACT_NONE, false); break;
          }
        }
#undef GH_LEAN
        if (p.prof) w_lean += clock64() - t_lean0;
      } else {
        // ---------------- general path (every epilogue option, edge tiles) ----------------
        int ns = (p.N - n0 + 15) / 16;   // slices of this tile that hold real columns
        if (ns > BN / 16) ns = BN / 16;
        SliceLoads LA, LB;
        auto issue = [&](int sl, SliceLoads& L) {
          const int gn = n0 + sl * 16 + col8;
          slice_issue_loads(ep, rows, gn < p.N ? rowmask : 0u, gn + 4 < p.N, gn, L);
        };
        auto run = [&](int sl, const SliceLoads& L, int next_sl, SliceLoads& Lnext) {
          uint32_t r[16];
          tmem_ld_32x16(t_row + sl * 16, r);
          if (next_sl < ns) issue(next_sl, Lnext);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(st + lane * Cfg::EPI_LD);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                 __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
          __syncwarp();
          const int gn = n0 + sl * 16 + col8;
          slice_finish(ep, st, Cfg::EPI_LD, rows, gn < p.N ? rowmask : 0u, gn + 4 < p.N, sub_row, col8, gn, L);
          __syncwarp();
        };
        if (half < ns) issue(half, LA);
        if (p.prof) {
          const long long t0 = clock64();
          mbar_wait(&tfull_bar[acc], acc_phase);
          w_tfull += clock64() - t0;
        } else {
          mbar_wait(&tfull_bar[acc], acc_phase);
        }
        tc_fence_after();
#pragma unroll 1
        for (int sl = half; sl < ns; sl += 4) {
          run(sl, LA, sl + 2, LB);
          if (sl + 2 < ns) run(sl + 2, LB, sl + 4, LA);
        }
      }
      tc_fence_before();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster_relaxed(tempty_leader[acc]);   // the leader's issuer waits for both CTAs' epilogues
        else mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (tma_epi) bulk_wait_group_all();   // every bulk store of this thread has left smem AND reached memory
    if (p.prof && warp == 4 && lane == 0) {
      p.prof[blockIdx.x * 8 + 4] = clock64() - e_begin;
      p.prof[blockIdx.x * 8 + 5] = w_tfull;
      p.prof[blockIdx.x * 8 + 7] = w_lean;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();  // neither CTA leaves while the other may still signal its barriers / use its smem
  if (warp == 2) {
    if (CTA2) tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace gh
