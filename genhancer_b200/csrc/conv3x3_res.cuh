// genhancer_b200 -- 3x3 / stride-1 / pad-1 convolution with 128 input and 128 output channels, weights RESIDENT in
// shared memory and the input patch read ONCE per column shift instead of once per filter tap.
//
// Why a second conv kernel.  The FLUX AE encoder's first level (src/flux/modules/autoencoder.py:62-82, ResnetBlock
// conv1 / conv2 at ch = 128 on the full-resolution map; 4 launches per 336x336 batch, 5.7 ms of a 97 ms step) is the
// one conv the generic implicit GEMM (umma_gemm_kernel<.., MODE_CONV>) cannot feed: N = 128 output channels give
// every A byte only 128 MACs, so a 256 x 128 pair tile needs 24 KB of operands per 256 MMA cycles and CTA --
// 96 B/clk/SM out of L2, ~80 % of what the L2 -> SM path delivers -- and the kernel sat at 750 TFLOP/s, L2-bandwidth
// bound, re-reading the same input patch for each of the 9 taps and the same 295 KB of weights for each tile.
//
// This kernel (CTA pairs, tcgen05 cta_group::2, one 256-pixel x 128-channel tile per pair):
//   * weights: each CTA keeps ITS 64 output channels x 1152 K (= 18 SWIZZLE_128B tiles of 8 KB, 144 KB) in shared
//     memory for the whole launch -- loaded once by TMA, never re-read;
//   * input: the output patch of a CTA is 16 rows x 8 columns.  For a column shift kw in {0,1,2} and a 64-channel
//     chunk, ONE 4-D TMA box [18 rows x 8 columns x 64 ch] (18 KB, out-of-image rows / columns zero-filled = padding)
//     holds the A operand of all three row taps kh: tap kh is the same buffer read from byte offset kh * 1024 (one
//     halo row = 8 pixels x 128 B = one swizzle atom, so every tap's descriptor stays 1024-byte aligned and uses the
//     ordinary K-major SWIZZLE_128B layout with SBO = 1024).  Six boxes per tile instead of eighteen 16 KB ones:
//     L2 -> SM traffic per CTA and tile 110 KB instead of 435 KB;
//   * 12 MMAs (3 taps x 4 k-steps of 16) per stage, 3 stages, two 128-column TMEM accumulators so the epilogue of
//     tile i overlaps the MMAs of tile i + 1; the epilogue is the lean one of the GEMM (bias [+ residual] -> bf16).
#pragma once
#include "umma_gemm.cuh"

namespace gh {

struct Conv3x3ResCfg {
  static constexpr int BN = 128;
  static constexpr int TW = 8, TH = 16;                     // output patch of one CTA (128 pixels)
  static constexpr int TAPS = 9, CHUNKS = 2;                // Cin = 128 = 2 chunks of 64 channels
  static constexpr int W_TILE_BYTES = 64 * 64 * 2;          // [64 Cout rows x 64 ch] per CTA, tap and chunk
  static constexpr int W_BYTES = TAPS * CHUNKS * W_TILE_BYTES;   // 147456
  static constexpr int A_STAGE_BYTES = (TH + 2) * TW * 128;      // 18432: [18 rows x 8 px] x 64 ch
  static constexpr int STAGES = 3;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int EPI_WARP_BYTES = 3072;               // fp32 transposition scratch (2560 B) + staged bias (256 B)
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = W_BYTES + STAGES * A_STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// Epilogue of one warp for one 128-channel tile: bf16 out = acc + bias [+ bf16 residual].  Same arithmetic and staging
// as lean_tile (umma_gemm.cuh), but the residual of the WHOLE tile (2 rows x 4 slices x 16 B per lane) is requested
// by the caller BEFORE it waits for the accumulator: the residual streams from HBM (925 MB per launch), and with the
// one-slice-ahead prefetch of lean_tile 29 % of all warp-stall samples of this kernel sat on the first use of those
// loads (ncu source page, profiles/r02_kernels_ncu.txt) -- four exposed DRAM round trips per tile and warp, which held
// the tensor pipe at 56 %.
template <bool HAS_RES>
__device__ __forceinline__ void conv_res_tile(const EpilogueParams& ep, float* __restrict__ st, const float* __restrict__ bias_s,
                                              uint32_t t_row, int half, int lane, int sub_row, int col8,
                                              const int (&rows)[2], uint32_t rowmask, const uint4 (&res)[4][2]) {
  constexpr int BN = Conv3x3ResCfg::BN;
  constexpr int LD = GemmCfg<BN>::EPI_LD;
  constexpr int NJ = BN / 32;
  const bool ok0 = rowmask & 1u, ok1 = (rowmask >> 1) & 1u;
  __nv_bfloat16* d0 = static_cast<__nv_bfloat16*>(ep.d) + static_cast<int64_t>(rows[0]) * ep.ldd + col8 + half * 16;
  __nv_bfloat16* d1 = static_cast<__nv_bfloat16*>(ep.d) + static_cast<int64_t>(rows[1]) * ep.ldd + col8 + half * 16;
  const float* sp0 = st + sub_row * LD + col8;
  const float* sp1 = sp0 + 16 * LD;
  float4* dst = reinterpret_cast<float4*>(st + lane * LD);
  uint32_t treg[16];
  tmem_ld_32x16(t_row + half * 16, treg);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_float4(__uint_as_float(treg[4 * q]), __uint_as_float(treg[4 * q + 1]),
                           __uint_as_float(treg[4 * q + 2]), __uint_as_float(treg[4 * q + 3]));
    __syncwarp();
    if (j + 1 < NJ) tmem_ld_32x16(t_row + half * 16 + (j + 1) * 32, treg);
    const float4 bb0 = *reinterpret_cast<const float4*>(bias_s + j * 16 + col8);
    const float4 bb1 = *reinterpret_cast<const float4*>(bias_s + j * 16 + col8 + 4);
    const float4 a0 = *reinterpret_cast<const float4*>(sp0), a1 = *reinterpret_cast<const float4*>(sp0 + 4);
    const float4 c0 = *reinterpret_cast<const float4*>(sp1), c1 = *reinterpret_cast<const float4*>(sp1 + 4);
    float v0[8] = {a0.x + bb0.x, a0.y + bb0.y, a0.z + bb0.z, a0.w + bb0.w, a1.x + bb1.x, a1.y + bb1.y, a1.z + bb1.z, a1.w + bb1.w};
    float v1[8] = {c0.x + bb0.x, c0.y + bb0.y, c0.z + bb0.z, c0.w + bb0.w, c1.x + bb1.x, c1.y + bb1.y, c1.z + bb1.z, c1.w + bb1.w};
    if (HAS_RES) {
      float r[8];
      unpack8(res[j][0], r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v0[k] += r[k];
      unpack8(res[j][1], r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v1[k] += r[k];
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

__global__ void __launch_bounds__(Conv3x3ResCfg::THREADS, 1)
conv3x3_res_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const GemmParams p) {
  using Cfg = Conv3x3ResCfg;
  constexpr int BN = Cfg::BN, STAGES = Cfg::STAGES;
  const uint32_t rank = cluster_ctarank();
  const int worker = static_cast<int>(blockIdx.x >> 1);
  const int num_workers = static_cast<int>(gridDim.x >> 1);
  const int num_tiles = (p.num_m_blocks + 1) / 2;          // pair tiles

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* w_base = smem;                                   // 18 x 8 KB, 1024-aligned
  uint8_t* a_base = smem + Cfg::W_BYTES;                    // STAGES x 18 KB, 1024-aligned
  uint8_t* epi_base = a_base + STAGES * Cfg::A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint64_t* w_bar = tempty_bar + 2;          // weights resident
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_img = p.cv.tiles_w * p.cv.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], Cfg::EPI_WARPS * 2);
    mbar_init(&tempty_bar[1], Cfg::EPI_WARPS * 2);
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc2<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // pixel origin of this CTA's patch in tile `t` (pair tile index): (image, row, column)
  auto patch_of = [&](int t, int& cb, int& h0, int& w0) {
    const int m_blk = t * 2 + static_cast<int>(rank);
    cb = m_blk / tiles_per_img;
    const int r = m_blk - cb * tiles_per_img;
    h0 = (r / p.cv.tiles_w) * Cfg::TH;
    w0 = (r % p.cv.tiles_w) * Cfg::TW;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // resident weights: this CTA's 64 output channels; both CTAs' bytes are counted on the leader's barrier
      const uint32_t wb = mapa_u32(smem_u32(w_bar), 0u);
      if (rank == 0) mbar_arrive_expect_tx(w_bar, 2u * Cfg::W_BYTES);
      for (int kb = 0; kb < Cfg::TAPS * Cfg::CHUNKS; ++kb)
        tma2_load_2d(w_base + kb * Cfg::W_TILE_BYTES, &tmap_w, wb, kb * 64, static_cast<int>(rank) * 64);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    // Three 18 KB stages hold ~2300 MMA cycles of input -- enough to cover an L2 hit, not a DRAM miss (the first version
    // ran the tensor pipe at 56 %: ncu profiles/r02_kernels_ncu.txt).  The weights leave no room for a fourth stage, so
    // the input of the tile AFTER the next one is pulled into L2 while this tile is loaded: the two outer column
    // shifts (kw = 0, 2) of each channel chunk cover every pixel the three shifts read.
    auto prefetch_tile = [&](int t) {
      if (t >= num_tiles || (p.dbg & 2)) return;     // (dbg bit 1: A/B switch GH_CONV_PREFETCH=0)
      int cb, h0, w0;
      patch_of(t, cb, h0, w0);
      if (elect_one()) {
#pragma unroll
        for (int chunk = 0; chunk < Cfg::CHUNKS; ++chunk) {
          tma_prefetch_l2_4d(&tmap_x, chunk * 64, w0 - 1, h0 - 1, cb);
          tma_prefetch_l2_4d(&tmap_x, chunk * 64, w0 + 1, h0 - 1, cb);
        }
      }
      __syncwarp();
    };
    prefetch_tile(worker + num_workers);
    for (int t = worker; t < num_tiles; t += num_workers) {
      int cb, h0, w0;
      patch_of(t, cb, h0, w0);
      prefetch_tile(t + 2 * num_workers);
      for (int s = 0; s < Cfg::CHUNKS * 3; ++s) {
        const int chunk = s / 3, kw = s - chunk * 3;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * Cfg::A_STAGE_BYTES);
          tma2_load_4d(a_base + stage * Cfg::A_STAGE_BYTES, &tmap_x, fb, chunk * 64, w0 + kw - 1, h0 - 1, cb);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(256, BN, false, false);
    const uint64_t desc_base = umma_desc_base(16u, 1024u);
    mbar_wait(w_bar, 0u);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t sW = smem_u32(w_base);
    for (int t = worker; t < num_tiles; t += num_workers) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int s = 0; s < Cfg::CHUNKS * 3; ++s) {
        const int chunk = s / 3, kw = s - chunk * 3;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_u32(a_base + stage * Cfg::A_STAGE_BYTES);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t a_tap = sA + kh * 1024;                                        // halo row kh: one swizzle atom down
            const uint32_t b_tap = sW + ((kh * 3 + kw) * Cfg::CHUNKS + chunk) * Cfg::W_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma2_ss(d_tmem, umma_desc_at(desc_base, a_tap + k * 32), umma_desc_at(desc_base, b_tap + k * 32), idesc,
                       (s | kh | k) != 0 ? 1u : 0u);
          }
          umma2_commit(&empty_bar[stage]);
          if (s == Cfg::CHUNKS * 3 - 1) umma2_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp >= 4) {
    // ===================== epilogue: bf16 out = alpha * acc + bias [+ bf16 residual] (lean path of the GEMM) ===========
    const int ew = warp - 4;
    const int quad = ew & 3;
    const int half = ew >> 2;
    float* st = reinterpret_cast<float*>(epi_base + ew * Cfg::EPI_WARP_BYTES);
    float* bias_s = st + 32 * GemmCfg<BN>::EPI_LD;
    const int sub_row = lane & 15;
    const int col8 = (lane >> 4) * 8;
    const EpilogueParams& ep = p.ep;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader[2] = {mapa_u32(smem_u32(&tempty_bar[0]), 0u), mapa_u32(smem_u32(&tempty_bar[1]), 0u)};
    // bias of this warp's columns (the same for every tile): lane + 32 i -> column (half + 2 (idx >> 4)) * 16 + (idx & 15)
    {
      constexpr int NB = (BN / 2) / 32;
      float bv[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int idx = lane + 32 * i;
        const int col = (half + 2 * (idx >> 4)) * 16 + (idx & 15);
        bv[i] = 0.f;
        if (ep.bias)
          bv[i] = ep.bias_f32 ? __ldg(static_cast<const float*>(ep.bias) + col)
                              : __bfloat162float(static_cast<const __nv_bfloat16*>(ep.bias)[col]);
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) bias_s[lane + 32 * i] = bv[i];
    }
    __syncwarp();
    for (int t = worker; t < num_tiles; t += num_workers) {
      int cb, h0, w0;
      patch_of(t, cb, h0, w0);
      int rows[2];
      uint32_t rowmask = 0;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int tr = quad * 32 + it * 16 + sub_row;      // row of the 128-row patch: (tr / 8, tr % 8)
        const int oh = h0 + tr / Cfg::TW, ow = w0 + tr % Cfg::TW;
        const bool ok = (oh < p.cv.Ho) && (ow < p.cv.Wo) && (cb < p.cv.B);
        rows[it] = (cb * p.cv.Ho + oh) * p.cv.Wo + ow;
        rowmask |= static_cast<uint32_t>(ok) << it;
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
      // the residual of the whole tile is requested now, a tile's worth of MMAs before it is used
      uint4 res[4][2];
      if (ep.residual) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(ep.residual) + static_cast<int64_t>(rows[it]) * ep.ld_res +
                                    col8 + half * 16;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            res[j][it] = (rowmask >> it) & 1u ? __ldg(reinterpret_cast<const uint4*>(rp + j * 32)) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (ep.residual) conv_res_tile<true>(ep, st, bias_s, t_row, half, lane, sub_row, col8, rows, rowmask, res);
      else conv_res_tile<false>(ep, st, bias_s, t_row, half, lane, sub_row, col8, rows, rowmask, res);
      tc_fence_before();
      if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace gh
