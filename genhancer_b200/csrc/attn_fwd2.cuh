// genhancer_b200 -- attention forward, second form (included by attn_sm100.cu after attn_bwd2.cuh): 128-key blocks, P handed
// to the PV MMA through TMEM.
//
// Why: tools/umma_bench.cu shows that a tcgen05.mma with M = 128 costs ~65-80 cycles whether N is 64 or 128 (the operand
// fetch, not the math, sets the floor for narrow N), and that every change of accumulator costs ~190 more.  The first form
// (flash_fwd_kernel) works on 64-key blocks: its S = Q K^T MMAs are N = 64 -- half the math per instruction for the same
// time -- and it switches accumulator twice per 64 keys.  Here a block is 128 keys: S = Q K_j^T is 128 x 128 x D, the
// softmax threads write P (bf16) over the fp32 score columns they read (tcgen05.st), and O += P V_j takes P out of TMEM
// (TS-mode MMA), so P never touches shared memory and half as many accumulator switches happen per key.
//
// RESULT (profiles/r02_attn_v2.txt): correct (same parity cases as the first form), but NOT faster -- 173 vs 161 us at
// B = 32, H = 24, L = 442, D = 128; 668 vs 668 us at L = 2169; 133 vs 105 us for the ViT's D = 64, L = 577.  The forward is
// not bound by MMA issue: with two (three) CTAs per SM its limit is the S -> softmax -> P -> PV chain of each CTA, and this
// form lengthens that chain (S single-buffered).  Kept for A/B builds (-DGH_ATTN_FWD_V2); the product runs flash_fwd_kernel.
//
// CTA = 128 queries of one (head, sample); 8 softmax warps -- TWO threads per query row, 64 columns each, the row maximum
// exchanged through shared memory -- and one control warp (TMA + MMA issue).  S single-buffered (128 columns) + O (D
// columns) = 256 TMEM columns and <= 99 KB of shared memory: two CTAs per SM, one's softmax under the other's MMAs.
#pragma once

namespace gh {

constexpr int ATT_FWD2_SW = 8;                          // softmax warps (also the index of the control warp)
constexpr int ATT_FWD2_THREADS = 32 * (ATT_FWD2_SW + 1);

template <int D>
struct AttnFwd2Cfg {
  static constexpr int T_BYTES = 128 * D * 2;          // Q, K_j or V_j
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = T_BYTES;
  static constexpr int OFF_V = 2 * T_BYTES;
  static constexpr int OFF_RED = 3 * T_BYTES;          // float [2][128]: per-half row maxima, then row sums
  static constexpr int OFF_BAR = OFF_RED + 2 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int TMEM_COLS = 256;
  static constexpr int TM_S = 0, TM_O = 128;
};

template <int D>
__global__ void __launch_bounds__(ATT_FWD2_THREADS, 2)
flash_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  using Cfg = AttnFwd2Cfg<D>;
  constexpr int DC = D / 64;
  constexpr int CH = 128 * 128;    // bytes between the 64-lane chunks of the head dim inside a 128-row tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  float* sRed = reinterpret_cast<float*>(smem + Cfg::OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;    // S_j in TMEM (=> the K tile is free)
  uint64_t* bar_p = bars + 4;    // P_j written, O rescaled (8 warps)
  uint64_t* bar_o = bars + 5;    // PV_j retired (=> the V tile is free, O stable)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Lk + 127) / 128;

  if (warp == ATT_FWD2_SW && lane == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, ATT_FWD2_SW);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == ATT_FWD2_SW) {
    // ---------------- control warp: TMA + MMA issue (all lanes run the flow, one elected lane issues) ----------------
    const uint32_t idesc_o = umma_idesc_bf16(128, p.dvalid, false, true);   // N = the head lanes that hold data
    const int ksteps = p.dvalid >> 4;
    const uint64_t kdesc = umma_desc_base(16u, 1024u);
    const uint64_t vdesc = umma_desc_base(static_cast<uint32_t>(CH), 1024u);   // V_j as MN-major B operand
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
    auto load_tile = [&](uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int row0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar, Cfg::T_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) tma_load_4d(dst + c * CH, m, bar, c * 64, row0, h, b);
      }
    };
    auto issue_s = [&](int j) {
      if (elect_one()) {
        mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_S, kdesc, aQ, aK, umma_idesc_bf16(128, bwd2_block_n(p.Lk - j * 128), false, false));
        umma_commit(bar_s);
      }
    };
    load_tile(sQ, &tm_q, bar_q, q0);
    load_tile(sK, &tm_k, bar_k, 0);
    load_tile(sV, &tm_v, bar_v, 0);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      const bool more = j + 1 < nkv;
      mbar_wait(bar_s, j & 1);                 // S_j complete: the K tile is free
      if (more) load_tile(sK, &tm_k, bar_k, (j + 1) * 128);
      mbar_wait(bar_v, j & 1);                 // (landed during the softmax of the block before)
      mbar_wait(bar_p, j & 1);                 // P_j is in TMEM, O has been rescaled
      tc_fence_after();
      if (elect_one()) {
        mma_ts_ksteps<64>(bwd2_block_n(p.Lk - j * 128) >> 4, tmem + Cfg::TM_O, tmem + Cfg::TM_S, vdesc, aV, idesc_o, j != 0 ? 1u : 0u);
        umma_commit(bar_o);
      }
      if (more) {
        mbar_wait(bar_k, (j + 1) & 1);         // (PV_j is queued: this wait costs the pipe nothing unless the tile is late)
        tc_fence_after();
        issue_s(j + 1);                        // behind PV_j in the pipe: P_j has been consumed
        mbar_wait(bar_o, j & 1);               // PV_j retired: the V tile is free
        load_tile(sV, &tm_v, bar_v, (j + 1) * 128);
      }
    }
    __syncwarp();
  } else {
    // ---------------- softmax / correction / epilogue: two threads per query row ----------------
    const int row = threadIdx.x & 127;   // == TMEM lane
    const int part = threadIdx.x >> 7;   // which 64 of a block's 128 key columns (and which half of the head lanes of O)
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t t_s = t_lane + Cfg::TM_S + part * 64;
    constexpr int OH = D / 2;            // O columns per thread
    const uint32_t t_o = t_lane + Cfg::TM_O + part * OH;
    float m_used = 0.f, l_sum = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int kv_left = p.Lk - j * 128;   // valid keys in this block
      const bool active = part * 64 < bwd2_block_n(kv_left);   // (warp-uniform) my columns exist in this block's MMAs
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      uint32_t s[64];
      if (active) {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        tmem_ld_32x32(t_s, s0);
        tmem_ld_32x32(t_s + 32, s1);
        tmem_ld_wait();
        if (kv_left < 128) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (part * 64 + c >= kv_left) s[c] = 0xff800000u;  // -inf
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = 0xff800000u;
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], __uint_as_float(s[c + u]));
      }
      const float mloc = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      sRed[part * 128 + row] = mloc;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_FWD2_SW) : "memory");
      const float mx = fmaxf(mloc, sRed[(part ^ 1) * 128 + row]) * p.scale_log2;  // scale > 0; both threads of a row agree
      float factor = 1.f;
      if (j == 0) {
        m_used = mx;
      } else if (mx > m_used + 8.f) {
        factor = ex2_approx(m_used - mx);
        m_used = mx;
      }
      const float neg_m = -m_used;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        const float e0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, neg_m));
        const float e1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, neg_m));
        pk[c >> 1] = pack_bf16x2(e0, e1);
        rs0 += e0;
        rs1 += e1;
      }
      l_sum = l_sum * factor + (rs0 + rs1);
      if (j > 0) {
        mbar_wait(bar_o, (j - 1) & 1);  // PV_{j-1} retired: O stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, factor != 1.f)) {
#pragma unroll
          for (int c = 0; c < OH / 32; ++c) {
            if (part * OH + c * 32 >= p.dvalid) break;     // (columns >= dvalid were never written by the PV MMA)
            uint32_t o[32];
            tmem_ld_32x32(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32(t_o + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      if (active) {
        tmem_st_32x32(t_s, pk);          // P over the first half of the scores it came from
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue ----
    mbar_wait(bar_o, (nkv - 1) & 1);
    tc_fence_after();
    sRed[part * 128 + row] = l_sum;      // (the maxima of the last block were read before that block's bar_p arrive)
    asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_FWD2_SW) : "memory");
    l_sum += sRed[(part ^ 1) * 128 + row];
    const int l = q0 + row;
    const float inv = 1.f / l_sum;
    if (part == 0 && l < p.Lq && p.lse2) p.lse2[(static_cast<int64_t>(b) * p.H + h) * p.Lq + l] = m_used + log2f(l_sum);
    bf16* orow = (l < p.Lq) ? p.o.row(b, l) + h * D + part * OH : nullptr;
#pragma unroll
    for (int c = 0; c < OH / 32; ++c) {
      uint32_t o[32];
      const int col0 = part * OH + c * 32;
      if (col0 < p.dvalid) {                   // (a warp-uniform branch: tcgen05.ld is .sync.aligned)
        tmem_ld_32x32(t_o + c * 32, o);
        tmem_ld_wait();
      }
      if (orow) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w = make_uint4(0u, 0u, 0u, 0u);          // pad lanes: exact zeros
          if (col0 + i * 8 < p.dvalid) {
            w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          }
          *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

}  // namespace gh
