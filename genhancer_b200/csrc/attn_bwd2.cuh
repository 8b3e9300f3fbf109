// genhancer_b200 -- attention backward, second form: 128 x 128 score tiles, P^T / dS^T / dS handed to the tensor core
// THROUGH TMEM (the A operand of tcgen05.mma may live there), included by attn_sm100.cu.
//
// Why (DESIGN.md finding 24, profiles/r02_attn_phase_counters.txt): the first form worked on 128 x 64 tiles, wrote P^T
// and dS^T to shared memory and fed every MMA from shared memory.  Two thirds of its MMAs were 128 x 64 x 16 -- 6 KB of
// operands for 32 cycles of math, i.e. paced by the operand fetch, not the tensor pipe -- and the 128-row K / V tiles
// were re-read for every 64 queries: the tensor pipe sat at 29 %.  Here every MMA is 128 x 128 x 16 (N = the whole tile
// or the valid head lanes), the element-wise warps write their bf16 results over the fp32 columns they were computed
// from (tcgen05.st) and the accumulating MMAs read that operand out of TMEM: per 128 x 128 tile the shared-memory
// operand traffic of the dK/dV kernel drops from 320 KB to 192 KB and no element-wise result touches shared memory.
//
//   dK/dV : CTA = (128 keys, h, b), loop over 128-query blocks i, on the TRANSPOSED tile (TMEM lane = key):
//             S^T = K Q_i^T           (SS)      P^T  = exp2(S^T c - lse_q)            -> bf16 over the S^T columns
//             dP^T = V dO_i^T         (SS)      dS^T = P^T (dP^T scale - delta_q scale) -> bf16 over the dP^T columns
//             dV += P^T dO_i          (TS)      dK += dS^T Q_i                        (TS)
//           TMEM: S^T 128 | dP^T 128 | dV D | dK D columns.  Issue order dV_i, S_{i+1}, dK_i, dP_{i+1}: the tensor pipe
//           runs its MMAs in order, so S_{i+1} may overwrite the columns P^T_i was read from, and while the compute warps
//           turn S_{i+1} into P^T_{i+1} the pipe works on dK_i and dP_{i+1}.
//   dQ    : CTA = (128 queries, h, b), loop over 128-key blocks j (TMEM lane = query):
//             S = Q K_j^T, dP = dO V_j^T (SS);  dS = exp2(S c - lse) (dP scale - delta scale) -> bf16 over S;  dQ += dS K_j (TS)
//           TMEM: S (2 buffers) | dP | dQ; dS_j is written over S_j (dead once its exponentials sit in registers), so
//           dP_{j+1} needs only "dP_j has been read" and is issued AHEAD of dQ_j: issue order dP_{j+1}, dQ_j, S_{j+2}.  The
//           exponentials of block j + 1 run under the MMAs of block j.
// Still 7 GEMMs for the algorithm's 5 (S and dP are computed in both kernels): no atomics, bit-reproducible.
//
// A k-step of a TMEM-resident A operand is 16 bf16 = 8 columns.  The compute thread that owns columns [32 q, 32 q + 32)
// of a row writes its 16 packed columns at [32 q, 32 q + 16) -- inside the range it read, so no thread overwrites
// another's unread scores -- and the MMA takes k-step ks from column 32 (ks / 2) + 8 (ks % 2).
#pragma once

namespace gh {

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}

// D[tmem] (+)= A[tmem, packed as described above] * B[smem, MN-major, 2048 B per k-step].  SPAN = fp32 columns a compute
// thread owns (32 in the backward kernels, 64 in the forward): its SPAN / 2 packed columns start at the span's first column.
template <int KS, int SPAN>
__device__ __forceinline__ void mma_group_ts(uint32_t tmem_d, uint32_t a_tmem, uint64_t b_base, uint32_t b_addr, uint32_t idesc,
                                             uint32_t acc_first) {
  constexpr int KPS = SPAN / 16;   // k-steps per span
  const uint32_t b_hi = static_cast<uint32_t>(b_base >> 32);
  const uint32_t b_lo = static_cast<uint32_t>(b_base) | ((b_addr >> 4) & 0x3FFFu);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
    umma_ts(tmem_d, a_tmem + static_cast<uint32_t>(SPAN * (ks / KPS) + 8 * (ks % KPS)),
            desc_join(b_hi, b_lo + static_cast<uint32_t>(ks * (2048 >> 4))), idesc, ks == 0 ? acc_first : 1u);
}
template <int SPAN = 32>
__device__ __forceinline__ void mma_ts_ksteps(int ksteps, uint32_t tmem_d, uint32_t a_tmem, uint64_t b_base, uint32_t b_addr,
                                              uint32_t idesc, uint32_t acc_first) {
  switch (ksteps) {
    case 1: mma_group_ts<1, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
    case 2: mma_group_ts<2, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
    case 3: mma_group_ts<3, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
    case 4: mma_group_ts<4, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
    case 5: mma_group_ts<5, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
    case 6: mma_group_ts<6, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
    case 7: mma_group_ts<7, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
    default: mma_group_ts<8, SPAN>(tmem_d, a_tmem, b_base, b_addr, idesc, acc_first); break;
  }
}

// Accumulator rows (one TMEM lane per thread, 32-column chunks shared out over the threads of a row) -> bf16, staged in
// shared memory as [128 rows x 64 lanes] SWIZZLE_128B tiles (one per 64 head lanes, 16 KB apart) for ONE bulk tensor
// store per tile: full 128-byte lines, rows past the sequence end clipped by the hardware.  (The first form stored
// 16 bytes per lane at a 256-byte stride straight from registers: 32 half-filled sectors per instruction, ~5000 cycles
// per CTA for 64 KB.)
template <int D>
__device__ __forceinline__ void bwd2_stage_rows(uint32_t t_acc, int part, int row, int dvalid, uint8_t* tile) {
#pragma unroll
  for (int c = part; c < D / 32; c += ATT_BWD_SPLIT) {
    uint32_t o[32];
    if (c * 32 < dvalid) {                     // (warp-uniform: part is per warp)
      tmem_ld_32x32(t_acc + c * 32, o);
      tmem_ld_wait();
    }
    uint8_t* t = tile + (c >> 1) * (128 * 128);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      uint4 w = make_uint4(0u, 0u, 0u, 0u);   // pad lanes: exact zeros
      if (c * 32 + q4 * 8 < dvalid) {
        w.x = pack_bf16x2(__uint_as_float(o[8 * q4 + 0]), __uint_as_float(o[8 * q4 + 1]));
        w.y = pack_bf16x2(__uint_as_float(o[8 * q4 + 2]), __uint_as_float(o[8 * q4 + 3]));
        w.z = pack_bf16x2(__uint_as_float(o[8 * q4 + 4]), __uint_as_float(o[8 * q4 + 5]));
        w.w = pack_bf16x2(__uint_as_float(o[8 * q4 + 6]), __uint_as_float(o[8 * q4 + 7]));
      }
      st_sw128(t, row, (c & 1) * 4 + q4, w);
    }
  }
}

__device__ __forceinline__ int bwd2_block_n(int left) {   // MMA N / reduction length of a (possibly ragged) 128-wide block
  return left >= 128 ? 128 : ((left + 15) & ~15);
}

constexpr int ATT_BWD2_MMA_WARP = ATT_BWD_CW;        // warp 16: issues the MMAs
constexpr int ATT_BWD2_TMA_WARP = ATT_BWD_CW + 1;    // warp 17: issues the loads
constexpr int ATT_BWD2_THREADS = 32 * (ATT_BWD_CW + 2);

// Roles.  The tensor core's instruction queue is short (~2-3 MMAs, tools/umma_bench.cu + the in-kernel counters of
// profiles/r02_attn_phase_counters.txt): whatever the issuing warp does BETWEEN two chains of MMAs -- an mbarrier
// wait costs ~100 cycles even when the phase has long completed, a commit, a TMA issue -- drains it.  So the MMA warp
// does nothing but { wait for the element-wise warps, issue two chains back to back, commit }, with the waits on its
// operand tiles hoisted to where the pipe has two chains queued, and a separate warp owns the loads; "slot free" is read
// off the barriers the MMA warp commits anyway (the pipe retires in order: S_{i+1} complete => dV_i complete).

template <int D>
struct AttnBwdKV2Cfg {
  static constexpr int T_BYTES = 128 * D * 2;   // a 128-row tile: K, V, Q_i or dO_i
  static constexpr int NQ = 3;                  // Q ring: Q_i feeds S_i (early) and dK_i (late), Q_{i+1} is live in between
  static constexpr int NDO = 2;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = T_BYTES;
  static constexpr int OFF_Q = 2 * T_BYTES;
  static constexpr int OFF_DO = OFF_Q + NQ * T_BYTES;
  static constexpr int OFF_STAT = OFF_DO + NDO * T_BYTES;   // float [2][2][128]: -lse, delta * scale per query, two blocks
  static constexpr int OFF_BAR = OFF_STAT + 2 * 2 * 128 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128;          // D = 128: 231552 of the 232448 a CTA may have -- no alignment slack
  static constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 384;
};

template <int D>
__global__ void __launch_bounds__(ATT_BWD2_THREADS, 1)
flash_bwd_dkv2_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                      const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                      const AttnBwdParams p) {
  using Cfg = AttnBwdKV2Cfg<D>;
  constexpr int DC = D / 64;
  constexpr int CH = 128 * 128;    // bytes between the 64-lane chunks of the head dim inside a 128-row tile
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sDO = smem + Cfg::OFF_DO;
  float* sStat = reinterpret_cast<float*>(smem + Cfg::OFF_STAT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_k = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_q = bars + 2;    // [3]
  uint64_t* bar_do = bars + 5;   // [2]
  uint64_t* bar_s = bars + 7;    // S^T_i in TMEM (=> dV_{i-1} retired: its dO slot is free)
  uint64_t* bar_dp = bars + 8;   // dP^T_i in TMEM (=> dK_{i-1} retired: its Q slot is free)
  uint64_t* bar_p = bars + 9;    // P^T_i written (16 warps)
  uint64_t* bar_ds = bars + 10;  // dS^T_i written (16 warps)
  uint64_t* bar_done = bars + 11;  // every MMA retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nq = (p.Lq + 127) / 128;

  if (warp == ATT_BWD2_MMA_WARP && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("genhancer_b200: flash_bwd_dkv2: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    for (int i = 0; i < Cfg::NQ; ++i) mbar_init(&bar_q[i], 1);
    for (int i = 0; i < Cfg::NDO; ++i) mbar_init(&bar_do[i], 1);
    mbar_init(bar_s, 1); mbar_init(bar_dp, 1);
    mbar_init(bar_p, ATT_BWD_CW); mbar_init(bar_ds, ATT_BWD_CW);
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == ATT_BWD2_TMA_WARP) {
    // ---------------- loads ----------------
    auto load_tile = [&](uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int row0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar, Cfg::T_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) tma_load_4d(dst + c * CH, m, bar, c * 64, row0, h, b);
      }
    };
    load_tile(sK, &tm_k, bar_k, k0);
    load_tile(sQ, &tm_q, &bar_q[0], 0);
    load_tile(sV, &tm_v, bar_v, k0);
    load_tile(sDO, &tm_do, &bar_do[0], 0);
    if (nq > 1) {
      load_tile(sQ + Cfg::T_BYTES, &tm_q, &bar_q[1], 128);
      load_tile(sDO + Cfg::T_BYTES, &tm_do, &bar_do[1], 128);
    }
    if (nq > 2) load_tile(sQ + 2 * Cfg::T_BYTES, &tm_q, &bar_q[2], 256);
    // (a parity wait can only tell the barrier's current phase from the one before it: every phase is waited for, in order)
    for (int i = 0; i < nq; ++i) {
      mbar_wait(bar_s, i & 1);           // S^T_i complete => dV_{i-1} (issued before it) retired: dO slot (i - 1) % 2 is free
      if (i >= 1 && i + 1 < nq)
        load_tile(sDO + ((i + 1) % Cfg::NDO) * Cfg::T_BYTES, &tm_do, &bar_do[(i + 1) % Cfg::NDO], (i + 1) * 128);
      mbar_wait(bar_dp, i & 1);          // dP^T_i complete => dK_{i-1} retired: Q slot (i - 1) % 3 is free
      if (i >= 1 && i + 2 < nq)
        load_tile(sQ + ((i + 2) % Cfg::NQ) * Cfg::T_BYTES, &tm_q, &bar_q[(i + 2) % Cfg::NQ], (i + 2) * 128);
    }
    __syncwarp();
  } else if (warp == ATT_BWD2_MMA_WARP) {
    // ---------------- MMA issue: all lanes run the flow, one elected lane issues (see flash_fwd_kernel) ----------------
    const uint32_t idesc_a = umma_idesc_bf16(128, p.dvalid, false, true);   // accumulators: N = the head lanes that hold data
    const int ksteps = p.dvalid >> 4;                                       // k-steps of the MMAs that reduce over the head dim
    const uint64_t kdesc = umma_desc_base(16u, 1024u);                      // K-major tiles
    const uint64_t mdesc = umma_desc_base(static_cast<uint32_t>(CH), 1024u);   // Q / dO as MN-major B operands
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ0 = smem_u32(sQ), aD0 = smem_u32(sDO);
    mbar_wait(bar_k, 0);
    mbar_wait(&bar_q[0], 0);
    tc_fence_after();
    if (elect_one()) {
      mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_S, kdesc, aK, aQ0, umma_idesc_bf16(128, bwd2_block_n(p.Lq), false, false));
      umma_commit(bar_s);
    }
    mbar_wait(bar_v, 0);
    mbar_wait(&bar_do[0], 0);
    tc_fence_after();
    if (elect_one()) {
      mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_DP, kdesc, aV, aD0, umma_idesc_bf16(128, bwd2_block_n(p.Lq), false, false));
      umma_commit(bar_dp);
    }
    for (int i = 0; i < nq; ++i) {
      const bool more = i + 1 < nq;
      const int kq = bwd2_block_n(p.Lq - i * 128) >> 4;   // k-steps over this block's queries
      const uint32_t idesc_n = umma_idesc_bf16(128, bwd2_block_n(p.Lq - (i + 1) * 128), false, false);
      const uint32_t aQ = aQ0 + (i % Cfg::NQ) * Cfg::T_BYTES, aQn = aQ0 + ((i + 1) % Cfg::NQ) * Cfg::T_BYTES;
      const uint32_t aD = aD0 + (i % Cfg::NDO) * Cfg::T_BYTES, aDn = aD0 + ((i + 1) % Cfg::NDO) * Cfg::T_BYTES;
      if (more) {   // the next block's tiles landed long ago: look at their barriers while the pipe still has work queued
        mbar_wait(&bar_q[(i + 1) % Cfg::NQ], ((i + 1) / Cfg::NQ) & 1);
        mbar_wait(&bar_do[(i + 1) % Cfg::NDO], ((i + 1) / Cfg::NDO) & 1);
      }
      mbar_wait(bar_p, i & 1);                            // P^T_i is in TMEM
      tc_fence_after();
      if (elect_one()) {
        mma_ts_ksteps(kq, tmem + Cfg::TM_DV, tmem + Cfg::TM_S, mdesc, aD, idesc_a, i != 0 ? 1u : 0u);
        if (more) {                                       // behind dV_i in the pipe: P^T_i has been consumed
          mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_S, kdesc, aK, aQn, idesc_n);
          umma_commit(bar_s);
        }
      }
      mbar_wait(bar_ds, i & 1);                           // dS^T_i is in TMEM
      tc_fence_after();
      if (elect_one()) {
        mma_ts_ksteps(kq, tmem + Cfg::TM_DK, tmem + Cfg::TM_DP, mdesc, aQ, idesc_a, i != 0 ? 1u : 0u);
        if (more) {                                       // behind dK_i: dS^T_i has been consumed
          mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_DP, kdesc, aV, aDn, idesc_n);
          umma_commit(bar_dp);
        } else {
          umma_commit(bar_done);
        }
      }
    }
    __syncwarp();
  } else {
    const int row = threadIdx.x & 127;  // key row within the block == TMEM lane
    const int part = threadIdx.x >> 7;  // which 32 of the 128 query columns this thread owns
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int64_t stat_base = (static_cast<int64_t>(b) * p.H + h) * p.Lq;
    // Per-query statistics of a block, staged as (-lse, delta * scale); queries past Lq get -lse = -inf, so their P^T
    // and dS^T columns are exact zeros without a mask in the loops.  Loaded one block ahead, kept raw in a register
    // until they are stored (a thread stalls at the first USE of a load).
    auto stat_load = [&](int i) -> float {
      const int ql = i * 128 + row;
      if (threadIdx.x >= 256 || i >= nq) return 0.f;
      if (ql >= p.Lq) return part == 0 ? INFINITY : 0.f;
      return part == 0 ? p.lse2[stat_base + ql] : p.delta[stat_base + ql];
    };
    float stat_raw = stat_load(0);
    for (int i = 0; i < nq; ++i) {
      float* st = sStat + (i & 1) * 256;
      if (threadIdx.x < 256) st[threadIdx.x] = part == 0 ? -stat_raw : stat_raw * p.scale;
      stat_raw = stat_load(i + 1);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_BWD_CW) : "memory");
      const int nb = bwd2_block_n(p.Lq - i * 128);
      const bool active = part * 32 < nb;       // (warp-uniform) columns past nb are neither written nor read by the MMAs
      float pf[32];
      mbar_wait(bar_s, i & 1);
      tc_fence_after();
      if (active) {
        uint32_t s[32];
        tmem_ld_32x32(t_lane + Cfg::TM_S + part * 32, s);
        tmem_ld_wait();
        const float4* nl4 = reinterpret_cast<const float4*>(st + part * 32);   // -lse of my columns (broadcast reads)
        uint32_t pk[16];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 nl = nl4[c4];
          const int c = c4 * 4;
          pf[c] = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, nl.x));
          pf[c + 1] = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, nl.y));
          pf[c + 2] = ex2_approx(fmaf(__uint_as_float(s[c + 2]), p.scale_log2, nl.z));
          pf[c + 3] = ex2_approx(fmaf(__uint_as_float(s[c + 3]), p.scale_log2, nl.w));
          pk[c4 * 2] = pack_bf16x2(pf[c], pf[c + 1]);
          pk[c4 * 2 + 1] = pack_bf16x2(pf[c + 2], pf[c + 3]);
        }
        tmem_st_32x16(t_lane + Cfg::TM_S + part * 32, pk);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);

      mbar_wait(bar_dp, i & 1);
      tc_fence_after();
      if (active) {
        uint32_t dp[32];
        tmem_ld_32x32(t_lane + Cfg::TM_DP + part * 32, dp);
        tmem_ld_wait();
        const float4* ds4 = reinterpret_cast<const float4*>(st + 128 + part * 32);   // delta * scale
        uint32_t dk[16];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 dsc = ds4[c4];
          const int c = c4 * 4;
          dk[c4 * 2] = pack_bf16x2(pf[c] * fmaf(__uint_as_float(dp[c]), p.scale, -dsc.x),
                                   pf[c + 1] * fmaf(__uint_as_float(dp[c + 1]), p.scale, -dsc.y));
          dk[c4 * 2 + 1] = pack_bf16x2(pf[c + 2] * fmaf(__uint_as_float(dp[c + 2]), p.scale, -dsc.z),
                                       pf[c + 3] * fmaf(__uint_as_float(dp[c + 3]), p.scale, -dsc.w));
        }
        tmem_st_32x16(t_lane + Cfg::TM_DP + part * 32, dk);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ds);
    }
    mbar_wait(bar_done, 0);   // every MMA has retired and no load is in flight: the K / V tiles' shared memory stages dV / dK
    tc_fence_after();
    bwd2_stage_rows<D>(t_lane + Cfg::TM_DV, part, row, p.dvalid, sK);
    bwd2_stage_rows<D>(t_lane + Cfg::TM_DK, part, row, p.dvalid, sV);
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_BWD_CW) : "memory");
    if (warp == 0 && elect_one()) {
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        tma_store_4d(&tm_dv, sK + c * CH, c * 64, k0, h, b);
        tma_store_4d(&tm_dk, sV + c * CH, c * 64, k0, h, b);
      }
      bulk_commit_group();
      bulk_wait_group_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int D>
struct AttnBwdQ2Cfg {
  static constexpr int T_BYTES = 128 * D * 2;
  static constexpr int NK = 3;                  // K_j feeds S_j (one block ahead) and dQ_j
  static constexpr int NV = 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_DO = T_BYTES;
  static constexpr int OFF_K = 2 * T_BYTES;
  static constexpr int OFF_V = OFF_K + NK * T_BYTES;
  static constexpr int OFF_BAR = OFF_V + NV * T_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
  static constexpr int TM_S0 = 0, TM_DP = 128, TM_DQ = 256, TM_S1 = 384;
};

template <int D>
__global__ void __launch_bounds__(ATT_BWD2_THREADS, 1)
flash_bwd_dq2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                     const __grid_constant__ CUtensorMap tm_dq, const AttnBwdParams p) {
  using Cfg = AttnBwdQ2Cfg<D>;
  constexpr int DC = D / 64;
  constexpr int CH = 128 * 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sDO = smem + Cfg::OFF_DO;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_do = bars + 1;
  uint64_t* bar_k = bars + 2;    // [3]
  uint64_t* bar_v = bars + 5;    // [2]
  uint64_t* bar_s = bars + 7;    // [2]
  uint64_t* bar_dp = bars + 9;   // dP_j in TMEM (=> V slot j % 2 free)
  uint64_t* bar_ds = bars + 10;  // dS_j written (16 warps)
  uint64_t* bar_dq = bars + 11;  // dQ_j retired: K slot j % 3 is free
  uint64_t* bar_done = bars + 12;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nk = (p.Lk + 127) / 128;

  if (warp == ATT_BWD2_MMA_WARP && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("genhancer_b200: flash_bwd_dq2: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(bar_q, 1); mbar_init(bar_do, 1);
    for (int i = 0; i < Cfg::NK; ++i) mbar_init(&bar_k[i], 1);
    for (int i = 0; i < Cfg::NV; ++i) mbar_init(&bar_v[i], 1);
    mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1);
    mbar_init(bar_dp, 1);
    mbar_init(bar_ds, ATT_BWD_CW);
    mbar_init(bar_dq, 1);
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == ATT_BWD2_TMA_WARP) {
    auto load_tile = [&](uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int row0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar, Cfg::T_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) tma_load_4d(dst + c * CH, m, bar, c * 64, row0, h, b);
      }
    };
    load_tile(sQ, &tm_q, bar_q, q0);
    load_tile(sK, &tm_k, &bar_k[0], 0);
    load_tile(sDO, &tm_do, bar_do, q0);
    load_tile(sV, &tm_v, &bar_v[0], 0);
    if (nk > 1) {
      load_tile(sK + Cfg::T_BYTES, &tm_k, &bar_k[1], 128);
      load_tile(sV + Cfg::T_BYTES, &tm_v, &bar_v[1], 128);
    }
    if (nk > 2) load_tile(sK + 2 * Cfg::T_BYTES, &tm_k, &bar_k[2], 256);
    // (a parity wait can only tell the barrier's current phase from the one before it: every phase is waited for, in order)
    for (int j = 0; j + 2 < nk; ++j) {
      mbar_wait(bar_dp, j & 1);                       // dP_j complete: V slot j % 2 is free
      load_tile(sV + (j % Cfg::NV) * Cfg::T_BYTES, &tm_v, &bar_v[j % Cfg::NV], (j + 2) * 128);
      mbar_wait(bar_dq, j & 1);                       // dQ_j retired: K slot j % 3 is free
      if (j + 3 < nk) load_tile(sK + (j % Cfg::NK) * Cfg::T_BYTES, &tm_k, &bar_k[j % Cfg::NK], (j + 3) * 128);
    }
    __syncwarp();
  } else if (warp == ATT_BWD2_MMA_WARP) {
    const uint32_t idesc_a = umma_idesc_bf16(128, p.dvalid, false, true);
    const int ksteps = p.dvalid >> 4;
    const uint64_t kdesc = umma_desc_base(16u, 1024u);
    const uint64_t mdesc = umma_desc_base(static_cast<uint32_t>(CH), 1024u);   // K_j as MN-major B operand
    const uint32_t aQ = smem_u32(sQ), aD = smem_u32(sDO), aK0 = smem_u32(sK), aV0 = smem_u32(sV);
    auto idesc_blk = [&](int j) { return umma_idesc_bf16(128, bwd2_block_n(p.Lk - j * 128), false, false); };
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_k[0], 0);
    tc_fence_after();
    if (elect_one()) {
      mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_S0, kdesc, aQ, aK0, idesc_blk(0));
      umma_commit(&bar_s[0]);
    }
    mbar_wait(bar_do, 0);
    mbar_wait(&bar_v[0], 0);
    tc_fence_after();
    if (elect_one()) {
      mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_DP, kdesc, aD, aV0, idesc_blk(0));
      umma_commit(bar_dp);
    }
    if (nk > 1) {
      mbar_wait(&bar_k[1], 0);
      tc_fence_after();
      if (elect_one()) {
        mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_S1, kdesc, aQ, aK0 + Cfg::T_BYTES, idesc_blk(1));
        umma_commit(&bar_s[1]);
      }
    }
    for (int j = 0; j < nk; ++j) {
      const bool more1 = j + 1 < nk, more2 = j + 2 < nk;
      const uint32_t id1 = idesc_blk(j + 1), id2 = idesc_blk(j + 2);
      const uint32_t t_s = tmem + ((j & 1) ? Cfg::TM_S1 : Cfg::TM_S0);
      if (more1) mbar_wait(&bar_v[(j + 1) % Cfg::NV], ((j + 1) / Cfg::NV) & 1);   // (landed long ago; looked at while the pipe is busy)
      mbar_wait(bar_ds, j & 1);                            // dS_j is in TMEM over S_j; S_j and dP_j have been read
      tc_fence_after();
      if (elect_one()) {
        if (more1) {                                       // first: the element-wise warps wait for it, not for dQ_j
          mma_over_head_dim<D, CH, CH>(ksteps, tmem + Cfg::TM_DP, kdesc, aD, aV0 + ((j + 1) % Cfg::NV) * Cfg::T_BYTES, id1);
          umma_commit(bar_dp);
        }
        mma_ts_ksteps(bwd2_block_n(p.Lk - j * 128) >> 4, tmem + Cfg::TM_DQ, t_s, mdesc, aK0 + (j % Cfg::NK) * Cfg::T_BYTES, idesc_a,
                      j != 0 ? 1u : 0u);
        umma_commit(more1 ? bar_dq : bar_done);             // frees K slot j % 3 / releases the epilogue
      }
      if (more2) {
        mbar_wait(&bar_k[(j + 2) % Cfg::NK], ((j + 2) / Cfg::NK) & 1);   // (two chains are queued: this wait costs the pipe nothing)
        tc_fence_after();
        if (elect_one()) {                                 // behind dQ_j: into the buffer dS_j was read from
          mma_over_head_dim<D, CH, CH>(ksteps, t_s, kdesc, aQ, aK0 + ((j + 2) % Cfg::NK) * Cfg::T_BYTES, id2);
          umma_commit(&bar_s[j & 1]);
        }
      }
    }
    __syncwarp();
  } else {
    const int row = threadIdx.x & 127;
    const int part = threadIdx.x >> 7;  // which 32 of the 128 key columns this thread owns
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int ql = q0 + row;
    const int64_t stat = (static_cast<int64_t>(b) * p.H + h) * p.Lq + ql;
    const float nlse = ql < p.Lq ? -p.lse2[stat] : 0.f;
    const float ndls = ql < p.Lq ? -p.delta[stat] * p.scale : 0.f;     // dS = p * fma(dP, scale, -delta * scale)
    for (int j = 0; j < nk; ++j) {
      const int kv_left = p.Lk - j * 128;
      const bool active = part * 32 < bwd2_block_n(kv_left);
      float pf[32];
      mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (active) {
        uint32_t s[32];
        tmem_ld_32x32(t_lane + ((j & 1) ? Cfg::TM_S1 : Cfg::TM_S0) + part * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) pf[c] = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, nlse));
        if (kv_left < 128) {   // ragged last key block only: keys past Lk contribute nothing
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (part * 32 + c >= kv_left) pf[c] = 0.f;
        }
      }
      mbar_wait(bar_dp, j & 1);
      tc_fence_after();
      if (active) {
        uint32_t dp[32];
        tmem_ld_32x32(t_lane + Cfg::TM_DP + part * 32, dp);
        tmem_ld_wait();
        uint32_t ds[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2)
          ds[c >> 1] = pack_bf16x2(pf[c] * fmaf(__uint_as_float(dp[c]), p.scale, ndls),
                                   pf[c + 1] * fmaf(__uint_as_float(dp[c + 1]), p.scale, ndls));
        tmem_st_32x16(t_lane + ((j & 1) ? Cfg::TM_S1 : Cfg::TM_S0) + part * 32, ds);   // over the scores it came from
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ds);
    }
    mbar_wait(bar_done, 0);            // the last dQ MMA (and, in order, everything before it) has retired
    tc_fence_after();
    bwd2_stage_rows<D>(t_lane + Cfg::TM_DQ, part, row, p.dvalid, sQ);   // (the Q tile's shared memory: all MMAs have retired)
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_BWD_CW) : "memory");
    if (warp == 0 && elect_one()) {
#pragma unroll
      for (int c = 0; c < DC; ++c) tma_store_4d(&tm_dq, sQ + c * CH, c * 64, q0, h, b);
      bulk_commit_group();
      bulk_wait_group_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace gh
