// genhancer_b200 -- flash attention for sm_100a: tcgen05 MMAs with S / O accumulators in TMEM,
// Q/K/V tiles staged by TMA, online softmax by one thread per query row.
//
// Forward, one CTA per (128-query block, head, sample), two CTAs resident per SM:
//   warp 4 (one lane)  : TMA loads + tcgen05.mma issue   S_j = Q K_j^T   (double-buffered in TMEM)
//                                                        O  += P_j V_j
//   warps 0..3         : thread r owns query row r: tcgen05.ld S_j -> scale/mask/max/exp2 -> P_j (bf16) into
//                        a SWIZZLE_128B K-major smem tile (the A operand of the PV MMA); lazy rescale of the
//                        TMEM-resident O only when the running max grows by > 2^8; final O / l and LSE.
// The S_{j+1} MMA is issued before softmax_j finishes, so tensor pipe and softmax overlap inside the CTA and
// across the two co-resident CTAs.
//
// Layout contract: q/k/v element [b,h,l,:] is D contiguous bf16 at ptr + b*batch_stride + h*head_stride +
// l*row_stride (so both a head-major [B,H,L,D] buffer and the ViT's fused QKV GEMM output are zero-copy).
// O is token-major: element [b,l,h*D+d], optionally split in two row segments (txt | img streams of the
// DiT double blocks, layers.py:328).
#include "common.cuh"
#include "internal.h"

namespace gh {

using bf16 = __nv_bfloat16;
constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 64;

struct SegOut {  // two-segment token-major addressing: rows [0,n_split) -> seg0, [n_split, L) -> seg1
  bf16* p0; int64_t bs0, rs0;
  bf16* p1; int64_t bs1, rs1;
  int n_split;
  __device__ __forceinline__ bf16* row(int b, int l) const {
    return l < n_split ? p0 + b * bs0 + static_cast<int64_t>(l) * rs0
                       : p1 + b * bs1 + static_cast<int64_t>(l - n_split) * rs1;
  }
};

struct AttnFwdParams {
  int B, H, Lq, Lk;
  float scale_log2;  // softmax scale * log2(e)
  // Leading lanes of the head dim that hold data (multiple of 16, <= D).  SigLIP's 72-wide and MetaCLIP-H's 80-wide
  // heads live zero-padded in 128-lane slots: the MMAs whose K runs over the head dim issue only dvalid / 16 of their D / 16
  // k-steps, the MMAs whose N is the head dim run with N = dvalid, and the pad lanes of the output are written as zeros
  // without being computed -- 5/8 of the tensor work of the full 128-lane form at dvalid = 80.
  int dvalid;
  SegOut o;
  float* lse2;       // [B, H, Lq]  log2-domain logsumexp of the scaled scores
};

#ifndef GH_FWD64_KSTAGES
#define GH_FWD64_KSTAGES 1
#endif
template <int D>
struct AttnFwdCfg {
  static constexpr int Q_BYTES = ATT_BQ * D * 2;
  static constexpr int K_BYTES = ATT_BKV * D * 2;
  static constexpr int V_BYTES = ATT_BKV * D * 2;
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;                 // K_STAGES stages (defined below)
  static constexpr int OFF_V = OFF_K + (D == 64 ? GH_FWD64_KSTAGES : 2) * K_BYTES;     // 1 stage
  static constexpr int OFF_P = OFF_V + V_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  // D = 128: S double-buffered (2 x 64 columns) + O (128) = 256 columns, two CTAs per SM.
  // D = 64 (the ViT towers): the softmax, not the tensor pipe, is the limit (8192 MUFU.EX2 per 64-key block = 512
  // cycles of the SM's MUFU against 256 of MMA), and with two CTAs per SM the MUFU idled half the time while a CTA
  // sat in its barrier chain.  S single-buffered + O = 128 columns lets THREE CTAs share an SM (384 of 512 TMEM
  // columns, 3 x 57 KB of shared memory): what a CTA loses by not overlapping S_{j+1} with softmax_j the third CTA
  // fills in.
#ifdef GH_ATTN_FWD_TWO_CTAS   // A/B builds: the double-buffered-S, two-CTA form for every head dim
  static constexpr bool LEAN64 = false;
#else
  static constexpr bool LEAN64 = D == 64;
#endif
  static constexpr int S_BUFS = LEAN64 ? 1 : 2;
  static constexpr int TMEM_COLS = LEAN64 ? 128 : 256;
  static constexpr int TM_S = 0, TM_O = LEAN64 ? 64 : 128;
  // D = 64 with ONE K stage (K_{j+1} is requested when S_j has retired, i.e. during softmax_j): 48 KB of shared memory and
  // 128 TMEM columns per CTA -> FOUR CTAs per SM (A/B: -DGH_FWD64_KSTAGES=2 is the three-CTA form)
  static constexpr int K_STAGES = D == 64 ? GH_FWD64_KSTAGES : 2;
  static constexpr int CTAS_PER_SM = LEAN64 ? (K_STAGES == 1 ? 4 : 3) : 2;
};


// ----------------------------------------------------------------------------------------------------------
// MMA issue, cheap.  A descriptor differs from its neighbours only in the 14-bit start-address field of its LOW word,
// so a k-loop is "low word + compile-time constant" per operand.  The first version rebuilt every descriptor from the
// byte address (shift, two masks, or) and, with a per-k-step `if` for the valid-lane count, the compiler could not
// group the MMAs: 14 uniform-datapath instructions + a branch sat between consecutive UTCHMMAs, ~100 cycles each, and
// the dK/dV control warp spent ~1900 of the ~2900 cycles of a query block ISSUING its 24 MMAs (in-kernel counters,
// tools/attn_prof.py) while the tensor pipe's math was 1024 -- the kernels were bound by one warp's instruction
// stream.  Here every group is fully unrolled for its k-step count (a switch outside the group, no branch inside).
//   A_CH / B_CH : byte distance between 64-wide chunks of the reduction dim (K-major operands wider than 64)
//   A_ST / B_ST : byte step per 16-deep k-step inside a chunk (32 K-major, 2048 MN-major)
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_join(uint32_t hi, uint32_t lo) { return (static_cast<uint64_t>(hi) << 32) | lo; }

template <int KS, int A_CH, int A_ST, int B_CH, int B_ST>
__device__ __forceinline__ void mma_group(uint32_t tmem_d, uint64_t a_base, uint32_t a_addr, uint64_t b_base, uint32_t b_addr,
                                          uint32_t idesc, uint32_t acc_first) {
  const uint32_t a_hi = static_cast<uint32_t>(a_base >> 32), b_hi = static_cast<uint32_t>(b_base >> 32);
  const uint32_t a_lo = static_cast<uint32_t>(a_base) | ((a_addr >> 4) & 0x3FFFu);
  const uint32_t b_lo = static_cast<uint32_t>(b_base) | ((b_addr >> 4) & 0x3FFFu);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint32_t ao = static_cast<uint32_t>(((ks >> 2) * A_CH + (ks & 3) * A_ST) >> 4);
    const uint32_t bo = static_cast<uint32_t>(((ks >> 2) * B_CH + (ks & 3) * B_ST) >> 4);
    umma_ss(tmem_d, desc_join(a_hi, a_lo + ao), desc_join(b_hi, b_lo + bo), idesc, ks == 0 ? acc_first : 1u);
  }
}
// reduction over the head dim: ksteps = (valid lanes) / 16, one of {D/16, ...}; unrolled per count
template <int D, int A_CH, int B_CH>
__device__ __forceinline__ void mma_over_head_dim(int ksteps, uint32_t tmem_d, uint64_t kdesc, uint32_t a_addr, uint32_t b_addr,
                                                  uint32_t idesc) {
  switch (ksteps) {
    case 1: mma_group<1, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
    case 2: mma_group<2, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
    case 3: mma_group<3, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
    case 4: mma_group<4, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
    case 5: if (D > 64) mma_group<5, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
    case 6: if (D > 64) mma_group<6, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
    case 7: if (D > 64) mma_group<7, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
    default: if (D > 64) mma_group<8, A_CH, 32, B_CH, 32>(tmem_d, kdesc, a_addr, kdesc, b_addr, idesc, 0u); break;
  }
}

// write 8 bf16 (16 B) of row r, logical 16-byte chunk c, into a SWIZZLE_128B K-major tile (128 B rows)
__device__ __forceinline__ void st_sw128(uint8_t* tile, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

template <int D>
__global__ void __launch_bounds__(160, AttnFwdCfg<D>::CTAS_PER_SM)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  using Cfg = AttnFwdCfg<D>;
  constexpr int DC = D / 64;  // 64-wide chunks of the head dim
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sP = smem + Cfg::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;  // [2]
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;  // [2]
  uint64_t* bar_p = bars + 6;
  uint64_t* bar_o = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Lk + ATT_BKV - 1) / ATT_BKV;

  if (warp == 4 && lane == 0) {
    mbar_init(bar_q, 1);
    mbar_init(&bar_k[0], 1); mbar_init(&bar_k[1], 1);
    mbar_init(bar_v, 1);
    mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 4) {
    {
      // ---------------- control warp: TMA + MMA issue ----------------
      // All 32 lanes run the control flow (waits, loop counters); every async-unit instruction is issued by one
      // ELECTED lane (elect.sync), so its operands sit in uniform registers -- under `lane == 0` the compiler wraps
      // each UTCHMMA / UTMALDG / UTCBAR in an ELECT/R2UR/BRA.U.ANY waterfall (~100+ cycles per instruction).
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, ATT_BKV, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(128, p.dvalid, false, true);   // N = the lanes that hold data
      const int ksteps = p.dvalid >> 4;                                       // k-steps of the MMAs that reduce over the head dim
      const uint64_t kdesc = umma_desc_base(16u, 1024u);      // K-major tiles (Q, K, P)
      const uint64_t vdesc = umma_desc_base(8192u, 1024u);    // V as MN-major B operand
      auto load_k = [&](int j) {
        if (elect_one()) {
          uint8_t* dst = sK + (j % Cfg::K_STAGES) * Cfg::K_BYTES;
          mbar_arrive_expect_tx(&bar_k[j % Cfg::K_STAGES], Cfg::K_BYTES);
  #pragma unroll
          for (int c = 0; c < DC; ++c) tma_load_4d(dst + c * 8192, &tm_k, &bar_k[j % Cfg::K_STAGES], c * 64, j * ATT_BKV, h, b);
        }
      };
      auto load_v = [&](int j) {
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_v, Cfg::V_BYTES);
  #pragma unroll
          for (int c = 0; c < DC; ++c) tma_load_4d(sV + c * 8192, &tm_v, bar_v, c * 64, j * ATT_BKV, h, b);
        }
      };
      auto issue_s = [&](int j) {
        if (elect_one()) {
          const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK + (j % Cfg::K_STAGES) * Cfg::K_BYTES);
          mma_over_head_dim<D, 16384, 8192>(ksteps, tmem + Cfg::TM_S + (Cfg::S_BUFS == 2 ? (j & 1) * ATT_BKV : 0), kdesc, aQ, aK,
                                            idesc_s);
          umma_commit(&bar_s[j & 1]);
        }
      };
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) tma_load_4d(sQ + c * 16384, &tm_q, bar_q, c * 64, q0, h, b);
      }
      load_k(0);
      load_v(0);
      if (Cfg::K_STAGES == 2 && nkv > 1) load_k(1);
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (Cfg::K_STAGES == 1 && j + 1 < nkv) {
          mbar_wait(&bar_s[j & 1], (j >> 1) & 1);   // S_j retired: the one K buffer is free (softmax_j runs meanwhile)
          load_k(j + 1);
        }
        if (Cfg::S_BUFS == 2 && j + 1 < nkv) {
          // S buffer (j+1)&1 was last read by softmax_{j-1}, which arrived on bar_p before we got here
          mbar_wait(&bar_k[(j + 1) % Cfg::K_STAGES], ((j + 1) / Cfg::K_STAGES) & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(bar_p, j & 1);  // P_j is in smem (and O has been rescaled if needed); softmax_j has read S_j
        if (Cfg::S_BUFS == 1 && j + 1 < nkv) {
          // single S buffer: S_{j+1} may overwrite S_j now; issued BEFORE PV_j so that softmax_{j+1} starts earlier
          mbar_wait(&bar_k[(j + 1) % Cfg::K_STAGES], ((j + 1) / Cfg::K_STAGES) & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
          mma_group<4, 0, 32, 0, 2048>(tmem + Cfg::TM_O, kdesc, aP, vdesc, aV, idesc_o, j != 0 ? 1u : 0u);
          umma_commit(bar_o);
        }
        if (Cfg::K_STAGES == 2 && j + 2 < nkv) load_k(j + 2);  // K buffer j&1 is free: S_j completed before softmax_j started
        if (j + 1 < nkv) {
          mbar_wait(bar_o, j & 1);       // PV_j done -> V (single buffer) and P are free
          load_v(j + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- softmax / correction / epilogue: thread = query row ----------------
    const int row = threadIdx.x;  // 0..127 == TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float m_used = 0.f, l_sum = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[64];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        tmem_ld_32x32(t_lane + Cfg::TM_S + (Cfg::S_BUFS == 2 ? (j & 1) * ATT_BKV : 0), s0);
        tmem_ld_32x32(t_lane + Cfg::TM_S + (Cfg::S_BUFS == 2 ? (j & 1) * ATT_BKV : 0) + 32, s1);
        tmem_ld_wait();
      }
      // ~300 issue slots per 64-key block and thread (it was ~900: a separate scale multiply, a per-element tail mask,
      // exp2f's denormal range handling, and a bf16 round trip for the row sum): the tail mask runs only in the last
      // block, the scale rides in the FFMA that subtracts the running maximum, MUFU.EX2 is issued bare.
      const int kv_left = p.Lk - j * ATT_BKV;  // valid keys in this block
      if (kv_left < ATT_BKV) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= kv_left) s[c] = 0xff800000u;  // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], __uint_as_float(s[c + u]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;  // scale > 0
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
      const float rs = rs0 + rs1;
      l_sum = l_sum * factor + rs;
      if (j > 0) {
        mbar_wait(bar_o, (j - 1) & 1);  // PV_{j-1} retired: P buffer reusable, O stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, factor != 1.f)) {
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            if (c * 32 >= p.dvalid) break;     // (columns >= dvalid were never written by the PV MMA)
            uint32_t o[32];
            tmem_ld_32x32(t_lane + Cfg::TM_O + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x32(t_lane + Cfg::TM_O + c * 32, o);
          }
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        st_sw128(sP, row, c, make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue ----
    mbar_wait(bar_o, (nkv - 1) & 1);
    tc_fence_after();
    const int l = q0 + row;
    const float inv = 1.f / l_sum;
    if (l < p.Lq && p.lse2) p.lse2[(static_cast<int64_t>(b) * p.H + h) * p.Lq + l] = m_used + log2f(l_sum);
    bf16* orow = (l < p.Lq) ? p.o.row(b, l) + h * D : nullptr;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t o[32];
      if (c * 32 < p.dvalid) {                 // (a warp-uniform branch: tcgen05.ld is .sync.aligned)
        tmem_ld_32x32(t_lane + Cfg::TM_O + c * 32, o);
        tmem_ld_wait();
      }
      if (orow) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w = make_uint4(0u, 0u, 0u, 0u);          // pad lanes: exact zeros
          if (c * 32 + i * 8 < p.dvalid) {
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

// ============================================================================
// Backward
//   prep : delta[b,h,l] = sum_d dO*O ; dO gathered token-major(2 segments) -> head-major [B,H,L,D]
//   dKV  : CTA = (128 keys, h, b), loops over 64-query blocks.  Works on the TRANSPOSED score tile so that the
//          thread-owned TMEM lane is a key row and P^T / dS^T land in smem K-major without any transpose:
//            S^T = K Q^T, dP^T = V dO^T, P^T = exp2(S^T c - lse_col), dS^T = P^T (dP^T - delta_col) scale
//            dV += P^T dO,  dK += dS^T Q          (Q / dO tiles are reused as MN-major B operands)
//   dQ   : CTA = (128 queries, h, b), loops over 64-key blocks:  S = Q K^T, dP = dO V^T, dQ += dS K
// 7 GEMMs instead of the minimal 5 (S and dP are recomputed in the dQ pass), no atomics, deterministic.
// ============================================================================
struct AttnBwdParams {
  int B, H, Lq, Lk;
  float scale, scale_log2;
  int dvalid;          // valid leading lanes of the head dim (see AttnFwdParams)
  // bring-up aid (gh_debug_attn_prof): cycle counters of CTA (0,0,0) of the dK/dV kernel, NULL in production.
  //   compute warp 0 / lane 0, summed over the query blocks: [0] waiting for S/dP  [1] bar.sync + TMEM loads  [2] math
  //   [3] waiting for the previous dV/dK MMAs  [4] smem stores + fence + arrive  [5] whole loop  [6] epilogue
  //   control warp: [8] waiting for Q/dO tiles  [9] waiting for P^T/dS^T  [10] whole loop  [11] K/V load wait
  long long* prof;
  const float* lse2;   // [B,H,Lq]
  const float* delta;  // [B,H,Lq]
  bf16 *dq, *dk, *dv;  // element [b,h,l,:] at ptr + b*bs + h*hs + l*rs
  int64_t dq_bs, dq_hs, dq_rs, dk_bs, dk_hs, dk_rs, dv_bs, dv_hs, dv_rs;
};

// One lane = 8 consecutive head lanes (16 bytes) of one (token, head) row: D / 8 lanes per row, 256 / D rows per warp.  (The first
// version gave a whole warp to a row -- 8-byte accesses for D = 128, 4-byte for D = 64 -- and was issue-bound at 2.6-3.7 TB/s.)
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(SegOut o, SegOut dout, int B, int H, int L,
                                                            float* __restrict__ delta, bf16* __restrict__ do_hm) {
  constexpr int LPR = D / 8;          // lanes per row (16 or 8)
  constexpr int RPW = 32 / LPR;       // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t rid = wid * RPW + lane / LPR;                     // (b, l, h) row, h fastest
  const bool ok = rid < static_cast<int64_t>(B) * L * H;
  float acc = 0.f;
  int h = 0, l = 0, b = 0;
  if (ok) {
    h = static_cast<int>(rid % H);
    const int64_t tok = rid / H;
    l = static_cast<int>(tok % L);
    b = static_cast<int>(tok / L);
    const uint4 a = *reinterpret_cast<const uint4*>(o.row(b, l) + h * D + sub * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(dout.row(b, l) + h * D + sub * 8);
    *reinterpret_cast<uint4*>(do_hm + ((static_cast<int64_t>(b) * H + h) * L + l) * D + sub * 8) = g;
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 x = unpack_bf16x2(aw[q]), y = unpack_bf16x2(gw[q]);
      acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
    }
  }
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (ok && sub == 0) delta[(static_cast<int64_t>(b) * H + h) * L + l] = acc;
}

// Backward kernels: 16 compute warps (thread = a QUARTER of a score row: 16 columns) + 1 control warp.  With one CTA per
// SM (TMEM: 2 x (S, dP) + dV + dK = 512 columns) the element-wise part is a chain of TMEM loads, MUFU, shared-memory
// stores and barrier round trips whose latencies only other warps can hide: the first version had 8 compute warps
// (half a row each) and ncu showed it latency-bound on everything at once (issue slots 22 % busy, tensor pipe 29 %,
// no single hot instruction: profiles/r02_kernels_ncu.txt).
constexpr int ATT_BWD_SPLIT = 4;                       // threads per row
constexpr int ATT_BWD_CPT = 64 / ATT_BWD_SPLIT;        // columns per thread
constexpr int ATT_BWD_CW = 4 * ATT_BWD_SPLIT;          // compute warps (also the index of the control warp)
constexpr int ATT_BWD_THREADS = 32 * (ATT_BWD_CW + 1);

#ifdef GH_ATTN_BWD_V1   // the first form of the backward (128 x 64 tiles, every operand in shared memory): A/B builds only
template <int D>
struct AttnBwdKVCfg {
  static constexpr int KV_BYTES = 128 * D * 2;  // K_j or V_j (resident)
  static constexpr int QD_BYTES = 64 * D * 2;   // Q_i or dO_i
  static constexpr int NQ = 3;                  // Q/dO ring depth
  static constexpr int T_BYTES = 128 * 64 * 2;  // P^T or dS^T
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = KV_BYTES;
  static constexpr int OFF_Q = 2 * KV_BYTES;
  static constexpr int OFF_DO = OFF_Q + NQ * QD_BYTES;
  static constexpr int OFF_PT = OFF_DO + NQ * QD_BYTES;
  static constexpr int OFF_DST = OFF_PT + T_BYTES;
  static constexpr int OFF_STAT = OFF_DST + T_BYTES;     // float [2][2][64]: lse, delta per S buffer
  static constexpr int OFF_BAR = OFF_STAT + 2 * 2 * 64 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int TM_DV = 256, TM_DK = 256 + D;
};

template <int D>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
flash_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                     const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                     const AttnBwdParams p) {
  using Cfg = AttnBwdKVCfg<D>;
  constexpr int DC = D / 64;
  constexpr int NQ = Cfg::NQ;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sDO = smem + Cfg::OFF_DO;
  uint8_t* sPT = smem + Cfg::OFF_PT;
  uint8_t* sDST = smem + Cfg::OFF_DST;
  float* sStat = reinterpret_cast<float*>(smem + Cfg::OFF_STAT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_qd = bars + 1;   // [3]
  uint64_t* bar_sd = bars + 4;   // [2]
  uint64_t* bar_pd = bars + 6;
  uint64_t* bar_acc = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nq = (p.Lq + 63) / 64;

  // warps 0..15: compute (warps w, w + 4, w + 8, w + 12 share TMEM lane quadrant w % 4; each thread owns a QUARTER of
  // its row's 64 columns); warp 16: control
  if (warp == ATT_BWD_CW && lane == 0) {
    mbar_init(bar_kv, 1);
    for (int i = 0; i < NQ; ++i) mbar_init(&bar_qd[i], 1);
    mbar_init(&bar_sd[0], 1); mbar_init(&bar_sd[1], 1);
    mbar_init(bar_pd, ATT_BWD_CW);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == ATT_BWD_CW) {
    {  // control warp: see flash_fwd_kernel (all lanes run the flow, one elected lane issues)
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
      const uint32_t idesc_a = umma_idesc_bf16(128, p.dvalid, false, true);   // accumulators: N = the lanes that hold data
      const int ksteps = p.dvalid >> 4;                                       // k-steps of the MMAs that reduce over the head dim
      const uint64_t kdesc = umma_desc_base(16u, 1024u);
      const uint64_t mdesc = umma_desc_base(8192u, 1024u);
      auto load_qd = [&](int i) {
        if (elect_one()) {
          const int s = i % NQ;
          mbar_arrive_expect_tx(&bar_qd[s], 2 * Cfg::QD_BYTES);
  #pragma unroll
          for (int c = 0; c < DC; ++c) {
            tma_load_4d(sQ + s * Cfg::QD_BYTES + c * 8192, &tm_q, &bar_qd[s], c * 64, i * 64, h, b);
            tma_load_4d(sDO + s * Cfg::QD_BYTES + c * 8192, &tm_do, &bar_qd[s], c * 64, i * 64, h, b);
          }
        }
      };
      auto issue_sd = [&](int i) {
        if (elect_one()) {
          const int s = i % NQ;
          const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
          const uint32_t aQ = smem_u32(sQ + s * Cfg::QD_BYTES), aD = smem_u32(sDO + s * Cfg::QD_BYTES);
          const uint32_t tS = tmem + (i & 1) * 128, tP = tS + 64;
          mma_over_head_dim<D, 16384, 8192>(ksteps, tS, kdesc, aK, aQ, idesc_s);
          mma_over_head_dim<D, 16384, 8192>(ksteps, tP, kdesc, aV, aD, idesc_s);
          umma_commit(&bar_sd[i & 1]);
        }
      };
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_kv, 2 * Cfg::KV_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          tma_load_4d(sK + c * 16384, &tm_k, bar_kv, c * 64, k0, h, b);
          tma_load_4d(sV + c * 16384, &tm_v, bar_kv, c * 64, k0, h, b);
        }
      }
      const bool prof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      long long w_qd = 0, w_pd = 0, t_kv = 0;
      const long long c_begin = prof ? clock64() : 0;
      load_qd(0);
      if (nq > 1) load_qd(1);
      mbar_wait(bar_kv, 0);
      mbar_wait(&bar_qd[0], 0);
      if (prof) t_kv = clock64() - c_begin;
      tc_fence_after();
      issue_sd(0);
      for (int i = 0; i < nq; ++i) {
        if (i + 1 < nq) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(&bar_qd[(i + 1) % NQ], ((i + 1) / NQ) & 1);
          if (prof) w_qd += clock64() - t0;
          tc_fence_after();
          issue_sd(i + 1);
        }
        const long long t1 = prof ? clock64() : 0;
        mbar_wait(bar_pd, i & 1);  // P^T_i, dS^T_i in smem; implies acc MMAs of i-1 have retired
        if (prof) w_pd += clock64() - t1;
        tc_fence_after();
        if (i + 2 < nq) load_qd(i + 2);  // ring slot (i+2)%3 == (i-1)%3 is free
        if (elect_one()) {
          const int s = i % NQ;
          const uint32_t aPT = smem_u32(sPT), aDST = smem_u32(sDST);
          const uint32_t aQ = smem_u32(sQ + s * Cfg::QD_BYTES), aD = smem_u32(sDO + s * Cfg::QD_BYTES);
          mma_group<4, 0, 32, 0, 2048>(tmem + Cfg::TM_DV, kdesc, aPT, mdesc, aD, idesc_a, i != 0 ? 1u : 0u);
          mma_group<4, 0, 32, 0, 2048>(tmem + Cfg::TM_DK, kdesc, aDST, mdesc, aQ, idesc_a, i != 0 ? 1u : 0u);
          umma_commit(bar_acc);
        }
      }
      if (prof && lane == 0) {
        p.prof[8] = w_qd; p.prof[9] = w_pd; p.prof[10] = clock64() - c_begin; p.prof[11] = t_kv;
      }
    }
    __syncwarp();
  } else {
    const int row = threadIdx.x & 127;  // key row within the block == TMEM lane
    const int part = threadIdx.x >> 7;  // which 16 of the 64 query columns this thread owns
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int64_t stat_base = (static_cast<int64_t>(b) * p.H + h) * p.Lq;
    // Per-column statistics of a query block, staged in smem as (-lse, delta * scale) so that the inner loop is
    // p = ex2(fma(s, c, -lse)) ; dS = p * fma(dP, scale, -delta * scale): threads 0-63 own lse, 64-127 delta.  The
    // global loads of block i + 1 are issued BEFORE the math of block i (they used to sit, dependent, at the top of every
    // iteration: one exposed L2 round trip per query block with a single CTA per SM to hide it).
    // (the loaded value is kept RAW in a register and only negated / scaled when it is stored one iteration later: a
    // thread stalls at the first USE of a load, so doing the arithmetic at load time put the L2 round trip back at the
    // top of every iteration -- 3 % of all stall samples on that one FMUL, plus the 256-thread barrier behind it)
    auto stat_load = [&](int i) -> float {
      const int ql = i * 64 + (row & 63);
      if (threadIdx.x >= 128 || i >= nq || ql >= p.Lq) return 0.f;
      return row < 64 ? p.lse2[stat_base + ql] : p.delta[stat_base + ql];
    };
    float stat_raw = stat_load(0);
    const bool prof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    long long pw_s = 0, pw_ld = 0, pw_math = 0, pw_acc = 0, pw_st = 0;
    const long long p_begin = prof ? clock64() : 0;
    for (int i = 0; i < nq; ++i) {
      float* st = sStat + (i & 1) * 128;
      if (threadIdx.x < 128) st[row] = row < 64 ? -stat_raw : stat_raw * p.scale;
      stat_raw = stat_load(i + 1);
      long long tp0 = prof ? clock64() : 0;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_BWD_CW) : "memory");
      long long tp1 = prof ? clock64() : 0;
      mbar_wait(&bar_sd[i & 1], (i >> 1) & 1);
      if (prof) { pw_s += clock64() - tp1; pw_ld += tp1 - tp0; tp0 = clock64(); }
      tc_fence_after();
      const int q_left = p.Lq - i * 64;
      constexpr int CPT = ATT_BWD_CPT;
      uint32_t pt[CPT / 2], dst[CPT / 2];
      {
        uint32_t s[CPT], dp[CPT];
        tmem_ld_32x16(t_lane + (i & 1) * 128 + part * CPT, s);
        tmem_ld_32x16(t_lane + (i & 1) * 128 + 64 + part * CPT, dp);
        tmem_ld_wait();
        const float4* nl4 = reinterpret_cast<const float4*>(st + part * CPT);        // -lse of my columns (broadcast reads)
        const float4* ds4 = reinterpret_cast<const float4*>(st + 64 + part * CPT);   // delta * scale
#pragma unroll
        for (int c4 = 0; c4 < CPT / 4; ++c4) {
          const float4 nl = nl4[c4], dsc = ds4[c4];
          const int c = c4 * 4;
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, nl.x));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, nl.y));
          const float p2 = ex2_approx(fmaf(__uint_as_float(s[c + 2]), p.scale_log2, nl.z));
          const float p3 = ex2_approx(fmaf(__uint_as_float(s[c + 3]), p.scale_log2, nl.w));
          pt[c4 * 2] = pack_bf16x2(p0, p1);
          pt[c4 * 2 + 1] = pack_bf16x2(p2, p3);
          dst[c4 * 2] = pack_bf16x2(p0 * fmaf(__uint_as_float(dp[c]), p.scale, -dsc.x),
                                    p1 * fmaf(__uint_as_float(dp[c + 1]), p.scale, -dsc.y));
          dst[c4 * 2 + 1] = pack_bf16x2(p2 * fmaf(__uint_as_float(dp[c + 2]), p.scale, -dsc.z),
                                        p3 * fmaf(__uint_as_float(dp[c + 3]), p.scale, -dsc.w));
        }
        if (q_left < 64) {   // ragged last query block only: columns past Lq contribute nothing
#pragma unroll
          for (int c = 0; c < CPT; c += 2) {
            const int cc = part * CPT + c;
            if (cc + 1 >= q_left) {
              const uint32_t keep = cc < q_left ? 0x0000ffffu : 0u;
              pt[c >> 1] &= keep;
              dst[c >> 1] &= keep;
            }
          }
        }
      }
      if (prof) { tp1 = clock64(); pw_math += tp1 - tp0; }
      if (i > 0) mbar_wait(bar_acc, (i - 1) & 1);  // previous dV/dK MMAs no longer read sPT/sDST
      if (prof) { tp0 = clock64(); pw_acc += tp0 - tp1; }
#pragma unroll
      for (int c = 0; c < CPT / 8; ++c) {
        st_sw128(sPT, row, part * (CPT / 8) + c, make_uint4(pt[4 * c], pt[4 * c + 1], pt[4 * c + 2], pt[4 * c + 3]));
        st_sw128(sDST, row, part * (CPT / 8) + c, make_uint4(dst[4 * c], dst[4 * c + 1], dst[4 * c + 2], dst[4 * c + 3]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pd);
      if (prof) pw_st += clock64() - tp0;
    }
    const long long p_loop = prof ? clock64() - p_begin : 0;
    mbar_wait(bar_acc, (nq - 1) & 1);
    tc_fence_after();
    const int kl = k0 + row;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* dstp = which ? p.dk + b * p.dk_bs + h * p.dk_hs + static_cast<int64_t>(kl) * p.dk_rs
                         : p.dv + b * p.dv_bs + h * p.dv_hs + static_cast<int64_t>(kl) * p.dv_rs;
#pragma unroll
      for (int c = part; c < D / 32; c += ATT_BWD_SPLIT) {   // the threads of a row take alternate 32-column chunks
        uint32_t o[32];
        if (c * 32 < p.dvalid) {                     // (warp-uniform: part is per warp)
          tmem_ld_32x32(t_lane + (which ? Cfg::TM_DK : Cfg::TM_DV) + c * 32, o);
          tmem_ld_wait();
        }
        if (kl < p.Lk) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 w = make_uint4(0u, 0u, 0u, 0u);   // pad lanes: exact zeros
            if (c * 32 + q4 * 8 < p.dvalid) {
              w.x = pack_bf16x2(__uint_as_float(o[8 * q4 + 0]), __uint_as_float(o[8 * q4 + 1]));
              w.y = pack_bf16x2(__uint_as_float(o[8 * q4 + 2]), __uint_as_float(o[8 * q4 + 3]));
              w.z = pack_bf16x2(__uint_as_float(o[8 * q4 + 4]), __uint_as_float(o[8 * q4 + 5]));
              w.w = pack_bf16x2(__uint_as_float(o[8 * q4 + 6]), __uint_as_float(o[8 * q4 + 7]));
            }
            *reinterpret_cast<uint4*>(dstp + c * 32 + q4 * 8) = w;
          }
        }
      }
    }
    if (prof) {
      p.prof[0] = pw_s; p.prof[1] = pw_ld; p.prof[2] = pw_math; p.prof[3] = pw_acc; p.prof[4] = pw_st; p.prof[5] = p_loop;
      p.prof[6] = clock64() - p_begin - p_loop; p.prof[7] = nq;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int D>
struct AttnBwdQCfg {
  static constexpr int Q_BYTES = 128 * D * 2;  // Q_i or dO_i (resident)
  static constexpr int KV_BYTES = 64 * D * 2;  // K_j or V_j
  static constexpr int NK = 3;
  static constexpr int T_BYTES = 128 * 64 * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_DO = Q_BYTES;
  static constexpr int OFF_K = 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + NK * KV_BYTES;
  static constexpr int OFF_DS = OFF_V + NK * KV_BYTES;
  static constexpr int OFF_BAR = OFF_DS + T_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int TM_DQ = 256;
};

template <int D>
__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
flash_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                    const AttnBwdParams p) {
  using Cfg = AttnBwdQCfg<D>;
  constexpr int DC = D / 64;
  constexpr int NK = Cfg::NK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sDO = smem + Cfg::OFF_DO;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sDS = smem + Cfg::OFF_DS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv = bars + 1;   // [3]
  uint64_t* bar_sd = bars + 4;   // [2]
  uint64_t* bar_pd = bars + 6;
  uint64_t* bar_acc = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Lk + 63) / 64;

  // warps 0..15: compute (warps w, w + 4, w + 8, w + 12 share TMEM lane quadrant w % 4; each thread owns a QUARTER of
  // its row's 64 columns); warp 16: control
  if (warp == ATT_BWD_CW && lane == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < NK; ++i) mbar_init(&bar_kv[i], 1);
    mbar_init(&bar_sd[0], 1); mbar_init(&bar_sd[1], 1);
    mbar_init(bar_pd, ATT_BWD_CW);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == ATT_BWD_CW) {
    {  // control warp: see flash_fwd_kernel (all lanes run the flow, one elected lane issues)
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, false, false);
      const uint32_t idesc_a = umma_idesc_bf16(128, p.dvalid, false, true);   // accumulators: N = the lanes that hold data
      const int ksteps = p.dvalid >> 4;                                       // k-steps of the MMAs that reduce over the head dim
      const uint64_t kdesc = umma_desc_base(16u, 1024u);
      const uint64_t mdesc = umma_desc_base(8192u, 1024u);
      auto load_kv = [&](int j) {
        if (elect_one()) {
          const int s = j % NK;
          mbar_arrive_expect_tx(&bar_kv[s], 2 * Cfg::KV_BYTES);
  #pragma unroll
          for (int c = 0; c < DC; ++c) {
            tma_load_4d(sK + s * Cfg::KV_BYTES + c * 8192, &tm_k, &bar_kv[s], c * 64, j * 64, h, b);
            tma_load_4d(sV + s * Cfg::KV_BYTES + c * 8192, &tm_v, &bar_kv[s], c * 64, j * 64, h, b);
          }
        }
      };
      auto issue_sd = [&](int j) {
        if (elect_one()) {
          const int s = j % NK;
          const uint32_t aQ = smem_u32(sQ), aD = smem_u32(sDO);
          const uint32_t aK = smem_u32(sK + s * Cfg::KV_BYTES), aV = smem_u32(sV + s * Cfg::KV_BYTES);
          const uint32_t tS = tmem + (j & 1) * 128, tP = tS + 64;
          mma_over_head_dim<D, 16384, 8192>(ksteps, tS, kdesc, aQ, aK, idesc_s);
          mma_over_head_dim<D, 16384, 8192>(ksteps, tP, kdesc, aD, aV, idesc_s);
          umma_commit(&bar_sd[j & 1]);
        }
      };
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, 2 * Cfg::Q_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          tma_load_4d(sQ + c * 16384, &tm_q, bar_q, c * 64, q0, h, b);
          tma_load_4d(sDO + c * 16384, &tm_do, bar_q, c * 64, q0, h, b);
        }
      }
      load_kv(0);
      if (nkv > 1) load_kv(1);
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[0], 0);
      tc_fence_after();
      issue_sd(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {
          mbar_wait(&bar_kv[(j + 1) % NK], ((j + 1) / NK) & 1);
          tc_fence_after();
          issue_sd(j + 1);
        }
        mbar_wait(bar_pd, j & 1);
        tc_fence_after();
        if (j + 2 < nkv) load_kv(j + 2);
        if (elect_one()) {
          const uint32_t aDS = smem_u32(sDS), aK = smem_u32(sK + (j % NK) * Cfg::KV_BYTES);
          mma_group<4, 0, 32, 0, 2048>(tmem + Cfg::TM_DQ, kdesc, aDS, mdesc, aK, idesc_a, j != 0 ? 1u : 0u);
          umma_commit(bar_acc);
        }
      }
    }
    __syncwarp();
  } else {
    const int row = threadIdx.x & 127;
    const int part = threadIdx.x >> 7;  // which 16 of the 64 key columns this thread owns
    constexpr int CPT = ATT_BWD_CPT;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int ql = q0 + row;
    const int64_t stat = (static_cast<int64_t>(b) * p.H + h) * p.Lq + ql;
    const float nlse = ql < p.Lq ? -p.lse2[stat] : 0.f;
    const float ndls = ql < p.Lq ? -p.delta[stat] * p.scale : 0.f;     // -delta * scale: dS = p * fma(dP, scale, -delta * scale)
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&bar_sd[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int kv_left = p.Lk - j * 64;
      uint32_t ds[CPT / 2];
      {
        uint32_t s[CPT], dp[CPT];
        tmem_ld_32x16(t_lane + (j & 1) * 128 + part * CPT, s);
        tmem_ld_32x16(t_lane + (j & 1) * 128 + 64 + part * CPT, dp);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < CPT; c += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[c]), p.scale_log2, nlse));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 1]), p.scale_log2, nlse));
          ds[c >> 1] = pack_bf16x2(p0 * fmaf(__uint_as_float(dp[c]), p.scale, ndls),
                                   p1 * fmaf(__uint_as_float(dp[c + 1]), p.scale, ndls));
        }
        if (kv_left < 64) {   // ragged last key block only
#pragma unroll
          for (int c = 0; c < CPT; c += 2) {
            const int cc = part * CPT + c;
            if (cc + 1 >= kv_left) ds[c >> 1] &= cc < kv_left ? 0x0000ffffu : 0u;
          }
        }
      }
      if (j > 0) mbar_wait(bar_acc, (j - 1) & 1);
#pragma unroll
      for (int c = 0; c < CPT / 8; ++c)
        st_sw128(sDS, row, part * (CPT / 8) + c, make_uint4(ds[4 * c], ds[4 * c + 1], ds[4 * c + 2], ds[4 * c + 3]));
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pd);
    }
    mbar_wait(bar_acc, (nkv - 1) & 1);
    tc_fence_after();
    bf16* dstp = p.dq + b * p.dq_bs + h * p.dq_hs + static_cast<int64_t>(ql) * p.dq_rs;
#pragma unroll
    for (int c = part; c < D / 32; c += ATT_BWD_SPLIT) {
      uint32_t o[32];
      if (c * 32 < p.dvalid) {
        tmem_ld_32x32(t_lane + Cfg::TM_DQ + c * 32, o);
        tmem_ld_wait();
      }
      if (ql < p.Lq) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 w = make_uint4(0u, 0u, 0u, 0u);   // pad lanes: exact zeros
          if (c * 32 + q4 * 8 < p.dvalid) {
            w.x = pack_bf16x2(__uint_as_float(o[8 * q4 + 0]), __uint_as_float(o[8 * q4 + 1]));
            w.y = pack_bf16x2(__uint_as_float(o[8 * q4 + 2]), __uint_as_float(o[8 * q4 + 3]));
            w.z = pack_bf16x2(__uint_as_float(o[8 * q4 + 4]), __uint_as_float(o[8 * q4 + 5]));
            w.w = pack_bf16x2(__uint_as_float(o[8 * q4 + 6]), __uint_as_float(o[8 * q4 + 7]));
          }
          *reinterpret_cast<uint4*>(dstp + c * 32 + q4 * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

#endif  // GH_ATTN_BWD_V1

}  // namespace gh
#include "attn_bwd2.cuh"
#ifdef GH_ATTN_FWD_V2
#include "attn_fwd2.cuh"
#endif
namespace gh {

// 4-D map over [b, h, l, d] with arbitrary (16-byte multiple) strides; box = 64 x box_rows x 1 x 1
static int make_qkv_map(CUtensorMap* m, const gh_attn_tensor* t, int B, int H, int L, int D, int box_rows) {
  const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(L), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(t->row_stride) * 2, static_cast<uint64_t>(t->head_stride) * 2,
                               static_cast<uint64_t>(t->batch_stride) * 2};
  const uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), 1, 1};
  return make_tmap_bf16(m, t->ptr, 4, dims, strides, box, nullptr);
}

static bool attn_tensor_ok(const gh_attn_tensor* t) {
  return t && t->ptr && aligned16(t->ptr) && t->row_stride % 8 == 0 && t->head_stride % 8 == 0 &&
         t->batch_stride % 8 == 0;
}

int attn_init() {
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnFwdCfg<64>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnFwdCfg<128>::SMEM_BYTES));
#ifdef GH_ATTN_BWD_V1
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dkv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdKVCfg<64>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dkv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdKVCfg<128>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdQCfg<64>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dq_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdQCfg<128>::SMEM_BYTES));
#endif
#ifdef GH_ATTN_FWD_V2
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnFwd2Cfg<64>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_fwd2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnFwd2Cfg<128>::SMEM_BYTES));
#endif
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dkv2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdKV2Cfg<64>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dkv2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdKV2Cfg<128>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dq2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdQ2Cfg<64>::SMEM_BYTES));
  GH_CHECK_CUDA(cudaFuncSetAttribute(flash_bwd_dq2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdQ2Cfg<128>::SMEM_BYTES));
  return GH_OK;
}

}  // namespace gh

using namespace gh;

extern "C" int gh_flash_attn_fwd(const gh_attn_tensor* q, const gh_attn_tensor* k, const gh_attn_tensor* v,
                                 int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t D, int32_t d_valid, float scale,
                                 const gh_attn_out* o, float* lse2, void* stream) {
  GH_REQUIRE(attn_tensor_ok(q) && attn_tensor_ok(k) && attn_tensor_ok(v), GH_ERR_ALIGN,
             "gh_flash_attn_fwd: q/k/v must be non-NULL, 16B aligned, strides multiples of 8 elements");
  GH_REQUIRE(o && o->seg1, GH_ERR_NULL, "gh_flash_attn_fwd: output is NULL");
  GH_REQUIRE(D == 64 || D == 128, GH_ERR_UNSUPPORTED, "gh_flash_attn_fwd: head dim %d unsupported (64, 128)", D);
  if (d_valid <= 0) d_valid = D;
  GH_REQUIRE(d_valid % 16 == 0 && d_valid <= D, GH_ERR_BAD_SHAPE, "gh_flash_attn_fwd: d_valid=%d must be a multiple of 16, <= D", d_valid);
  GH_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, GH_ERR_BAD_SHAPE, "gh_flash_attn_fwd: bad shape");
  if (B == 0 || Lq == 0) return GH_OK;
  GH_REQUIRE(o->n_split >= 0 && o->n_split <= Lq && (o->n_split == 0 || o->seg0), GH_ERR_BAD_SHAPE,
             "gh_flash_attn_fwd: bad output split");
  GH_REQUIRE(o->seg1_row_stride % 8 == 0 && o->seg1_batch_stride % 8 == 0 && aligned16(o->seg1) &&
                 (o->n_split == 0 || (o->seg0_row_stride % 8 == 0 && o->seg0_batch_stride % 8 == 0 && aligned16(o->seg0))),
             GH_ERR_ALIGN, "gh_flash_attn_fwd: output strides must be multiples of 8 elements");
#ifdef GH_ATTN_FWD_V2    // (A/B builds, tools/build_ab.sh: attn_fwd2.cuh -- 128-key blocks, P through TMEM; measured no faster)
  constexpr int kv_box = 128;
#else
  constexpr int kv_box = ATT_BKV;
#endif
  CUtensorMap mq, mk_, mv;
  if (int e = make_qkv_map(&mq, q, B, H, Lq, D, ATT_BQ)) return e;
  if (int e = make_qkv_map(&mk_, k, B, H, Lk, D, kv_box)) return e;
  if (int e = make_qkv_map(&mv, v, B, H, Lk, D, kv_box)) return e;
  AttnFwdParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.dvalid = d_valid;
  p.o.p0 = static_cast<bf16*>(o->seg0); p.o.bs0 = o->seg0_batch_stride; p.o.rs0 = o->seg0_row_stride;
  p.o.p1 = static_cast<bf16*>(o->seg1); p.o.bs1 = o->seg1_batch_stride; p.o.rs1 = o->seg1_row_stride;
  p.o.n_split = o->n_split;
  p.lse2 = lse2;
  dim3 grid((Lq + ATT_BQ - 1) / ATT_BQ, H, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#ifdef GH_ATTN_FWD_V2
  if (D == 64)
    flash_fwd2_kernel<64><<<grid, ATT_FWD2_THREADS, AttnFwd2Cfg<64>::SMEM_BYTES, s>>>(mq, mk_, mv, p);
  else
    flash_fwd2_kernel<128><<<grid, ATT_FWD2_THREADS, AttnFwd2Cfg<128>::SMEM_BYTES, s>>>(mq, mk_, mv, p);
#else
  if (D == 64)
    flash_fwd_kernel<64><<<grid, 160, AttnFwdCfg<64>::SMEM_BYTES, s>>>(mq, mk_, mv, p);
  else
    flash_fwd_kernel<128><<<grid, 160, AttnFwdCfg<128>::SMEM_BYTES, s>>>(mq, mk_, mv, p);
#endif
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}

extern "C" int64_t gh_flash_attn_bwd_workspace_bytes(int32_t B, int32_t H, int32_t Lq, int32_t D, int32_t which) {
  const int64_t rows = static_cast<int64_t>(B) * H * Lq;
  return which == 0 ? rows * D * 2 : rows * 4;
}

static long long* g_attn_prof = nullptr;
extern "C" int gh_debug_attn_prof(void* device_buf) {
  g_attn_prof = static_cast<long long*>(device_buf);
  return GH_OK;
}

static void seg_from(const gh_attn_out* o, SegOut* s) {
  s->p0 = static_cast<bf16*>(o->seg0); s->bs0 = o->seg0_batch_stride; s->rs0 = o->seg0_row_stride;
  s->p1 = static_cast<bf16*>(o->seg1); s->bs1 = o->seg1_batch_stride; s->rs1 = o->seg1_row_stride;
  s->n_split = o->n_split;
}

extern "C" int gh_flash_attn_bwd(const gh_attn_tensor* q, const gh_attn_tensor* k, const gh_attn_tensor* v,
                                 const gh_attn_out* o, const gh_attn_out* d_o, const float* lse2, int32_t B,
                                 int32_t H, int32_t Lq, int32_t Lk, int32_t D, int32_t d_valid, float scale, const gh_attn_tensor* dq,
                                 const gh_attn_tensor* dk, const gh_attn_tensor* dv, void* ws_do_headmajor,
                                 float* ws_delta, void* stream) {
  GH_REQUIRE(attn_tensor_ok(q) && attn_tensor_ok(k) && attn_tensor_ok(v) && attn_tensor_ok(dq) && attn_tensor_ok(dk) &&
                 attn_tensor_ok(dv),
             GH_ERR_ALIGN, "gh_flash_attn_bwd: q/k/v/dq/dk/dv must be non-NULL, 16B aligned, strides multiples of 8");
  GH_REQUIRE(o && d_o && o->seg1 && d_o->seg1 && lse2 && ws_do_headmajor && ws_delta, GH_ERR_NULL,
             "gh_flash_attn_bwd: NULL pointer");
  GH_REQUIRE(D == 64 || D == 128, GH_ERR_UNSUPPORTED, "gh_flash_attn_bwd: head dim %d unsupported (64, 128)", D);
  if (d_valid <= 0) d_valid = D;
  GH_REQUIRE(d_valid % 16 == 0 && d_valid <= D, GH_ERR_BAD_SHAPE, "gh_flash_attn_bwd: d_valid=%d must be a multiple of 16, <= D", d_valid);
  GH_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, GH_ERR_BAD_SHAPE, "gh_flash_attn_bwd: bad shape");
  if (B == 0 || Lq == 0) return GH_OK;
  GH_REQUIRE(aligned16(ws_do_headmajor), GH_ERR_ALIGN, "gh_flash_attn_bwd: workspace must be 16B aligned");
  for (const gh_attn_out* t : {o, d_o})   // (the prep pass reads o / d_o with 16-byte accesses)
    GH_REQUIRE(t->n_split >= 0 && t->n_split <= Lq && t->seg1_row_stride % 8 == 0 && t->seg1_batch_stride % 8 == 0 && aligned16(t->seg1) &&
                   (t->n_split == 0 || (t->seg0 && t->seg0_row_stride % 8 == 0 && t->seg0_batch_stride % 8 == 0 && aligned16(t->seg0))),
               GH_ERR_ALIGN, "gh_flash_attn_bwd: o / d_o segments must be 16B aligned with strides that are multiples of 8 elements");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SegOut so, sdo;
  seg_from(o, &so);
  seg_from(d_o, &sdo);
  const int64_t n = static_cast<int64_t>(B) * Lq * H;
  if (D == 64)
    attn_bwd_prep_kernel<64><<<static_cast<int>((n + 31) / 32), 256, 0, s>>>(so, sdo, B, H, Lq, ws_delta,
                                                                            static_cast<bf16*>(ws_do_headmajor));
  else
    attn_bwd_prep_kernel<128><<<static_cast<int>((n + 15) / 16), 256, 0, s>>>(so, sdo, B, H, Lq, ws_delta,
                                                                             static_cast<bf16*>(ws_do_headmajor));
  GH_CHECK_CUDA(cudaGetLastError());

  gh_attn_tensor dot;
  dot.ptr = ws_do_headmajor;
  dot.row_stride = D;
  dot.head_stride = static_cast<int64_t>(Lq) * D;
  dot.batch_stride = static_cast<int64_t>(H) * Lq * D;
  AttnBwdParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.dvalid = d_valid;
  p.prof = g_attn_prof;
  p.lse2 = lse2; p.delta = ws_delta;
  p.dq = static_cast<bf16*>(const_cast<void*>(dq->ptr)); p.dq_bs = dq->batch_stride; p.dq_hs = dq->head_stride; p.dq_rs = dq->row_stride;
  p.dk = static_cast<bf16*>(const_cast<void*>(dk->ptr)); p.dk_bs = dk->batch_stride; p.dk_hs = dk->head_stride; p.dk_rs = dk->row_stride;
  p.dv = static_cast<bf16*>(const_cast<void*>(dv->ptr)); p.dv_bs = dv->batch_stride; p.dv_hs = dv->head_stride; p.dv_rs = dv->row_stride;
#ifndef GH_ATTN_BWD_V1   // (A/B builds, tools/build_ab.sh: the first form -- 128 x 64 tiles, every operand in shared memory)
  {
    CUtensorMap mk_, mv, mq, mdo;   // every tile of the second form is 128 rows x 64 lanes x D / 64 boxes
    if (int e = make_qkv_map(&mk_, k, B, H, Lk, D, 128)) return e;
    if (int e = make_qkv_map(&mv, v, B, H, Lk, D, 128)) return e;
    if (int e = make_qkv_map(&mq, q, B, H, Lq, D, 128)) return e;
    if (int e = make_qkv_map(&mdo, &dot, B, H, Lq, D, 128)) return e;
    CUtensorMap mdq, mdk, mdv;      // outputs: bulk tensor stores of [128 rows x 64 lanes] tiles, clipped at the sequence end
    if (int e = make_qkv_map(&mdq, dq, B, H, Lq, D, 128)) return e;
    if (int e = make_qkv_map(&mdk, dk, B, H, Lk, D, 128)) return e;
    if (int e = make_qkv_map(&mdv, dv, B, H, Lk, D, 128)) return e;
    const dim3 grid_kv((Lk + 127) / 128, H, B), grid_q((Lq + 127) / 128, H, B);
    if (D == 64) {
      flash_bwd_dkv2_kernel<64><<<grid_kv, ATT_BWD2_THREADS, AttnBwdKV2Cfg<64>::SMEM_BYTES, s>>>(mk_, mv, mq, mdo, mdk, mdv, p);
      flash_bwd_dq2_kernel<64><<<grid_q, ATT_BWD2_THREADS, AttnBwdQ2Cfg<64>::SMEM_BYTES, s>>>(mq, mdo, mk_, mv, mdq, p);
    } else {
      flash_bwd_dkv2_kernel<128><<<grid_kv, ATT_BWD2_THREADS, AttnBwdKV2Cfg<128>::SMEM_BYTES, s>>>(mk_, mv, mq, mdo, mdk, mdv, p);
      flash_bwd_dq2_kernel<128><<<grid_q, ATT_BWD2_THREADS, AttnBwdQ2Cfg<128>::SMEM_BYTES, s>>>(mq, mdo, mk_, mv, mdq, p);
    }
    GH_CHECK_CUDA(cudaGetLastError());
    return GH_OK;
  }
#else   // GH_ATTN_BWD_V1
  {
    CUtensorMap mk_, mv, mq, mdo;
    if (int e = make_qkv_map(&mk_, k, B, H, Lk, D, 128)) return e;
    if (int e = make_qkv_map(&mv, v, B, H, Lk, D, 128)) return e;
    if (int e = make_qkv_map(&mq, q, B, H, Lq, D, 64)) return e;
    if (int e = make_qkv_map(&mdo, &dot, B, H, Lq, D, 64)) return e;
    dim3 grid((Lk + 127) / 128, H, B);
    if (D == 64)
      flash_bwd_dkv_kernel<64><<<grid, ATT_BWD_THREADS, AttnBwdKVCfg<64>::SMEM_BYTES, s>>>(mk_, mv, mq, mdo, p);
    else
      flash_bwd_dkv_kernel<128><<<grid, ATT_BWD_THREADS, AttnBwdKVCfg<128>::SMEM_BYTES, s>>>(mk_, mv, mq, mdo, p);
    GH_CHECK_CUDA(cudaGetLastError());
  }
  {
    CUtensorMap mk_, mv, mq, mdo;
    if (int e = make_qkv_map(&mk_, k, B, H, Lk, D, 64)) return e;
    if (int e = make_qkv_map(&mv, v, B, H, Lk, D, 64)) return e;
    if (int e = make_qkv_map(&mq, q, B, H, Lq, D, 128)) return e;
    if (int e = make_qkv_map(&mdo, &dot, B, H, Lq, D, 128)) return e;
    dim3 grid((Lq + 127) / 128, H, B);
    if (D == 64)
      flash_bwd_dq_kernel<64><<<grid, ATT_BWD_THREADS, AttnBwdQCfg<64>::SMEM_BYTES, s>>>(mq, mdo, mk_, mv, p);
    else
      flash_bwd_dq_kernel<128><<<grid, ATT_BWD_THREADS, AttnBwdQCfg<128>::SMEM_BYTES, s>>>(mq, mdo, mk_, mv, p);
    GH_CHECK_CUDA(cudaGetLastError());
  }
  return GH_OK;
#endif
}
