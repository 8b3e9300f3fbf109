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
  SegOut o;
  float* lse2;       // [B, H, Lq]  log2-domain logsumexp of the scaled scores
};

template <int D>
struct AttnFwdCfg {
  static constexpr int Q_BYTES = ATT_BQ * D * 2;
  static constexpr int K_BYTES = ATT_BKV * D * 2;
  static constexpr int V_BYTES = ATT_BKV * D * 2;
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = Q_BYTES;                 // 2 stages
  static constexpr int OFF_V = OFF_K + 2 * K_BYTES;     // 1 stage
  static constexpr int OFF_P = OFF_V + V_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int TMEM_COLS = 256;                 // S: 2 x 64, O: D (<=128)
  static constexpr int TM_S = 0, TM_O = 128;
};

// write 8 bf16 (16 B) of row r, logical 16-byte chunk c, into a SWIZZLE_128B K-major tile (128 B rows)
__device__ __forceinline__ void st_sw128(uint8_t* tile, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

template <int D>
__global__ void __launch_bounds__(160, 2)
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
    if (lane == 0) {
      // ---------------- control thread: TMA + MMA issue ----------------
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, ATT_BKV, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);
      const uint64_t kdesc = umma_desc_base(16u, 1024u);      // K-major tiles (Q, K, P)
      const uint64_t vdesc = umma_desc_base(8192u, 1024u);    // V as MN-major B operand
      auto load_k = [&](int j) {
        uint8_t* dst = sK + (j & 1) * Cfg::K_BYTES;
        mbar_arrive_expect_tx(&bar_k[j & 1], Cfg::K_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) tma_load_4d(dst + c * 8192, &tm_k, &bar_k[j & 1], c * 64, j * ATT_BKV, h, b);
      };
      auto load_v = [&](int j) {
        mbar_arrive_expect_tx(bar_v, Cfg::V_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c) tma_load_4d(sV + c * 8192, &tm_v, bar_v, c * 64, j * ATT_BKV, h, b);
      };
      auto issue_s = [&](int j) {
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK + (j & 1) * Cfg::K_BYTES);
#pragma unroll
        for (int c = 0; c < DC; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem + Cfg::TM_S + (j & 1) * ATT_BKV, umma_desc_at(kdesc, aQ + c * 16384 + k * 32),
                    umma_desc_at(kdesc, aK + c * 8192 + k * 32), idesc_s, (c | k) != 0 ? 1u : 0u);
        umma_commit(&bar_s[j & 1]);
      };
      mbar_arrive_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
      for (int c = 0; c < DC; ++c) tma_load_4d(sQ + c * 16384, &tm_q, bar_q, c * 64, q0, h, b);
      load_k(0);
      load_v(0);
      if (nkv > 1) load_k(1);
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {
          // S buffer (j+1)&1 was last read by softmax_{j-1}, which arrived on bar_p before we got here
          mbar_wait(&bar_k[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(bar_p, j & 1);  // P_j is in smem (and O has been rescaled if needed)
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
        {
          const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem + Cfg::TM_O, umma_desc_at(kdesc, aP + k * 32), umma_desc_at(vdesc, aV + k * 2048), idesc_o,
                    (j | k) != 0 ? 1u : 0u);
          umma_commit(bar_o);
        }
        if (j + 2 < nkv) load_k(j + 2);  // K buffer j&1 is free: S_j completed before softmax_j started
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
        tmem_ld_32x32(t_lane + Cfg::TM_S + (j & 1) * ATT_BKV, s0);
        tmem_ld_32x32(t_lane + Cfg::TM_S + (j & 1) * ATT_BKV + 32, s1);
        tmem_ld_wait();
      }
      const int kv_left = p.Lk - j * ATT_BKV;  // valid keys in this block
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float v = __uint_as_float(s[c]) * p.scale_log2;
        if (c >= kv_left) v = -INFINITY;
        s[c] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
      float factor = 1.f;
      if (j == 0) {
        m_used = mx;
      } else if (mx > m_used + 8.f) {
        factor = exp2f(m_used - mx);
        m_used = mx;
      }
      float rs = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        const float e0 = exp2f(__uint_as_float(s[c]) - m_used);
        const float e1 = exp2f(__uint_as_float(s[c + 1]) - m_used);
        pk[c >> 1] = pack_bf16x2(e0, e1);
        const float2 r2 = unpack_bf16x2(pk[c >> 1]);  // sum what the MMA will actually see
        rs += r2.x + r2.y;
      }
      l_sum = l_sum * factor + rs;
      if (j > 0) {
        mbar_wait(bar_o, (j - 1) & 1);  // PV_{j-1} retired: P buffer reusable, O stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, factor != 1.f)) {
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
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
      tmem_ld_32x32(t_lane + Cfg::TM_O + c * 32, o);
      tmem_ld_wait();
      if (orow) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

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
  return GH_OK;
}

}  // namespace gh

using namespace gh;

extern "C" int gh_flash_attn_fwd(const gh_attn_tensor* q, const gh_attn_tensor* k, const gh_attn_tensor* v,
                                 int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t D, float scale,
                                 const gh_attn_out* o, float* lse2, void* stream) {
  GH_REQUIRE(attn_tensor_ok(q) && attn_tensor_ok(k) && attn_tensor_ok(v), GH_ERR_ALIGN,
             "gh_flash_attn_fwd: q/k/v must be non-NULL, 16B aligned, strides multiples of 8 elements");
  GH_REQUIRE(o && o->seg1, GH_ERR_NULL, "gh_flash_attn_fwd: output is NULL");
  GH_REQUIRE(D == 64 || D == 128, GH_ERR_UNSUPPORTED, "gh_flash_attn_fwd: head dim %d unsupported (64, 128)", D);
  GH_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, GH_ERR_BAD_SHAPE, "gh_flash_attn_fwd: bad shape");
  if (B == 0 || Lq == 0) return GH_OK;
  GH_REQUIRE(o->n_split >= 0 && o->n_split <= Lq && (o->n_split == 0 || o->seg0), GH_ERR_BAD_SHAPE,
             "gh_flash_attn_fwd: bad output split");
  GH_REQUIRE(o->seg1_row_stride % 8 == 0 && o->seg1_batch_stride % 8 == 0 && aligned16(o->seg1) &&
                 (o->n_split == 0 || (o->seg0_row_stride % 8 == 0 && o->seg0_batch_stride % 8 == 0 && aligned16(o->seg0))),
             GH_ERR_ALIGN, "gh_flash_attn_fwd: output strides must be multiples of 8 elements");
  CUtensorMap mq, mk_, mv;
  if (int e = make_qkv_map(&mq, q, B, H, Lq, D, ATT_BQ)) return e;
  if (int e = make_qkv_map(&mk_, k, B, H, Lk, D, ATT_BKV)) return e;
  if (int e = make_qkv_map(&mv, v, B, H, Lk, D, ATT_BKV)) return e;
  AttnFwdParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o.p0 = static_cast<bf16*>(o->seg0); p.o.bs0 = o->seg0_batch_stride; p.o.rs0 = o->seg0_row_stride;
  p.o.p1 = static_cast<bf16*>(o->seg1); p.o.bs1 = o->seg1_batch_stride; p.o.rs1 = o->seg1_row_stride;
  p.o.n_split = o->n_split;
  p.lse2 = lse2;
  dim3 grid((Lq + ATT_BQ - 1) / ATT_BQ, H, B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (D == 64)
    flash_fwd_kernel<64><<<grid, 160, AttnFwdCfg<64>::SMEM_BYTES, s>>>(mq, mk_, mv, p);
  else
    flash_fwd_kernel<128><<<grid, 160, AttnFwdCfg<128>::SMEM_BYTES, s>>>(mq, mk_, mv, p);
  GH_CHECK_CUDA(cudaGetLastError());
  return GH_OK;
}
