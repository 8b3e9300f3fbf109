// genhancer_b200 -- sm_100a device-side primitives (inline PTX).
//
// Thin wrappers over the Blackwell async machinery used by every tensor-core
// kernel in this library: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (TMEM alloc / mma / commit / ld) and the shared-memory matrix descriptors
// they consume.  Nothing here is a port of reference code: the reference
// (Jam1ezhang/GenHancer) ships no native code at all.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace gh {

// ----------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
// One lane of a converged warp.  Unlike `lane == 0`, the compiler knows that exactly one thread runs the guarded
// region, so operands of the async-unit instructions issued there (UTCHMMA / UTMALDG / UTCBAR take uniform
// registers) need no ELECT / R2UR / BRA.U.ANY waterfall loop per instruction -- which cost ~100 cycles per MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // make generic-proxy shared-memory writes visible to the async proxy (TMA / tcgen05)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin on a phase; a wait that lasts > ~10 s of SM clocks is a protocol bug: trap instead of
// hanging the GPU (turns a deadlock into a reportable CUDA error).
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar_addr, uint32_t parity) {
  printf("genhancer_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
         threadIdx.x, bar_addr, parity);
  __trap();
}
// try_wait with a suspend-time hint: the hardware parks the warp (no issue slots burned by a spin loop, which
// matters when a waiting warp shares its scheduler with a working one) until the phase flips or ~hint_ns pass.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (clock64() - t0 > 20000000000LL) mbar_timeout_trap(smem_u32(bar), parity);
  }
}

// ----------------------------------------------------------------------------
// TMA loads (tile mode). Coordinates are innermost-first, in elements.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// L2 prefetch of a tensor box (no shared memory, no barrier): turns the DRAM latency of a later cp.async.bulk.tensor of
// the same box into an L2 hit, for pipelines whose smem budget cannot hold enough stages to cover DRAM latency.
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// TMA store (smem tile -> global, clipped at the tensor's bounds) and its bulk-group bookkeeping
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {  // at most N groups still READING their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): the two SMs of a TPC run ONE 256-row MMA; each CTA stages its own
// 128 rows of A and HALF of the B tile, so the shared-memory traffic per SM and per MMA drops from
// (128 + BN) * 32 B to (128 + BN/2) * 32 B.  Single-CTA 128 x 256 tiles are shared-memory-bandwidth bound
// (12 KB operand reads + 12 KB TMA writes per 128-cycle MMA > 128 B/clk); the pair is not.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta address of THIS CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(rank));
  return r;
}
// wait with cluster-scope acquire: the data the barrier guards was written from the OTHER CTA of the cluster
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 20000000000LL) mbar_timeout_trap(addr, parity);
  }
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Arrive WITHOUT release semantics: for barriers that guard no memory, only "I am done reading TMEM" (the accumulator
// hand-back of the epilogue warps, ordered by tcgen05.fence::before_thread_sync).  The releasing form costs a
// MEMBAR.ALL.CTA + ERRBAR per arrive at cluster scope -- 15 % of all warp-stall samples of the K = 1024 GEMMs
// (ncu source page, profiles/r02_gemm_vit_fc1_hot_instructions.txt), once per epilogue warp and tile, on the path that
// frees the accumulator for the next tile's MMAs.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
#ifdef GH_TEMPTY_RELEASE   // A/B builds (tools/build_ab.sh): the releasing form
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// TMA loads of a CTA pair: data lands in THIS CTA's smem, the transaction bytes are counted on the mbarrier at
// cluster address `bar_cluster` (the leader CTA's full barrier).
__device__ __forceinline__ void tma2_load_2d(void* smem, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* smem, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {  // cta_group::2: one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: shared-memory matrix descriptors (SWIZZLE_128B, 16-bit elements)
//
// bit layout (PTX ISA "tcgen05 shared memory descriptor"):
//   [ 0,14) start address >> 4      [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4 [46,48) version (1 on sm_100)
//   [49,52) base offset             [61,64) layout type (2 = SWIZZLE_128B)
//
// K-major tile  (rows = M/N index, 64 elements = 128 B of K per row):
//     8-row core groups are 1024 B apart  -> SBO = 1024, LBO unused (=1)
//     advancing 16 elements of K          -> start address + 32 B
// MN-major tile (rows = K index, 64 elements = 128 B of M/N per row,
//                successive 64-wide M/N chunks `chunk_bytes` apart):
//     8-K-row groups are 1024 B apart     -> SBO = 1024
//     next 64-element M/N chunk           -> LBO = chunk_bytes
//     advancing 16 elements of K          -> start address + 2048 B
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version = 1 (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint64_t umma_desc_at(uint64_t base, uint32_t smem_addr_bytes) {
  return base | static_cast<uint64_t>((smem_addr_bytes >> 4) & 0x3FFFu);
}

// instruction descriptor for kind::f16, bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                                   // D format  = F32
         | (1u << 7)                                 // A format  = BF16
         | (1u << 10)                                // B format  = BF16
         | (static_cast<uint32_t>(a_mn_major) << 15) // A major
         | (static_cast<uint32_t>(b_mn_major) << 16) // B major
         | (static_cast<uint32_t>(N >> 3) << 17)     // N / 8
         | (static_cast<uint32_t>(M >> 4) << 24);    // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA-pair MMA (issued by the leader CTA only): M = 256 over the two CTAs' TMEM, A rows / B rows split across them.
__device__ __forceinline__ void umma2_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the pair's previously issued MMAs retire) on the barrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all tcgen05 ops previously issued by this thread retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05.ld: 32 lanes x 32 consecutive fp32 columns -> 32 registers/thread.
// Warp w of a 4-warp group may only touch TMEM lanes [32*(w%4), 32*(w%4)+32).
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------
// numerics shared by epilogues and elementwise kernels
// ----------------------------------------------------------------------------
enum Act : int { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_QUICK_GELU = 2, ACT_GELU_ERF = 3, ACT_SILU = 4 };

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU op, max rel. error ~2^-11 (outputs are bf16)
  return y;
}
__device__ __forceinline__ float sigmoidf_(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case ACT_GELU_TANH: {
      // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  == x * sigmoid(2u)
      float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return x * sigmoidf_(2.f * u);
    }
    case ACT_QUICK_GELU: return x * sigmoidf_(1.702f * x);
    case ACT_GELU_ERF: return 0.5f * x * (1.f + erff(x * 0.7071067811865476f));
    case ACT_SILU: return x * sigmoidf_(x);
    default: return x;
  }
}
__device__ __forceinline__ float act_bwd(int act, float x) {  // d act(x) / dx
  switch (act) {
    case ACT_GELU_TANH: {
      float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      float s = sigmoidf_(2.f * u);
      float du = 0.7978845608028654f * (1.f + 3.f * 0.044715f * x * x);
      return s + x * s * (1.f - s) * 2.f * du;
    }
    case ACT_QUICK_GELU: {
      float s = sigmoidf_(1.702f * x);
      return s + 1.702f * x * s * (1.f - s);
    }
    case ACT_GELU_ERF: {
      float cdf = 0.5f * (1.f + erff(x * 0.7071067811865476f));
      return cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
    }
    case ACT_SILU: {
      float s = sigmoidf_(x);
      return s + x * s * (1.f - s);
    }
    default: return 1.f;
  }
}

// 8-wide forms for the GEMM epilogue: ONE switch per vector so the 8 independent chains interleave (a switch per
// element serialises them: a lone warp per scheduler then runs at instruction latency, not throughput).
__device__ __forceinline__ void act_fwd8(int act, float (&v)[8]) {
  switch (act) {
    case ACT_GELU_TANH:
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float x = v[k];
        const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
        v[k] = 0.5f * x * (1.f + tanh_approx(u));
      }
      break;
    case ACT_QUICK_GELU:
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] * fmaf(0.5f, tanh_approx(0.851f * v[k]), 0.5f);
      break;
    case ACT_SILU:
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] * fmaf(0.5f, tanh_approx(0.5f * v[k]), 0.5f);
      break;
    case ACT_GELU_ERF:
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.5f * v[k] * (1.f + erff(v[k] * 0.7071067811865476f));
      break;
    default: break;
  }
}
// v[k] *= act'(x[k])
__device__ __forceinline__ void act_bwd_mul8(int act, const float (&x)[8], float (&v)[8]) {
  switch (act) {
    case ACT_GELU_TANH:
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xx = x[k];
        const float u = 0.7978845608028654f * fmaf(0.044715f * xx * xx, xx, xx);
        const float t = tanh_approx(u);
        const float du = 0.7978845608028654f * fmaf(0.134145f * xx, xx, 1.f);
        v[k] *= 0.5f * (1.f + t) + 0.5f * xx * (1.f - t * t) * du;
      }
      break;
    case ACT_QUICK_GELU:
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float s = fmaf(0.5f, tanh_approx(0.851f * x[k]), 0.5f);
        v[k] *= s + 1.702f * x[k] * s * (1.f - s);
      }
      break;
    case ACT_SILU:
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float s = fmaf(0.5f, tanh_approx(0.5f * x[k]), 0.5f);
        v[k] *= s + x[k] * s * (1.f - s);
      }
      break;
    case ACT_GELU_ERF:
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float cdf = 0.5f * (1.f + erff(x[k] * 0.7071067811865476f));
        v[k] *= cdf + x[k] * 0.3989422804014327f * __expf(-0.5f * x[k] * x[k]);
      }
      break;
    default: break;
  }
}

// bare MUFU.EX2 (exp2f without fast-math adds a denormal-range rescale: 2 FMUL + FSETP + FSEL per call)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

}  // namespace gh
