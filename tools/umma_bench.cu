// Micro-benchmark: how fast does ONE SM retire chains of tcgen05.mma (kind::f16, M = 128, K = 16 per instruction)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I genhancer_b200/csrc tools/umma_bench.cu -o gpurun_out/umma_bench
// Variants: operand source (SS / TS), N, one accumulator vs several independent ones, with / without concurrent TMEM
// loads by 16 other warps.  Prints SM cycles per MMA instruction; the math floor is N / 2 cycles.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

using namespace gh;

struct Args {
  int n;        // MMA N
  int ts;       // A operand from TMEM
  int nacc;     // number of independent accumulators the chain rotates over (1, 2, 4)
  int chain;    // k-steps per accumulation chain (acc flag cleared at its start)
  int reps;     // chains per measurement
  int ldwarps;  // 1: the 16 other warps stream tcgen05.ld over the accumulators meanwhile
  int bshift;   // B operand k-step stride in bytes (32 = K-major, 2048 = MN-major)
  int commit;   // 1: tcgen05.commit (to a barrier nobody waits on) after every chain
  long long* out;
};

__global__ void __launch_bounds__(544, 1) bench_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // small bf16 values
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 16) {
    const uint64_t kdesc = umma_desc_base(16u, 1024u);
    const uint64_t mdesc = umma_desc_base(16384u, 1024u);
    const uint32_t idesc = umma_idesc_bf16(128, a.n, false, a.bshift != 32);
    const uint64_t bbase = a.bshift == 32 ? kdesc : mdesc;
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < a.reps; ++r) {
          const uint32_t d = tmem + (a.nacc == 1 ? 0u : static_cast<uint32_t>((r % a.nacc) * a.n));
          for (int ks = 0; ks < a.chain; ++ks) {
            const uint64_t bd = umma_desc_at(bbase, sb + (a.bshift == 32 ? (ks & 3) * 32 : (ks & 7) * 2048));
            umma_ss(d, umma_desc_at(kdesc, sa + (ks & 3) * 32 + (ks >> 2 & 1) * 16384), bd, idesc, ks ? 1u : 0u);
          }
          if (a.commit) umma_commit(&bar2);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, pass & 1);
      t1 = clock64();
    }
    if (lane == 0) {
      a.out[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (a.ldwarps) {
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    while (!stop) {
      uint32_t r[32];
      tmem_ld_32x32(t_lane + (warp >> 2) * 32, r);
      tmem_ld_wait();
      acc += r[3];
    }
    if (acc == 0x12345678u) a.out[200] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// TS variant kept separate so that the accumulator / operand columns are right: A (bf16 pairs) at columns 384.., D at 0..
__global__ void __launch_bounds__(544, 1) bench_ts_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp < 4) {   // fill the A columns with something finite
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = 0x3c003c00u;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    tmem_st_32x32(t_lane + 384, r);
    tmem_st_32x32(t_lane + 416, r);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 16) {
    const uint64_t mdesc = umma_desc_base(16384u, 1024u);
    const uint32_t idesc = umma_idesc_bf16(128, a.n, false, true);
    const uint32_t sb = smem_u32(smem + 32768);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < a.reps; ++r) {
          const uint32_t d = tmem + (a.nacc == 1 ? 0u : static_cast<uint32_t>((r % a.nacc) * a.n));
          for (int ks = 0; ks < a.chain; ++ks)
            umma_ts(d, tmem + 384u + 8u * (ks & 7), umma_desc_at(mdesc, sb + (ks & 7) * 2048), idesc, ks ? 1u : 0u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, pass & 1);
      t1 = clock64();
    }
    if (lane == 0) a.out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out;
  cudaMalloc(&out, 4096);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(bench_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  struct Case { const char* name; int n, ts, nacc, chain, reps, ldwarps, bshift, ctas, commit; };
  const Case cases[] = {
      {"SS N=128 one accumulator, chains of 8        ", 128, 0, 1, 8, 64, 0, 32, 1},
      {"SS N=128 one accumulator, chains of 8, 148 CTAs", 128, 0, 1, 8, 64, 0, 32, 148},
      {"SS N=128 one accumulator, ONE chain of 512   ", 128, 0, 1, 512, 1, 0, 32, 1},
      {"SS N=128 two accumulators alternating chains ", 128, 0, 2, 8, 64, 0, 32, 1},
      {"SS N=64  one accumulator, chains of 8        ", 64, 0, 1, 8, 64, 0, 32, 1},
      {"SS N=64  four accumulators                   ", 64, 0, 4, 8, 64, 0, 32, 1},
      {"SS N=256 one accumulator, chains of 8        ", 256, 0, 1, 8, 64, 0, 32, 1},
      {"SS N=128 B MN-major, chains of 8             ", 128, 0, 1, 8, 64, 0, 2048, 1},
      {"SS N=128 chains of 8 + 16 warps of tcgen05.ld", 128, 0, 1, 8, 64, 1, 32, 1},
      {"TS N=128 one accumulator, chains of 8        ", 128, 1, 1, 8, 64, 0, 2048, 1},
      {"TS N=128 two accumulators                    ", 128, 1, 2, 8, 64, 0, 2048, 1},
      {"TS N=128 one accumulator, 148 CTAs           ", 128, 1, 1, 8, 64, 0, 2048, 148},
      {"SS N=128 chains of 4                         ", 128, 0, 1, 4, 128, 0, 32, 1},
      {"SS N=128 chains of 1 (no accumulation)       ", 128, 0, 1, 1, 512, 0, 32, 1},
      {"SS N=128 four accumulators, chains of 8      ", 128, 0, 4, 8, 64, 0, 32, 1, 0},
      {"SS N=128 one accumulator, chains of 8, commit after each", 128, 0, 1, 8, 64, 0, 32, 1, 1},
      {"SS N=128 four accumulators, chains of 8, commit after each", 128, 0, 4, 8, 64, 0, 32, 1, 1},
      {"SS N=128 four accumulators, chains of 8, commit after each, 16 ld warps", 128, 0, 4, 8, 64, 1, 32, 1, 1},
      {"SS N=256 two accumulators, chains of 8, commit after each", 256, 0, 2, 8, 64, 0, 32, 1, 1},
      {"SS N=128 four accumulators, chains of 16, commit after each", 128, 0, 4, 16, 32, 0, 32, 1, 1},
  };
  for (const Case& c : cases) {
    Args a{c.n, c.ts, c.nacc, c.chain, c.reps, c.ldwarps, c.bshift, c.commit, out};
    cudaMemset(out, 0, 4096);
    if (c.ts)
      bench_ts_kernel<<<c.ctas, 544, 65536>>>(a);
    else
      bench_kernel<<<c.ctas, 544, 65536>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < c.ctas; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = static_cast<double>(mx) / (c.chain * c.reps);
    printf("%s : %7.1f cycles / MMA (math floor %d)  [%s]\n", c.name, per, c.n / 2, cudaGetErrorString(e));
  }
  return 0;
}
