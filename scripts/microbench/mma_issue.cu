// Microbenchmark (profiling aid, not product code): what does the single MMA-issuing thread pay per tcgen05.mma, per
// tcgen05.commit and per mbarrier wait?  One CTA, garbage operands (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I semantic-superpoint_b200/csrc scripts/microbench/mma_issue.cu -o gpurun_out/mma_issue
#include "tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>

__global__ void __launch_bounds__(128, 1) k(long long* out, int mode, int groups, int per_group, int n) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[16];
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { for (int i = 0; i < 16; ++i) tc::mbar_init(bars + i, 1); tc::fence_barrier_init(); }
    __syncwarp();
    tc::tmem_alloc(&tptr, 512);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tptr;
  if (warp == 2 && (mode & 32)) {
    // whole warp runs the loop (uniform control flow, operands in uniform registers), one elected lane issues
    const uint32_t su = tc::smem_u32(smem);
    const uint32_t idesc_ts = tc::idesc_bf16_f32(128, n, 0, 1);
    const uint32_t idesc_ss = tc::idesc_bf16_f32(128, n, 0, 0);
    const uint64_t db_mn = tc::smem_desc_sw128(su, 8192, 1024), da = tc::smem_desc_sw128(su + 65536, 16, 1024), db_k = tc::smem_desc_sw128(su, 16, 1024);
    const bool el = tc::elect_one();
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (mode & 4) {
        tc::mbar_wait(bars + 8, 1);
        tc::mbar_wait(bars + 9, 1);
        tc::fence_after_sync();
      }
      if (mode & 8) {
        if (g >= 4) { const int s = (g - 4) & 3; tc::mbar_wait(bars + s, ((g - 4) >> 2) & 1); }
      }
      for (int i = 0; i < per_group; i += 4) {
        if (el) {
          if (mode & 16) tc::mma_ss_x4(0u, da, db_k, idesc_ss, 1u);
          else tc::mma_ts_x4(0u, 256u, db_mn + (uint64_t)((g & 3) * 512), idesc_ts, 1u);
        }
      }
      if ((mode & 2) && el) tc::mma_commit(bars + (g & 3));
      __syncwarp();
    }
    long long t1 = clock64();
    if (el) tc::mma_commit(bars + 10);
    tc::mbar_wait(bars + 10, 0);
    long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (warp == 1 && lane == 0 && !(mode & 32)) {
    const uint32_t su = tc::smem_u32(smem);
    const uint32_t idesc_ts = tc::idesc_bf16_f32(128, n, 0, 1);   // A TMEM, B MN-major
    const uint32_t idesc_ss = tc::idesc_bf16_f32(128, n, 0, 0);   // A, B K-major smem
    const uint64_t db_mn = tc::smem_desc_sw128(su, 8192, 1024), da = tc::smem_desc_sw128(su + 65536, 16, 1024), db_k = tc::smem_desc_sw128(su, 16, 1024);
    long long t0 = clock64();
    int committed = 0;
    for (int g = 0; g < groups; ++g) {
      if (mode & 4) {  // two waits on barriers that are already complete (fresh barrier, parity 1 passes)
        tc::mbar_wait(bars + 8, 1);
        tc::mbar_wait(bars + 9, 1);
        tc::fence_after_sync();
      }
      if (mode & 8) {  // ring back-pressure: wait for the commit issued 4 groups ago
        if (g >= 4) { const int s = (g - 4) & 3; tc::mbar_wait(bars + s, ((g - 4) >> 2) & 1); }
      }
      for (int i = 0; i < per_group; i += 4) {
        if (mode & 16) tc::mma_ss_x4(tb, da, db_k, idesc_ss, 1u);
        else tc::mma_ts_x4(tb, tb + 256, db_mn, idesc_ts, 1u);
      }
      if (mode & 2) { tc::mma_commit(bars + (g & 3)); ++committed; }
    }
    long long t1 = clock64();
    tc::mma_commit(bars + 10);
    tc::mbar_wait(bars + 10, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char* name; int mode, groups, per, n; } E[] = {
      {"TS N=128: 1024 MMAs, no commit", 0, 128, 8, 128},
      {"TS N=128: 8 MMAs + commit", 2, 128, 8, 128},
      {"TS N=128: 2 passed waits + 8 MMAs + commit", 6, 128, 8, 128},
      {"TS N=128: ring wait(g-4) + 8 MMAs + commit", 10, 128, 8, 128},
      {"TS N=128: 2 waits + ring + 8 MMAs + commit", 14, 128, 8, 128},
      {"TS N=128: 2 waits + ring + 4 MMAs + commit", 14, 128, 4, 128},
      {"TS N=128: 2 waits + ring + 16 MMAs + commit", 14, 128, 16, 128},
      {"SS N=256: 1024 MMAs, no commit", 16, 128, 8, 256},
      {"SS N=256: 1 wait-pair + ring + 12 MMAs + commit", 16 | 14, 128, 12, 256},
      {"SS N=128: 1024 MMAs, no commit", 16, 128, 8, 128},
      {"TS N=64: 1024 MMAs, no commit", 0, 128, 8, 64},
      {"uniform TS N=128: 1024 MMAs, no commit", 32, 128, 8, 128},
      {"uniform TS N=128: 8 MMAs + commit", 32 | 2, 128, 8, 128},
      {"uniform TS N=128: 2 waits + ring + 8 MMAs + commit", 32 | 14, 128, 8, 128},
      {"uniform TS N=128: 2 waits + ring + 4 MMAs + commit", 32 | 14, 128, 4, 128},
      {"uniform TS N=64: 1024 MMAs, no commit", 32, 128, 8, 64},
      {"uniform TS N=32: 1024 MMAs, no commit", 32, 128, 8, 32},
      {"uniform SS N=256: 2 waits + ring + 12 MMAs + commit", 32 | 16 | 14, 128, 12, 256},
      {"uniform SS N=128: 1024 MMAs, no commit", 32 | 16, 128, 8, 128},
      {"TS N=256: 1024 MMAs, no commit", 0, 128, 8, 256},
  };
  for (auto& e : E) {
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
      k<<<1, 128, 200 * 1024>>>(d, e.mode, e.groups, e.per, e.n);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("%s: %s\n", e.name, cudaGetErrorString(err)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    const int nm = e.groups * e.per;
    printf("%-52s issue %7.1f cyc/MMA  total %7.1f cyc/MMA  (%6.1f cyc/group)\n", e.name, (double)h[0] / nm, (double)h[1] / nm, (double)h[1] / e.groups);
  }
  return 0;
}
