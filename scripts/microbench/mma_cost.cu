// Microbenchmark: cycles for a chain of tcgen05.mma kind::f16 instructions as a function of M, N, operand
// source (A from smem = SS, A from TMEM = TS) and number of independent accumulators.  One CTA, one issuing
// thread; smem contents are irrelevant (zeros).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I onssen_b200/csrc scripts/microbench/mma_cost.cu -o gpurun_out/mma_cost
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace tc05;

__global__ void __launch_bounds__(128, 1) mma_cost_kernel(long long* out, int M, int N, int nmma, int ts, int nacc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(&tmem_ptr, 512);
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    // converged warp, elected lane issues (descriptors stay in uniform registers)
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(M, N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 96 * 1024);
    for (int rep = 0; rep < 3; ++rep) {
      uint64_t da = make_smem_desc(a_addr, M * 16, 128, 0);
      uint64_t db = make_smem_desc(b_addr, N * 16, 128, 0);
      uint32_t ta = tb + 256;
      const long long t0 = clock64();
      for (int k = 0; k < nmma; ++k) {
        const uint32_t d = tb + (k % nacc) * N;
        if (leader) {
          if (ts) umma_f16_ts(d, ta, db, idesc, k >= nacc);
          else umma_f16(d, da, db, idesc, k >= nacc);
        }
        da += (2 * M * 16) >> 4; db += (2 * N * 16) >> 4; ta += 8;
        if (((k + 1) & 3) == 0) { da = make_smem_desc(a_addr, M * 16, 128, 0); db = make_smem_desc(b_addr, N * 16, 128, 0); ta = tb + 256; }
      }
      const long long t1 = clock64();
      if (leader) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (lane == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int nmma = 38;
  printf("M   N   src nacc | issue cyc | total cyc (issue->commit observed) | per MMA\n");
  for (int M : {128, 64}) for (int ts : {0, 1}) for (int nacc : {1, 4}) for (int N : {8, 16, 32, 64, 128, 256}) {
    if (nacc * N > 256 || (M == 128 && N < 16)) continue;
    mma_cost_kernel<<<1, 128, 200 * 1024>>>(d, M, N, nmma, ts, nacc);
    long long h[6]; cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("M=%d N=%d ts=%d: %s\n", M, N, ts, cudaGetErrorString(e)); return 1; }
    printf("%3d %3d %s  %d    | %6lld | %6lld | %.1f\n", M, N, ts ? "TS" : "SS", nacc, h[4], h[5], h[5] / (double)nmma);
  }
  return 0;
}
