// Microbenchmark: all-to-all of 16-byte chunks inside a thread-block cluster with st.async (completion counted in
// bytes on the receiver's mbarrier) -- the exchange of the cluster variant of the BLSTM forward recurrence: every
// CTA of a cluster of C pushes `nch` 16-byte chunks into every member's tile each round and waits until the C*nch
// chunks addressed to it have landed.  Pure ping-pong (no compute between rounds): cycles per round = floor of the
// per-step exchange latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I onssen_b200/csrc scripts/microbench/cluster_stasync.cu -o gpurun_out/cluster_stasync
#include <cstdio>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace tc05;

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t dst, uint4 v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

__device__ __forceinline__ void st_cluster_v4(uint32_t dst, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_volatile_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
// mode 0: every thread waits on the mbarrier; mode 1: one warp waits, then __syncthreads;
// mode 2: plain st.shared::cluster.v4 whose 4th word is the round tag, every thread polls the chunks it consumes
// (tid, tid+256, ...) in its own shared memory
__global__ void __launch_bounds__(288, 1) xchg_kernel(long long* out, int iters, int nch, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t rank, C;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(C));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);               // [2]
  uint8_t* tile = smem + 128;                                        // [2][C][nch] x 16 B
  const int tid = threadIdx.x;
  const uint32_t expect = C * nch * 16;
  if (tid == 0) {
    mbar_init(bars, 1); mbar_init(bars + 1, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(bars, expect);
    mbar_arrive_expect_tx(bars + 1, expect);
  }
  __syncthreads();
  cluster_sync_all();
  long long t0 = 0;
  unsigned bad = 0;
  for (int it = 0; it < iters; ++it) {
    if (it == 16) t0 = clock64();
    const int par = it & 1;
    if (tid < 256) {
      // chunk-major: thread handles (chunk, dest) pairs; consecutive threads -> consecutive destinations
      for (int idx = tid; idx < nch * (int)C; idx += 256) {
        const int dest = idx % C, ch = idx / C;
        const uint32_t v = it * 977 + ch + 131 * rank;
        if (mode == 2)
          st_cluster_v4(mapa(smem_u32(tile + (((size_t)par * C + rank) * nch + ch) * 16), dest), make_uint4(v, v + 1, v + 2, (uint32_t)it + 7u));
        else
          st_async_v4(mapa(smem_u32(tile + (((size_t)par * C + rank) * nch + ch) * 16), dest), make_uint4(v, v + 1, v + 2, v + 3),
                      mapa(smem_u32(bars + par), dest));
      }
    }
    if (mode == 2) {
      if (tid < 256)
        for (int idx = tid; idx < nch * (int)C; idx += 256) {
          const uint32_t a = smem_u32(tile + ((size_t)par * C * nch + idx) * 16);
          while (lds_volatile_v4(a).w != (uint32_t)it + 7u) {
          }
        }
      __syncthreads();
    } else {
      if (mode == 0 || tid >= 256) mbar_wait_cluster(bars + par, (it >> 1) & 1);
      if (mode == 1) __syncthreads();
      if (tid == 256) mbar_arrive_expect_tx(bars + par, expect);   // next use of this parity: round it+2
    }
    // payload check: chunk (tid % nch) of producer (rank+1)%C
    {
      const uint32_t src = (rank + 1) % C, ch = tid % nch;
      const uint4 v = *reinterpret_cast<const uint4*>(tile + (((size_t)par * C + src) * nch + ch) * 16);
      const uint32_t e = it * 977 + ch + 131 * src;
      if (v.x != e || v.z != e + 2) bad = it + 1;
    }
  }
  const long long t1 = clock64();
  cluster_sync_all();
  if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / (iters - 16);
  if (bad) out[8] = bad;
}

int main() {
  long long* d;
  cudaMalloc(&d, 128);
  cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("cluster chunks/dest bytes_out/CTA clusters mode | cycles per all-to-all round, payload check\n");
  for (int C : {2, 4, 8, 16})
    for (int nch : {8, 55, 96})
      for (int nclusters : {6})
        for (int mode : {0, 2}) {
          cudaMemset(d, 0, 128);
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(C * nclusters);
          cfg.blockDim = dim3(288);
          cfg.dynamicSmemBytes = 180 * 1024;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          int nact = -1;
          cudaOccupancyMaxActiveClusters(&nact, xchg_kernel, &cfg);
          if (nclusters > nact) { printf("%4d %4d: %d clusters do not fit (max %d)\n", C, nch, nclusters, nact); continue; }
          cudaError_t e = cudaLaunchKernelEx(&cfg, xchg_kernel, d, 2016, nch, mode);
          long long h[16] = {0};
          if (e == cudaSuccess) e = cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("C=%d nch=%d: %s\n", C, nch, cudaGetErrorString(e)); cudaGetLastError(); continue; }
          printf("%4d %6d %10d %6d (max %d) %d | %6lld  %s\n", C, nch, nch * 16 * C, nclusters, nact, mode, h[0], h[8] ? "PAYLOAD MISMATCH" : "ok");
        }
  return 0;
}
