// Microbenchmark: all-to-all exchange of small pieces inside a thread-block cluster through distributed shared
// memory, the pattern of the cluster variant of the BLSTM recurrence: every CTA of a cluster of C
// pushes its piece (bytes) into every member's receive tile with cp.async.bulk.shared::cluster (complete_tx on the
// receiver's per-producer mbarrier) and waits for the C pieces addressed to it.  Reports cycles per exchange round
// (pure ping-pong: no compute between rounds) -- the floor of the per-step exchange latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I onssen_b200/csrc scripts/microbench/cluster_xchg.cu -o gpurun_out/cluster_xchg
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
__device__ __forceinline__ void dsmem_bulk(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}

__global__ void __launch_bounds__(288, 1) xchg_kernel(long long* out, int iters, int bytes, int extra_smem_touch) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t rank, C;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(C));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);               // [2][16]
  uint8_t* stage = smem + 512;                                       // [2][bytes]
  uint8_t* tile = stage + 2 * bytes;                                 // [2][C][bytes]
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid < (int)C) {   // arm both parities
    mbar_arrive_expect_tx(bars + tid, bytes);
    mbar_arrive_expect_tx(bars + 16 + tid, bytes);
  }
  cluster_sync_all();
  long long t0 = 0;
  for (int it = 0; it < iters; ++it) {
    if (it == 16) t0 = clock64();
    const int par = it & 1;
    // "gate phase": generic-proxy writes of the piece, made visible to the async proxy
    if (tid < 256) {
      for (int i = tid; i < bytes / 4; i += 256) reinterpret_cast<uint32_t*>(stage + par * bytes)[i] = it * 977 + i + rank;
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 "gate warps"
    }
    if (tid < (int)C)
      dsmem_bulk(mapa(smem_u32(tile + ((size_t)par * C + rank) * bytes), tid), smem_u32(stage + par * bytes), bytes,
                 mapa(smem_u32(bars + par * 16 + rank), tid));
    // consumer: lane p of the last warp waits for producer p, then re-arms the barrier for its next use
    if (tid >= 256 && tid - 256 < (int)C) {
      mbar_wait(bars + par * 16 + (tid - 256), (it >> 1) & 1);
      mbar_arrive_expect_tx(bars + par * 16 + (tid - 256), bytes);
    }
    __syncthreads();
    if (extra_smem_touch) {   // check the payload of one peer (correctness of the protocol)
      const uint32_t v = reinterpret_cast<uint32_t*>(tile + ((size_t)par * C + (rank + 1) % C) * bytes)[tid % (bytes / 4)];
      if (v != (uint32_t)(it * 977 + tid % (bytes / 4) + (rank + 1) % C)) out[8] = it + 1;
    }
  }
  const long long t1 = clock64();
  cluster_sync_all();
  if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / (iters - 16);
}

int main() {
  long long* d;
  cudaMalloc(&d, 128);
  cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("cluster  piece_B  clusters | cycles per all-to-all round (publish -> all C pieces received), payload check\n");
  for (int C : {2, 4, 8, 10, 12, 16})
    for (int bytes : {512, 1024, 2048})
      for (int nclusters : {1, 6}) {
        cudaMemset(d, 0, 128);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C * nclusters);
        cfg.blockDim = dim3(288);
        cfg.dynamicSmemBytes = 180 * 1024;   // like the real kernel: one CTA per SM
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nact = -1;
        cudaOccupancyMaxActiveClusters(&nact, xchg_kernel, &cfg);
        cudaError_t e = cudaLaunchKernelEx(&cfg, xchg_kernel, d, 2016, bytes, 1);
        long long h[16] = {0};
        if (e == cudaSuccess) e = cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("C=%d bytes=%d: %s\n", C, bytes, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        printf("%4d %8d %6d (max active clusters %d) | %6lld  %s\n", C, bytes, nclusters, nact, h[0], h[8] ? "PAYLOAD MISMATCH" : "ok");
      }
  return 0;
}
