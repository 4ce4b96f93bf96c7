// Shared host/device helpers for the onssen_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/onssen_b200.h"

namespace onssen {

int num_sms();

// internal C++ entry (defined in gemm_tc05.cu); the C ABI wrapper lives in capi.cu
int gemm_f16(const void* A, const void* W, const float* bias, float* out, int M, int N, int K, long long lda,
             long long ldw, long long ld_out, int epi, int group, int remap_inner, int remap_outer,
             cudaStream_t stream, const float* out_scale = nullptr, float* inv_norm = nullptr);
bool gemm_l2norm_group_supported(int group);
int gemm_f16_rows(const void* X, const void* Y, float* out, int M, int N, int Kc, long long ldx, long long ldy,
                  long long ld_out, int y_row_shift, const float* out_scale, cudaStream_t stream);

inline int hp_of(int H) { return ((H + 31) / 32) * 32; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// saturating fp32 -> fp16 (round to nearest even, clamp to +-65504 instead of producing inf)
__device__ __forceinline__ __half to_half_sat(float x) {
  x = fminf(fmaxf(x, -65504.0f), 65504.0f);
  return __float2half_rn(x);
}

}  // namespace onssen

#define ONSSEN_CHECK_LAUNCH() (cudaGetLastError() == cudaSuccess ? ONSSEN_OK : ONSSEN_ERR_CUDA)
