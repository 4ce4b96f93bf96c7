// C-ABI glue: version / error strings / device attribute cache / GEMM entry point.
#include "common.cuh"

namespace onssen {
int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace onssen

extern "C" const char* onssen_version(void) { return "onssen_b200 0.1 (sm_100a)"; }

extern "C" const char* onssen_error_string(int code) {
  switch (code) {
    case ONSSEN_OK: return "ok";
    case ONSSEN_ERR_ARG: return "bad argument";
    case ONSSEN_ERR_UNSUPPORTED: return "unsupported shape";
    case ONSSEN_ERR_CUDA: return "CUDA runtime error";
    case ONSSEN_ERR_DRIVER: return "CUDA driver entry point unavailable";
    case ONSSEN_ERR_RESIDENCY: return "persistent grid does not fit on the device";
    default: return "unknown error";
  }
}

extern "C" int onssen_num_sms(void) { return onssen::num_sms(); }

extern "C" int onssen_gemm_f16(const void* A, const void* W, const float* bias, float* out, int M, int N, int K,
                               long long lda, long long ldw, long long ld_out, int epi, int group,
                               int remap_inner, int remap_outer, void* stream) {
  if (!A || !W || !out) return ONSSEN_ERR_ARG;
  return onssen::gemm_f16(A, W, bias, out, M, N, K, lda, ldw, ld_out, epi, group, remap_inner, remap_outer,
                          (cudaStream_t)stream);
}

extern "C" int onssen_gemm_f16_rows(const void* X, const void* Y, float* out, int M, int N, int Kc, long long ldx,
                                    long long ldy, long long ld_out, int y_row_shift, const float* out_scale,
                                    void* stream) {
  if (!X || !Y || !out) return ONSSEN_ERR_ARG;
  return onssen::gemm_f16_rows(X, Y, out, M, N, Kc, ldx, ldy, ld_out, y_row_shift, out_scale, (cudaStream_t)stream);
}

extern "C" int onssen_gemm_f16_ex(const void* A, const void* W, const float* bias, float* out, int M, int N,
                                  int K, long long lda, long long ldw, long long ld_out, int epi, int group,
                                  int remap_inner, int remap_outer, const float* out_scale, float* inv_norm,
                                  void* stream) {
  if (!A || !W || !out) return ONSSEN_ERR_ARG;
  return onssen::gemm_f16(A, W, bias, out, M, N, K, lda, ldw, ld_out, epi, group, remap_inner, remap_outer,
                          (cudaStream_t)stream, out_scale, inv_norm);
}

extern "C" int onssen_gemm_l2norm_supported(int group) { return onssen::gemm_l2norm_group_supported(group) ? 1 : 0; }

extern "C" size_t onssen_bn_scratch_bytes(int M, int H) {
  const int C = 2 * onssen::hp_of(H);
  return (size_t)2 * onssen_bn_num_chunks(M) * C * sizeof(double) + (size_t)2 * C * sizeof(float);
}
