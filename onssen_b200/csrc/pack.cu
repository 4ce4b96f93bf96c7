// Layout/precision conversion kernels around the BLSTM GEMMs, and BatchNorm1d.
//   * input / weight packing into the padded fp16 layouts the tcgen05 kernels consume
//   * BatchNorm1d over (B,T) (deep_clustering.py:36-38, enhancement.py:45-47) fused with the fp16 cast
//     that produces the head GEMM's A operand
// All HBM-bound elementwise / column-reduction work: coalesced, vector width limited by the fp32 source.
#include "common.cuh"

namespace onssen {
namespace {

__global__ void pack_input_kernel(const float* __restrict__ x, int B, int T, int I, __half* __restrict__ xh,
                                  int Kp) {
  const long long total = (long long)B * T * Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const long long m = idx / Kp;
    const int b = (int)(m % B);
    const int t = (int)(m / B);
    float v = 0.f;
    if (k < I) v = x[((long long)b * T + t) * I + k];
    xh[idx] = to_half_sat(v);
  }
}

struct PackArgs {
  const float* w_ih[2];
  const float* w_hh[2];
  const float* b_ih[2];
  const float* b_hh[2];
  int H, Hp, I, Kp, in_is_blstm, Hin, Hinp;
};

__device__ __forceinline__ bool decode_row(int rem, int H, int& src_row) {
  const int rb = rem >> 7;
  const int r = rem & 127;
  const int ul = r >> 2;
  const int gate = r & 3;
  const int u = rb * 32 + ul;
  src_row = gate * H + u;
  return u < H;
}

__global__ void pack_wih_kernel(PackArgs a, __half* __restrict__ out) {
  const int G4 = 4 * a.Hp;
  const long long total = 2LL * G4 * a.Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % a.Kp);
    const int n = (int)(idx / a.Kp);
    const int dir = n / G4;
    int src_row;
    const bool row_ok = decode_row(n % G4, a.H, src_row);
    int j = -1;
    if (a.in_is_blstm) {
      const int din = k / a.Hinp;
      const int uin = k % a.Hinp;
      if (din < 2 && uin < a.Hin) j = din * a.Hin + uin;
    } else if (k < a.I) {
      j = k;
    }
    float v = 0.f;
    if (row_ok && j >= 0) v = a.w_ih[dir][(long long)src_row * a.I + j];
    out[idx] = to_half_sat(v);
  }
}

__global__ void pack_whh_kernel(PackArgs a, __half* __restrict__ out) {
  // layout [dir][rb][row 0..127][Hp]: one contiguous row-major slab per (dir, row block); a gate thread
  // streams its own row into its TMEM lane (csrc/lstm_rec.cu)
  const int nrb = a.Hp / 32;
  const long long total = 2LL * nrb * 128 * a.Hp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % a.Hp);
    long long t = idx / a.Hp;
    const int r = (int)(t & 127); t >>= 7;
    const int rb = (int)(t % nrb);
    const int dir = (int)(t / nrb);
    int src_row;
    const bool row_ok = decode_row(rb * 128 + r, a.H, src_row);
    float v = 0.f;
    if (row_ok && k < a.H) v = a.w_hh[dir][(long long)src_row * a.H + k];
    out[idx] = to_half_sat(v);
  }
}

// Linear weight [N][K] fp32 -> fp16 [N][Kp]; blstm != 0 maps input j of 2*Hin to column (j/Hin)*Hinp + j%Hin
__global__ void pack_linear_kernel(const float* __restrict__ w, int N, int K, int blstm, int Hin, int Hinp,
                                   int Kp, __half* __restrict__ out) {
  const long long total = (long long)N * Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const long long n = idx / Kp;
    int j = -1;
    if (blstm) {
      const int din = k / Hinp, uin = k % Hinp;
      if (din < 2 && uin < Hin) j = din * Hin + uin;
    } else if (k < K) {
      j = k;
    }
    out[idx] = to_half_sat(j >= 0 ? w[n * K + j] : 0.f);
  }
}

__global__ void pack_bias_kernel(PackArgs a, float* __restrict__ out) {
  const int G4 = 4 * a.Hp;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * G4) return;
  const int dir = idx / G4;
  int src_row;
  const bool row_ok = decode_row(idx % G4, a.H, src_row);
  out[idx] = row_ok ? a.b_ih[dir][src_row] + a.b_hh[dir][src_row] : 0.f;
}

// ---------------------------------------------------------------- BatchNorm
constexpr int BN_ROWS_PER_CHUNK = 256;

__global__ void __launch_bounds__(256) bn_partial_kernel(const float* __restrict__ y, int M, int C,
                                                         double* __restrict__ part) {
  // grid (C/32, nchunk); block 256 = 8 warps; lane = column
  __shared__ double s_sum[8][32];
  __shared__ double s_sq[8][32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * BN_ROWS_PER_CHUNK;
  const int r1 = min(M, r0 + BN_ROWS_PER_CHUNK);
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const double v = (double)y[(long long)r * C + c];
      s += v;
      q += v * v;
    }
  }
  s_sum[warp][lane] = s;
  s_sq[warp][lane] = q;
  __syncthreads();
  if (warp == 0 && c < C) {
    double ts = 0.0, tq = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      ts += s_sum[w][lane];
      tq += s_sq[w][lane];
    }
    const int nchunk = gridDim.y;
    part[(long long)blockIdx.y * C + c] = ts;
    part[(long long)(nchunk + blockIdx.y) * C + c] = tq;
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ part, int nchunk, int M, int H, int Hp,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float eps,
                                   float momentum, int training, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
  const int C = 2 * Hp;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int dir = c / Hp;
  const int u = c % Hp;
  if (u >= H) {
    scale[c] = 0.f;
    shift[c] = 0.f;
    return;
  }
  const int j = dir * H + u;
  float mean, var;
  if (training) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < nchunk; ++k) {
      s += part[(long long)k * C + c];
      q += part[(long long)(nchunk + k) * C + c];
    }
    const double mu = s / M;
    double v = q / M - mu * mu;
    if (v < 0.0) v = 0.0;
    mean = (float)mu;
    var = (float)v;
    const double unbiased = M > 1 ? v * ((double)M / (double)(M - 1)) : v;
    running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * mean;
    running_var[j] = (1.f - momentum) * running_var[j] + momentum * (float)unbiased;
  } else {
    mean = running_mean[j];
    var = running_var[j];
  }
  const float invstd = 1.0f / sqrtf(var + eps);
  const float sc = gamma[j] * invstd;
  scale[c] = sc;
  shift[c] = beta[j] - mean * sc;
  if (save_mean) save_mean[j] = mean;
  if (save_invstd) save_invstd[j] = invstd;
}

__global__ void bn_apply_f16_kernel(const float* __restrict__ y, long long M, int C,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    __half* __restrict__ out) {
  // 4 elements / thread; C multiple of 8
  const long long total4 = M * C / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % C);
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    const float4 sc = *reinterpret_cast<const float4*>(scale + c);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c);
    __half2 lo = __halves2half2(to_half_sat(fmaf(v.x, sc.x, sh.x)), to_half_sat(fmaf(v.y, sc.y, sh.y)));
    __half2 hi = __halves2half2(to_half_sat(fmaf(v.z, sc.z, sh.z)), to_half_sat(fmaf(v.w, sc.w, sh.w)));
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

__global__ void cast_f16_kernel(const float* __restrict__ y, long long n4, __half* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    __half2 lo = __halves2half2(to_half_sat(v.x), to_half_sat(v.y));
    __half2 hi = __halves2half2(to_half_sat(v.z), to_half_sat(v.w));
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" int onssen_pack_input_f16(const float* x, int B, int T, int I, void* xh, int Kp, void* stream) {
  if (!x || !xh || B <= 0 || T <= 0 || I <= 0 || Kp < I || (Kp & 7)) return ONSSEN_ERR_ARG;
  const long long total = (long long)B * T * Kp;
  pack_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, T, I, (__half*)xh, Kp);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_lstm_pack_layer(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f,
                                      const float* b_hh_f, const float* w_ih_r, const float* w_hh_r,
                                      const float* b_ih_r, const float* b_hh_r, int H, int I, int in_is_blstm,
                                      int Hin, void* wih_p, void* whh_p, float* bias_p, void* stream) {
  if (!w_ih_f || !w_hh_f || !b_ih_f || !b_hh_f || !w_ih_r || !w_hh_r || !b_ih_r || !b_hh_r) return ONSSEN_ERR_ARG;
  if (!wih_p || !whh_p || !bias_p || H <= 0 || I <= 0) return ONSSEN_ERR_ARG;
  if (in_is_blstm && I != 2 * Hin) return ONSSEN_ERR_ARG;
  PackArgs a;
  a.w_ih[0] = w_ih_f; a.w_ih[1] = w_ih_r;
  a.w_hh[0] = w_hh_f; a.w_hh[1] = w_hh_r;
  a.b_ih[0] = b_ih_f; a.b_ih[1] = b_ih_r;
  a.b_hh[0] = b_hh_f; a.b_hh[1] = b_hh_r;
  a.H = H; a.Hp = hp_of(H); a.I = I; a.in_is_blstm = in_is_blstm; a.Hin = Hin;
  a.Hinp = in_is_blstm ? hp_of(Hin) : 0;
  a.Kp = in_is_blstm ? 2 * a.Hinp : ((I + 63) / 64) * 64;
  cudaStream_t s = (cudaStream_t)stream;
  const long long n_ih = 2LL * 4 * a.Hp * a.Kp;
  pack_wih_kernel<<<grid_for(n_ih, 256), 256, 0, s>>>(a, (__half*)wih_p);
  const long long n_hh = 2LL * 4 * a.Hp * a.Hp;
  pack_whh_kernel<<<grid_for(n_hh, 256), 256, 0, s>>>(a, (__half*)whh_p);
  pack_bias_kernel<<<(2 * 4 * a.Hp + 255) / 256, 256, 0, s>>>(a, bias_p);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_pack_linear_f16(const float* w, int N, int K, int in_is_blstm, int Hin, void* out, int Kp,
                                      void* stream) {
  if (!w || !out || N <= 0 || K <= 0 || (Kp & 7)) return ONSSEN_ERR_ARG;
  if (in_is_blstm ? (K != 2 * Hin || Kp != 2 * hp_of(Hin)) : (Kp < K)) return ONSSEN_ERR_ARG;
  pack_linear_kernel<<<grid_for((long long)N * Kp, 256), 256, 0, (cudaStream_t)stream>>>(
      w, N, K, in_is_blstm, Hin, in_is_blstm ? hp_of(Hin) : 0, Kp, (__half*)out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_bn_num_chunks(int M) { return (M + BN_ROWS_PER_CHUNK - 1) / BN_ROWS_PER_CHUNK; }

extern "C" int onssen_bn_forward_f16(const float* y, int M, int H, const float* gamma, const float* beta,
                                     float* running_mean, float* running_var, float eps, float momentum,
                                     int training, void* out_h, float* save_mean, float* save_invstd,
                                     void* scratch, void* stream) {
  if (!y || !gamma || !beta || !running_mean || !running_var || !out_h || !scratch || M <= 0 || H <= 0)
    return ONSSEN_ERR_ARG;
  const int Hp = hp_of(H);
  const int C = 2 * Hp;
  const int nchunk = onssen_bn_num_chunks(M);
  cudaStream_t s = (cudaStream_t)stream;
  double* part = (double*)scratch;
  // scale/shift live behind the partial sums in the scratch buffer
  float* scale = (float*)(part + 2LL * nchunk * C);
  float* shift = scale + C;
  if (training) {
    dim3 grid((C + 31) / 32, nchunk);
    bn_partial_kernel<<<grid, 256, 0, s>>>(y, M, C, part);
  }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(part, nchunk, M, H, Hp, gamma, beta, running_mean,
                                                     running_var, eps, momentum, training, scale, shift,
                                                     save_mean, save_invstd);
  const long long total4 = (long long)M * C / 4;
  bn_apply_f16_kernel<<<grid_for(total4, 256), 256, 0, s>>>(y, M, C, scale, shift, (__half*)out_h);
  return ONSSEN_CHECK_LAUNCH();
}

// ---- cross-rank (SyncBN-style) batch statistics: local sums -> [caller all-reduces 2*C doubles] -> apply ----------
namespace onssen { namespace {
__global__ void chunk_sum_kernel(const double* __restrict__ part, int nchunk, int C2, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;   // over 2*C (sum rows then square rows)
  if (c >= C2) return;
  const int C = C2 / 2;
  const int half = c / C, col = c % C;
  double t = 0.0;
  for (int k = 0; k < nchunk; ++k) t += part[(long long)(half * nchunk + k) * C + col];
  sums[c] = t;
}
} }

extern "C" int onssen_bn_stats(const float* y, int M, int H, void* sums_f64, void* scratch, void* stream) {
  if (!y || !sums_f64 || !scratch || M <= 0 || H <= 0) return ONSSEN_ERR_ARG;
  const int C = 2 * hp_of(H);
  const int nchunk = onssen_bn_num_chunks(M);
  cudaStream_t s = (cudaStream_t)stream;
  bn_partial_kernel<<<dim3((C + 31) / 32, nchunk), 256, 0, s>>>(y, M, C, (double*)scratch);
  chunk_sum_kernel<<<(2 * C + 127) / 128, 128, 0, s>>>((const double*)scratch, nchunk, 2 * C, (double*)sums_f64);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_bn_forward_f16_stats(const float* y, int M, int M_total, int H, const void* sums_f64,
                                           const float* gamma, const float* beta, float* running_mean,
                                           float* running_var, float eps, float momentum, void* out_h,
                                           float* save_mean, float* save_invstd, void* scratch, void* stream) {
  if (!y || !sums_f64 || !gamma || !beta || !running_mean || !running_var || !out_h || !scratch || M <= 0 ||
      M_total < M || H <= 0)
    return ONSSEN_ERR_ARG;
  const int Hp = hp_of(H);
  const int C = 2 * Hp;
  cudaStream_t s = (cudaStream_t)stream;
  float* scale = (float*)((double*)scratch + 2LL * onssen_bn_num_chunks(M) * C);
  float* shift = scale + C;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>((const double*)sums_f64, 1, M_total, H, Hp, gamma, beta,
                                                     running_mean, running_var, eps, momentum, 1, scale, shift,
                                                     save_mean, save_invstd);
  bn_apply_f16_kernel<<<grid_for((long long)M * C / 4, 256), 256, 0, s>>>(y, M, C, scale, shift, (__half*)out_h);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_cast_f16(const float* y, long long n, void* out_h, void* stream) {
  if (!y || !out_h || n <= 0 || (n & 3)) return ONSSEN_ERR_ARG;
  cast_f16_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(y, n / 4, (__half*)out_h);
  return ONSSEN_CHECK_LAUNCH();
}
