// Small streaming kernels for the enhancement / phase-network variants of the path:
//   enhance   (onssen/nn/enhancement.py:48-51): mask * relu(fc_pre(mag_noisy)) -> fp16 operand of fc_post
//   phase_net (onssen/nn/phase_network.py:46-66, repaired): cat(x_mag*mask, x_phase) -> fp16 operand of the
//             second BLSTM; fc_phase output + x_phase -> F.normalize over the (re,im) pair
//   loss_mask_msa / loss_mask_psa (onssen/loss/loss_mask.py:6-40), phase term of loss_phase
//             (onssen/loss/loss_phase.py:26-35)
// All HBM-bound, one read of each input and one write of each output.
#include "common.cuh"

namespace onssen {
namespace {

inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

__global__ void mul_pack_f16_kernel(const float* __restrict__ a, const float* __restrict__ b, long long M, int F,
                                    int Kp, __half* __restrict__ out) {
  const long long total = M * Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const long long m = idx / Kp;
    out[idx] = to_half_sat(k < F ? a[m * F + k] * b[m * F + k] : 0.f);
  }
}

// rows time-major m = t*B + b; columns [0,F): x_mag*mask_s, [F,3F): x_phase (re,im interleaved), rest 0
__global__ void pack_phase_input_kernel(const float* __restrict__ x_mag, const float* __restrict__ mask,
                                        long long mask_stride, const float* __restrict__ x_phase, int B, int T,
                                        int F, int Kp, __half* __restrict__ out) {
  const long long total = (long long)B * T * Kp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kp);
    const long long m = idx / Kp;
    const int b = (int)(m % B), t = (int)(m / B);
    const long long bt = (long long)b * T + t;
    float v = 0.f;
    if (k < F) v = x_mag[bt * F + k] * mask[(bt * F + k) * mask_stride];
    else if (k < 3 * F) v = x_phase[bt * 2 * F + (k - F)];
    out[idx] = to_half_sat(v);
  }
}

__global__ void add_l2norm_pairs_kernel(const float* __restrict__ x, const float* __restrict__ res, long long n2,
                                        float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    const float2 a = reinterpret_cast<const float2*>(x)[i];
    const float2 r = reinterpret_cast<const float2*>(res)[i];
    const float re = a.x + r.x, im = a.y + r.y;
    const float inv = 1.0f / fmaxf(sqrtf(re * re + im * im), 1e-12f);
    reinterpret_cast<float2*>(out)[i] = make_float2(re * inv, im * inv);
  }
}

// per-utterance reduction helper: block per utterance, fp64 accumulation, fixed order
template <typename F>
__device__ __forceinline__ double block_reduce_utt(int N, F f) {
  double q = 0.0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) q += (double)f(n);
  __shared__ double sp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  q = warp_sum(q);
  if (lane == 0) sp[warp] = q;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sp[w];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(1024)
l1_psa_kernel(const float* __restrict__ mask, const float* __restrict__ noisy, const float* __restrict__ clean,
              const float* __restrict__ cosd, int N, float* __restrict__ out) {
  const long long off = (long long)blockIdx.x * N;
  const double t = block_reduce_utt(N, [&](int n) {
    const float m = noisy[off + n];
    const float tgt = fminf(m, fmaxf(clean[off + n] * cosd[off + n], 0.f));
    return fabsf(mask[off + n] * m - tgt);
  });
  if (threadIdx.x == 0) out[blockIdx.x] = (float)t;
}

__global__ void __launch_bounds__(1024)
sqdiff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                      double* __restrict__ part) {
  double q = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    q += (double)d * d;
  }
  __shared__ double sp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  q = warp_sum(q);
  if (lane == 0) sp[warp] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sp[w];
    part[blockIdx.x] = t;
  }
}
__global__ void mse_final_kernel(const double* __restrict__ part, int nblk, long long n, float* __restrict__ out) {
  double t = 0.0;
  for (int i = 0; i < nblk; ++i) t += part[i];
  out[0] = (float)(t / (double)n);
}

// -sum_n mag * (cos(pX,s1) + cos(pY,s2)) with (X,Y) = (A,B) if perm==0 else (B,A); F.cosine_similarity eps 1e-8
__global__ void __launch_bounds__(1024)
phase_cos_kernel(const float* __restrict__ pa, const float* __restrict__ pb, const float* __restrict__ s1,
                 const float* __restrict__ s2, const float* __restrict__ mag, const int32_t* __restrict__ perm,
                 int N, float* __restrict__ out) {
  const long long off = (long long)blockIdx.x * N;
  const bool swap = perm[blockIdx.x] != 0;
  const float2* A = reinterpret_cast<const float2*>(swap ? pb : pa) + off;
  const float2* Bv = reinterpret_cast<const float2*>(swap ? pa : pb) + off;
  const float2* S1 = reinterpret_cast<const float2*>(s1) + off;
  const float2* S2 = reinterpret_cast<const float2*>(s2) + off;
  auto cs = [](float2 x, float2 y) {
    const float nx = fmaxf(sqrtf(x.x * x.x + x.y * x.y), 1e-8f), ny = fmaxf(sqrtf(y.x * y.x + y.y * y.y), 1e-8f);
    return (x.x * y.x + x.y * y.y) / (nx * ny);
  };
  const double t = block_reduce_utt(N, [&](int n) { return -mag[off + n] * (cs(A[n], S1[n]) + cs(Bv[n], S2[n])); });
  if (threadIdx.x == 0) out[blockIdx.x] = (float)t;
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" int onssen_mul_pack_f16(const float* a, const float* b, long long M, int F, void* out, int Kp,
                                   void* stream) {
  if (!a || !b || !out || M <= 0 || F <= 0 || Kp < F || (Kp & 7)) return ONSSEN_ERR_ARG;
  mul_pack_f16_kernel<<<grid_for(M * Kp, 256), 256, 0, (cudaStream_t)stream>>>(a, b, M, F, Kp, (__half*)out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_pack_phase_input_f16(const float* x_mag, const float* mask, long long mask_stride,
                                           const float* x_phase, int B, int T, int F, void* out, int Kp,
                                           void* stream) {
  if (!x_mag || !mask || !x_phase || !out || B <= 0 || T <= 0 || F <= 0 || Kp < 3 * F || (Kp & 7) || mask_stride <= 0)
    return ONSSEN_ERR_ARG;
  pack_phase_input_kernel<<<grid_for((long long)B * T * Kp, 256), 256, 0, (cudaStream_t)stream>>>(
      x_mag, mask, mask_stride, x_phase, B, T, F, Kp, (__half*)out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_add_l2norm_pairs(const float* x, const float* residual, long long npairs, float* out,
                                       void* stream) {
  if (!x || !residual || !out || npairs <= 0) return ONSSEN_ERR_ARG;
  add_l2norm_pairs_kernel<<<grid_for(npairs, 256), 256, 0, (cudaStream_t)stream>>>(x, residual, npairs, out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_loss_l1_psa_fwd(const float* mask, const float* mag_noisy, const float* mag_clean,
                                      const float* cos_diff, int B, int N, float* out, void* stream) {
  if (!mask || !mag_noisy || !mag_clean || !cos_diff || !out || B <= 0 || N <= 0) return ONSSEN_ERR_ARG;
  l1_psa_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(mask, mag_noisy, mag_clean, cos_diff, N, out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_loss_mse_fwd(const float* a, const float* b, long long n, float* out, void* scratch,
                                   void* stream) {
  if (!a || !b || !out || !scratch || n <= 0) return ONSSEN_ERR_ARG;
  const int nblk = 256;   // scratch: 256 doubles
  sqdiff_partial_kernel<<<nblk, 1024, 0, (cudaStream_t)stream>>>(a, b, n, (double*)scratch);
  mse_final_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const double*)scratch, nblk, n, out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_loss_phase_cos_fwd(const float* phase_a, const float* phase_b, const float* phase_s1,
                                         const float* phase_s2, const float* mag_mix, const int32_t* perm, int B,
                                         int N, float* out, void* stream) {
  if (!phase_a || !phase_b || !phase_s1 || !phase_s2 || !mag_mix || !perm || !out || B <= 0 || N <= 0)
    return ONSSEN_ERR_ARG;
  phase_cos_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(phase_a, phase_b, phase_s1, phase_s2, mag_mix, perm, N, out);
  return ONSSEN_CHECK_LAUNCH();
}
