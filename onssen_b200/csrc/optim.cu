// Optimiser step of the training loop (onssen/utils/train.py:83-84: torch.nn.utils.clip_grad_norm_(params, 5)
// followed by optimizer.step() of the torch.optim.Adam built at onssen/utils/basic.py:6-7), as three multi-tensor
// kernels over a device table of (param, grad, exp_avg, exp_avg_sq, numel) records: global squared norm ->
// clip coefficient (device scalar, no host sync) -> in-place gradient scale, and the Adam update.
// HBM-bound: clip reads every gradient once (+ one read-modify-write when it clips); Adam reads p, g, m, v and
// writes p, m, v (28 B per parameter).
#include "common.cuh"

namespace onssen {
namespace {

struct OptTensor {
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
};
struct OptChunk {
  long long tensor;
  long long start;
};

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const OptTensor* __restrict__ tensors, const OptChunk* __restrict__ chunks, int chunk_elems,
                  double* __restrict__ partials) {
  const OptChunk ck = chunks[blockIdx.x];
  const OptTensor tn = tensors[ck.tensor];
  const long long end = min(tn.n, ck.start + chunk_elems);
  const float* g = tn.g;
  float q = 0.f;
  double acc = 0.0;
  // fp32 inside a thread's strided run of <= chunk_elems/256 values, fp64 across threads/blocks
  for (long long i = ck.start + threadIdx.x; i < end; i += 256) q += g[i] * g[i];
  acc = warp_sum((double)q);
  __shared__ double sp[8];
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sp[w];
    partials[blockIdx.x] = t;
  }
}

// out2 = {total_norm, clip_coef}; clip_coef = min(1, max_norm / (total_norm + 1e-6))  (torch's clip_grad_norm_)
__global__ void clip_coef_kernel(const double* __restrict__ partials, int n, float max_norm, float* __restrict__ out2) {
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) t += partials[i];
  t = warp_sum(t);
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(t);
    out2[0] = norm;
    out2[1] = fminf(1.0f, max_norm / (norm + 1e-6f));
  }
}

__global__ void __launch_bounds__(256)
grad_scale_kernel(const OptTensor* __restrict__ tensors, const OptChunk* __restrict__ chunks, int chunk_elems,
                  const float* __restrict__ out2) {
  const float coef = out2[1];
  if (coef >= 1.0f) return;              // nothing to clip: gradients stay untouched (bit-exact with torch)
  const OptChunk ck = chunks[blockIdx.x];
  const OptTensor tn = tensors[ck.tensor];
  const long long end = min(tn.n, ck.start + chunk_elems);
  for (long long i = ck.start + threadIdx.x; i < end; i += 256) tn.g[i] *= coef;
}

// torch.optim.Adam (no amsgrad, L2 weight decay folded into the gradient like torch's default):
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(const OptTensor* __restrict__ tensors, const OptChunk* __restrict__ chunks, int chunk_elems, float lr,
            float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt) {
  const OptChunk ck = chunks[blockIdx.x];
  const OptTensor tn = tensors[ck.tensor];
  const long long end = min(tn.n, ck.start + chunk_elems);
  const float step_size = lr / bc1;
  auto upd = [&](float& pv, float g, float& m, float& v) {
    if (weight_decay != 0.f) g += weight_decay * pv;
    m = beta1 * m + (1.0f - beta1) * g;
    v = beta2 * v + (1.0f - beta2) * g * g;
    pv -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  };
  // chunk starts are multiples of chunk_elems; with 16-byte aligned tensors the body runs on float4
  const bool vec = (((uintptr_t)tn.p | (uintptr_t)tn.g | (uintptr_t)tn.m | (uintptr_t)tn.v) & 15) == 0 &&
                   (ck.start & 3) == 0;
  long long i0 = ck.start;
  if (vec) {
    const long long n4 = (end - ck.start) / 4;
    float4* p4 = reinterpret_cast<float4*>(tn.p + ck.start);
    const float4* g4 = reinterpret_cast<const float4*>(tn.g + ck.start);
    float4* m4 = reinterpret_cast<float4*>(tn.m + ck.start);
    float4* v4 = reinterpret_cast<float4*>(tn.v + ck.start);
    for (long long i = threadIdx.x; i < n4; i += 256) {
      float4 pv = p4[i], m = m4[i], v = v4[i];
      const float4 g = g4[i];
      upd(pv.x, g.x, m.x, v.x); upd(pv.y, g.y, m.y, v.y); upd(pv.z, g.z, m.z, v.z); upd(pv.w, g.w, m.w, v.w);
      p4[i] = pv; m4[i] = m; v4[i] = v;
    }
    i0 = ck.start + n4 * 4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += 256) {
    float pv = tn.p[i], m = tn.m[i], v = tn.v[i];
    upd(pv, tn.g[i], m, v);
    tn.p[i] = pv; tn.m[i] = m; tn.v[i] = v;
  }
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" int onssen_clip_grad_norm(const void* tensors, const void* chunks, int nchunks, int chunk_elems,
                                     float max_norm, void* partials_f64, float* out2, void* stream) {
  if (!tensors || !chunks || !partials_f64 || !out2 || nchunks <= 0 || chunk_elems <= 0 || !(max_norm > 0.f))
    return ONSSEN_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  grad_sumsq_kernel<<<nchunks, 256, 0, s>>>((const OptTensor*)tensors, (const OptChunk*)chunks, chunk_elems,
                                            (double*)partials_f64);
  clip_coef_kernel<<<1, 32, 0, s>>>((const double*)partials_f64, nchunks, max_norm, out2);
  grad_scale_kernel<<<nchunks, 256, 0, s>>>((const OptTensor*)tensors, (const OptChunk*)chunks, chunk_elems, out2);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_adam_step(const void* tensors, const void* chunks, int nchunks, int chunk_elems, float lr,
                                float beta1, float beta2, float eps, float weight_decay, long long step,
                                void* stream) {
  if (!tensors || !chunks || nchunks <= 0 || chunk_elems <= 0 || step <= 0) return ONSSEN_ERR_ARG;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>((const OptTensor*)tensors, (const OptChunk*)chunks,
                                                         chunk_elems, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                         (float)sqrt(bc2));
  return ONSSEN_CHECK_LAUNCH();
}
