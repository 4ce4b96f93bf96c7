// K-means on the active-bin embeddings of one utterance -> binary masks: the clustering step of deep-clustering
// inference (egs/wsj0-2mix/deep_clustering/evaluate.py:36-41: VAD at max(feature) - 40/20, sklearn
// KMeans(n_clusters=num_spk, random_state=0).fit_predict on embedding[active], mask[0] = label, mask[1] = 1 - label).
// sklearn's k-means++ draws from numpy's RNG and cannot be matched bit for bit; this is Lloyd's algorithm with a
// deterministic farthest-point initialisation, checked against sklearn's partition (up to the label permutation).
// One pass per iteration over N x D floats (HBM/L2-bound, N <= ~150 k points, D <= 64): each block accumulates
// per-cluster sums in fp64 and the last block to finish folds them into the new centroids (no host round trip).
#include "common.cuh"

namespace onssen {
namespace {

constexpr int KM_MAXK = 4;
constexpr int KM_MAXD = 64;

struct KmState {
  float cent[KM_MAXK][KM_MAXD];
  unsigned long long far_key;    // (distance bits << 32) | index, for the farthest-point searches
  unsigned int blocks_done;
  unsigned int n_active;
};

__device__ __forceinline__ bool is_active(const float* feature, const float* thr, long long n) {
  return feature == nullptr || feature[n] >= thr[0];
}

// thr[0] = max(feature) - db/20
__global__ void km_threshold_kernel(const float* __restrict__ feature, long long N, float db_over_20, float* thr) {
  float m = -INFINITY;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, feature[i]);
  __shared__ float sm[32];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sm[w]);
    thr[0] = m - db_over_20;
  }
}

// mode 0: partial sums of all active points (-> mean);  mode 1: farthest active point from st->cent[ref]
template <int MODE>
__global__ void __launch_bounds__(256)
km_init_kernel(const float* __restrict__ emb, const float* __restrict__ feature, const float* __restrict__ thr,
               long long N, int D, int ref, KmState* st, double* __restrict__ part) {
  __shared__ float c[KM_MAXD];
  if (MODE == 1)
    for (int d = threadIdx.x; d < D; d += 256) c[d] = st->cent[ref][d];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[2] = {0.0, 0.0};   // MODE 0: lane handles dims lane, lane+32
  double cnt = 0.0;
  unsigned long long best = 0ull;
  for (long long n = (long long)blockIdx.x * 8 + warp; n < N; n += (long long)gridDim.x * 8) {   // one point per warp
    if (!is_active(feature, thr, n)) continue;
    const float v0 = lane < D ? emb[n * D + lane] : 0.f;
    const float v1 = lane + 32 < D ? emb[n * D + lane + 32] : 0.f;
    if (MODE == 0) {
      acc[0] += v0; acc[1] += v1; cnt += 1.0;
    } else {
      float d0 = lane < D ? v0 - c[lane] : 0.f, d1 = lane + 32 < D ? v1 - c[lane + 32] : 0.f;
      const float dist = warp_sum(d0 * d0 + d1 * d1);
      const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned int)(0xFFFFFFFFu - (unsigned int)n);
      best = key > best ? key : best;   // ties -> smallest index
    }
  }
  if (MODE == 0) {
    __shared__ double sp[8][KM_MAXD + 1];
    sp[warp][lane] = acc[0]; sp[warp][lane + 32] = acc[1];
    if (lane == 0) sp[warp][KM_MAXD] = cnt;
    __syncthreads();
    for (int d = threadIdx.x; d <= KM_MAXD; d += 256) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += sp[w][d];
      part[(long long)blockIdx.x * (KM_MAXD + 1) + d] = t;
    }
  } else if (lane == 0 && best) {
    atomicMax(&st->far_key, best);
  }
}

__global__ void km_mean_kernel(const double* __restrict__ part, int nblk, int D, KmState* st) {
  const int d = threadIdx.x;
  double t = 0.0, c = 0.0;
  for (int b = 0; b < nblk; ++b) {
    c += part[(long long)b * (KM_MAXD + 1) + KM_MAXD];
    if (d < D) t += part[(long long)b * (KM_MAXD + 1) + d];
  }
  if (d < D) st->cent[KM_MAXK - 1][d] = c > 0 ? (float)(t / c) : 0.f;   // scratch slot: the mean
  if (d == 0) { st->n_active = (unsigned int)c; st->far_key = 0ull; }
}

__global__ void km_take_far_kernel(const float* __restrict__ emb, int D, int slot, KmState* st) {
  const unsigned int idx = st->far_key ? 0xFFFFFFFFu - (unsigned int)(st->far_key & 0xFFFFFFFFull) : 0u;   // no active point: any row
  const int d = threadIdx.x;
  const float v = d < D ? emb[(long long)idx * D + d] : 0.f;
  __syncthreads();
  if (d < D) st->cent[slot][d] = v;
  if (d == 0) st->far_key = 0ull;
}

// later seeds: farthest from the NEAREST already chosen centroid
__global__ void __launch_bounds__(256)
km_far_from_set_kernel(const float* __restrict__ emb, const float* __restrict__ feature, const float* __restrict__ thr,
                       long long N, int D, int nchosen, KmState* st) {
  __shared__ float c[KM_MAXK][KM_MAXD];
  for (int i = threadIdx.x; i < nchosen * KM_MAXD; i += 256) c[i / KM_MAXD][i % KM_MAXD] = st->cent[i / KM_MAXD][i % KM_MAXD];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long best = 0ull;
  for (long long n = (long long)blockIdx.x * 8 + warp; n < N; n += (long long)gridDim.x * 8) {
    if (!is_active(feature, thr, n)) continue;
    const float v0 = lane < D ? emb[n * D + lane] : 0.f;
    const float v1 = lane + 32 < D ? emb[n * D + lane + 32] : 0.f;
    float dmin = INFINITY;
    for (int k = 0; k < nchosen; ++k) {
      const float d0 = lane < D ? v0 - c[k][lane] : 0.f, d1 = lane + 32 < D ? v1 - c[k][lane + 32] : 0.f;
      dmin = fminf(dmin, warp_sum(d0 * d0 + d1 * d1));
    }
    const unsigned long long key = ((unsigned long long)__float_as_uint(dmin) << 32) | (unsigned int)(0xFFFFFFFFu - (unsigned int)n);
    best = key > best ? key : best;
  }
  if (lane == 0 && best) atomicMax(&st->far_key, best);
}

// one Lloyd iteration: assign + accumulate; the last block folds the partial sums into the new centroids.
// With write_masks the labels of this (final) assignment are written as masks[k][n].
__global__ void __launch_bounds__(256)
km_lloyd_kernel(const float* __restrict__ emb, const float* __restrict__ feature, const float* __restrict__ thr,
                long long N, int D, int K, KmState* st, double* __restrict__ part, float* __restrict__ masks,
                int32_t* __restrict__ labels) {
  __shared__ float c[KM_MAXK][KM_MAXD];
  __shared__ double sp[8][KM_MAXK][KM_MAXD + 1];
  for (int i = threadIdx.x; i < K * KM_MAXD; i += 256) c[i / KM_MAXD][i % KM_MAXD] = st->cent[i / KM_MAXD][i % KM_MAXD];
  for (int i = threadIdx.x; i < 8 * KM_MAXK * (KM_MAXD + 1); i += 256) (&sp[0][0][0])[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long n = (long long)blockIdx.x * 8 + warp; n < N; n += (long long)gridDim.x * 8) {
    const bool act = is_active(feature, thr, n);
    int lab = -1;
    if (act) {
      const float v0 = lane < D ? emb[n * D + lane] : 0.f;
      const float v1 = lane + 32 < D ? emb[n * D + lane + 32] : 0.f;
      float dmin = INFINITY;
      for (int k = 0; k < K; ++k) {
        const float d0 = lane < D ? v0 - c[k][lane] : 0.f, d1 = lane + 32 < D ? v1 - c[k][lane + 32] : 0.f;
        const float dist = warp_sum(d0 * d0 + d1 * d1);
        if (dist < dmin) { dmin = dist; lab = k; }     // ties -> lowest cluster index
      }
      sp[warp][lab][lane] += v0;
      sp[warp][lab][lane + 32] += v1;
      if (lane == 0) sp[warp][lab][KM_MAXD] += 1.0;
    }
    if (lane == 0) {
      if (labels) labels[n] = lab;
      if (masks) {
        // evaluate.py:39-41 for two speakers: mask[0] = label, mask[1] = 1 - label on active bins, 0 elsewhere;
        // for K > 2 the natural extension mask[k] = (label == k)
        for (int k = 0; k < K; ++k)
          masks[(long long)k * N + n] = !act ? 0.f : (K == 2 ? (k == 0 ? (float)lab : 1.0f - (float)lab) : (lab == k ? 1.f : 0.f));
      }
    }
  }
  __syncthreads();
  const int per = KM_MAXK * (KM_MAXD + 1);
  for (int i = threadIdx.x; i < per; i += 256) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (&sp[w][0][0])[i];
    part[(long long)blockIdx.x * per + i] = t;
  }
  __threadfence();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&st->blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int i = threadIdx.x; i < K * KM_MAXD; i += 256) {
    const int k = i / KM_MAXD, d = i % KM_MAXD;
    if (d >= D) continue;
    double t = 0.0, cn = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) {
      t += part[(long long)b * per + k * (KM_MAXD + 1) + d];
      cn += part[(long long)b * per + k * (KM_MAXD + 1) + KM_MAXD];
    }
    if (cn > 0.0) st->cent[k][d] = (float)(t / cn);       // an empty cluster keeps its centroid
  }
  if (threadIdx.x == 0) st->blocks_done = 0u;
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" size_t onssen_kmeans_scratch_bytes(void) {
  const int nblk = num_sms() * 2;
  return sizeof(KmState) + 16 + (size_t)nblk * KM_MAXK * (KM_MAXD + 1) * sizeof(double);
}

extern "C" int onssen_kmeans_masks(const float* emb, const float* feature, long long N, int D, int K,
                                   float db_threshold, int iters, float* masks, int32_t* labels, void* scratch,
                                   void* stream) {
  if (!emb || N <= 0 || D <= 0 || K < 2 || iters <= 0 || !scratch || (!masks && !labels)) return ONSSEN_ERR_ARG;
  if (D > KM_MAXD || K > KM_MAXK - 1 || N > 0xFFFFFFFEll) return ONSSEN_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  KmState* st = (KmState*)scratch;
  float* thr = (float*)((uint8_t*)scratch + sizeof(KmState));
  double* part = (double*)((uint8_t*)scratch + sizeof(KmState) + 16);
  const int nblk = num_sms() * 2;
  if (cudaMemsetAsync(st, 0, sizeof(KmState), s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  if (feature) km_threshold_kernel<<<1, 1024, 0, s>>>(feature, N, db_threshold / 20.0f, thr);
  // seeds: farthest active point from the mean, then repeatedly the point farthest from the chosen set
  km_init_kernel<0><<<nblk, 256, 0, s>>>(emb, feature, thr, N, D, 0, st, part);
  km_mean_kernel<<<1, KM_MAXD, 0, s>>>(part, nblk, D, st);
  km_init_kernel<1><<<nblk, 256, 0, s>>>(emb, feature, thr, N, D, KM_MAXK - 1, st, part);
  km_take_far_kernel<<<1, KM_MAXD, 0, s>>>(emb, D, 0, st);
  for (int k = 1; k < K; ++k) {
    km_far_from_set_kernel<<<nblk, 256, 0, s>>>(emb, feature, thr, N, D, k, st);
    km_take_far_kernel<<<1, KM_MAXD, 0, s>>>(emb, D, k, st);
  }
  for (int it = 0; it < iters; ++it)
    km_lloyd_kernel<<<nblk, 256, 0, s>>>(emb, feature, thr, N, D, K, st, part, nullptr, nullptr);
  km_lloyd_kernel<<<nblk, 256, 0, s>>>(emb, feature, thr, N, D, K, st, part, masks, labels);
  return ONSSEN_CHECK_LAUNCH();
}
