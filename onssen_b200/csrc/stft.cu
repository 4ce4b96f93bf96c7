// STFT featurizer, ideal-binary/VAD labels and masked iSTFT overlap-add on the device.
//
// Replaces the per-item host featurizer of the reference (librosa + numpy, single thread):
//   get_stft            onssen/data/feature_utils.py:5-21   (librosa.core.stft: periodic Hann, win=n_fft,
//                                                            center=True, reflect padding, hop)
//   crop / tiling       onssen/data/wsj0_2mix.py:118-128
//   get_log_magnitude   feature_utils.py:49-51,  |.| wsj0_2mix.py:132-134
//   get_cos_difference  feature_utils.py:67-80,  get_phase feature_utils.py:54-64
//   get_one_hot         feature_utils.py:83-95
//   masked istft        egs/wsj0-2mix/deep_clustering/evaluate.py:42-45, chimera/evaluate.py:40-43
// One warp per (frame, signal) computes a spectrum (half-length complex FFT in shared memory, fp32) and writes
// every requested feature directly in the (B,T,F) layout the model consumes; only the cropped T frames are ever
// computed.  HBM-bound (waveforms in, features out), FLOPs are negligible.
#include "common.cuh"

namespace onssen {
namespace {

struct StftParams {
  const float* wav[3];
  const int32_t* crop_start;
  const int32_t* ns_per_utt;   // optional true lengths (<= ns = row pitch); NULL: every row has ns samples
  int B, ns, N, logN, hop, T, F, nsig;
  float* feature;
  float* mag[3];
  float* cosd[2];
  float* ph[3];
  float* feat_max;
};

template <bool INVERSE>
__device__ __forceinline__ void fft_radix2(float2* data, const float2* tw, int N, int tx, int nthr) {
  // decimation in time: input in bit-reversed order, output in natural order
  for (int h = 1; h < N; h <<= 1) {
    const int tstride = N / (2 * h);
    for (int t = tx; t < N / 2; t += nthr) {
      const int pos = t & (h - 1);
      const int i0 = ((t - pos) << 1) + pos;
      const int i1 = i0 + h;
      float2 w = tw[pos * tstride];
      if (INVERSE) w.y = -w.y;
      const float2 u = data[i0];
      const float2 x = data[i1];
      const float2 v = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
      data[i0] = make_float2(u.x + v.x, u.y + v.y);
      data[i1] = make_float2(u.x - v.x, u.y - v.y);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float hann_periodic(int n, int N) {
  const float s = sinpif((float)n / (float)N);
  return s * s;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float val) {
  if (val >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(val));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(val));
}

__global__ void fill_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// One WARP per (frame, signal).  A real N-point frame is transformed as ONE N/2-point complex FFT of
// z[m] = x[2m] + i x[2m+1] (radix-4 passes = two fused radix-2 stages, in the warp's private shared-memory
// buffer, __syncwarp between passes) followed by the even/odd split X[k] = E[k] + W_N^k O[k],
// X[N/2-k] = conj(E[k] - W_N^k O[k]).  A CTA of STFT_WARPS warps holds STFT_WARPS/nsig frames in flight and walks
// the (utterance, frame) list persistently; the twiddle table exp(-2 pi i k/N), k <= N/2 -- which is also the Hann
// window, 0.5 - 0.5 Re tw[k] -- is built once per CTA.  Two CTA barriers per batch of frames (spectra complete ->
// outputs, which for the cos-difference read the mixture's spectrum of the sibling warp).
constexpr int STFT_WARPS = 12;

__global__ void __launch_bounds__(STFT_WARPS * 32) stft_feat_kernel(const StftParams p, int items) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int N = p.N, M = N >> 1, logM = p.logN - 1;
  float2* tw = reinterpret_cast<float2*>(smem_raw);          // [N/2 + 1]
  float2* zall = tw + (M + 2);                               // [STFT_WARPS][M + 2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slots = STFT_WARPS / p.nsig;
  const int slot = warp / p.nsig, sig = warp - slot * p.nsig;
  float2* z = zall + (size_t)warp * (M + 2);
  const float2* z_mix = zall + (size_t)(slot * p.nsig) * (M + 2);
  for (int k = tid; k <= M; k += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  __syncthreads();
  for (int base = blockIdx.x * slots; base < items; base += gridDim.x * slots) {
    const int item = base + slot;
    const bool active = item < items && slot < slots;
    int b = 0, tt = 0;
    if (active) {
      b = item / p.T;
      tt = item - b * p.T;
      const int ns = p.ns_per_utt ? p.ns_per_utt[b] : p.ns;
      const int frames = 1 + ns / p.hop;                       // librosa center=True frame count
      const int fr = (p.crop_start[b] + tt) % frames;          // tiling of short utterances (wsj0_2mix.py:118-123)
      const float* wav = p.wav[sig] + (long long)b * p.ns;
      const int pos0 = fr * p.hop - M;
      for (int m = lane; m < M; m += 32) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = 2 * m + e;
          int pos = pos0 + n;
          if (pos < 0) pos = -pos;                             // reflect padding (librosa <= 0.9 default)
          if (pos >= ns) pos = 2 * (ns - 1) - pos;
          pos = max(0, min(ns - 1, pos));
          const float hw = 0.5f - 0.5f * tw[n <= M ? n : N - n].x;   // periodic Hann
          v[e] = __ldg(wav + pos) * hw;
        }
        z[__brev((unsigned)m) >> (32 - logM)] = make_float2(v[0], v[1]);
      }
      __syncwarp();
      // decimation in time on bit-reversed input; pass with half-size h and 2h fused (radix 4)
      int h = 1;
      for (; h * 4 <= M; h <<= 2) {
        const int ts1 = N / (2 * h), ts2 = N / (4 * h);       // twiddle strides in the N-point table
        for (int t = lane; t < M / 4; t += 32) {
          const int pos = t & (h - 1);
          const int i0 = ((t - pos) << 2) + pos;
          const float2 w1 = tw[pos * ts1], w2 = tw[pos * ts2], w3 = tw[(pos + h) * ts2];
          float2 a = z[i0], bb = z[i0 + h], c = z[i0 + 2 * h], d = z[i0 + 3 * h];
          float2 q = make_float2(bb.x * w1.x - bb.y * w1.y, bb.x * w1.y + bb.y * w1.x);
          bb = make_float2(a.x - q.x, a.y - q.y);
          a = make_float2(a.x + q.x, a.y + q.y);
          q = make_float2(d.x * w1.x - d.y * w1.y, d.x * w1.y + d.y * w1.x);
          d = make_float2(c.x - q.x, c.y - q.y);
          c = make_float2(c.x + q.x, c.y + q.y);
          q = make_float2(c.x * w2.x - c.y * w2.y, c.x * w2.y + c.y * w2.x);
          z[i0] = make_float2(a.x + q.x, a.y + q.y);
          z[i0 + 2 * h] = make_float2(a.x - q.x, a.y - q.y);
          q = make_float2(d.x * w3.x - d.y * w3.y, d.x * w3.y + d.y * w3.x);
          z[i0 + h] = make_float2(bb.x + q.x, bb.y + q.y);
          z[i0 + 3 * h] = make_float2(bb.x - q.x, bb.y - q.y);
        }
        __syncwarp();
      }
      if (h < M) {                                            // odd log2(M): one radix-2 pass left (h == M/2)
        const int ts1 = N / (2 * h);
        for (int t = lane; t < M / 2; t += 32) {
          const float2 w = tw[t * ts1];
          const float2 a = z[t], x = z[t + h];
          const float2 q = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
          z[t] = make_float2(a.x + q.x, a.y + q.y);
          z[t + h] = make_float2(a.x - q.x, a.y - q.y);
        }
        __syncwarp();
      }
      // even/odd split, in place: pair (k, M-k) belongs to one lane
      for (int k = lane; k <= M / 2; k += 32) {
        const float2 zk = z[k], zm = z[(M - k) & (M - 1)];
        const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        const float2 O = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));   // (zk - conj zm) / (2i)
        const float2 w = tw[k];
        const float2 P = make_float2(O.x * w.x - O.y * w.y, O.x * w.y + O.y * w.x);
        z[k] = make_float2(E.x + P.x, E.y + P.y);
        if (2 * k != M) z[M - k] = make_float2(E.x - P.x, -(E.y - P.y));
      }
    }
    __syncthreads();
    if (active) {
      const long long obase = ((long long)b * p.T + tt) * p.F;
      float lmax = -INFINITY;
      for (int k = lane; k < p.F; k += 32) {
        const float2 X = z[k];
        const float mg = sqrtf(X.x * X.x + X.y * X.y);
        if (p.mag[sig]) p.mag[sig][obase + k] = mg;
        if (p.ph[sig]) reinterpret_cast<float2*>(p.ph[sig])[obase + k] = X;
        if (sig == 0) {
          const float ft = log10f(mg + 1e-7f);
          if (p.feature) p.feature[obase + k] = ft;
          lmax = fmaxf(lmax, ft);
        } else if (p.cosd[sig - 1]) {
          const float2 X0 = z_mix[k];
          p.cosd[sig - 1][obase + k] = cosf(atan2f(X0.y, X0.x) - atan2f(X.y, X.x));
        }
      }
      if (p.feat_max != nullptr && sig == 0) {
        lmax = warp_max(lmax);
        if (lane == 0) atomic_max_float(p.feat_max + b, lmax);
      }
    }
    __syncthreads();
  }
}

template <typename OT>
__global__ void one_hot_kernel(const float* __restrict__ feature, const float* __restrict__ m1,
                               const float* __restrict__ m2, const float* __restrict__ fmax, float thr_off,
                               long long per_utt, long long total, OT* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_utt);
    const float thr = fmax[b] - thr_off;
    const bool active = !(feature[i] < thr);
    const bool first = m1[i] >= m2[i];   // np.argmax: ties -> index 0
    out[2 * i] = (OT)((active && first) ? 1 : 0);
    out[2 * i + 1] = (OT)((active && !first) ? 1 : 0);
  }
}

// ---------------------------------------------------------------- iSTFT
struct IstftParams {
  const float* re;
  const float* im;
  const float* mask;
  int B, S, frames, N, logN, hop, F, nsample;
  float* frames_buf;   // [B][S][frames][N]
  float* out;
};

__global__ void istft_frames_kernel(const IstftParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float2* d = reinterpret_cast<float2*>(smem_raw);  // [N]
  float2* tw = d + p.N;                              // [N/2]
  const int N = p.N, tx = threadIdx.x, nthr = blockDim.x;
  const int fr = blockIdx.x, s = blockIdx.y, b = blockIdx.z;
  const long long ibase = ((long long)b * p.frames + fr) * p.F;
  const long long mbase = (((long long)b * p.S + s) * p.frames + fr) * p.F;
  for (int k = tx; k < N; k += nthr) {
    const int kk = k <= N / 2 ? k : N - k;
    float mk = p.mask ? p.mask[mbase + kk] : 1.0f;
    float xr = p.re[ibase + kk] * mk;
    float xi = p.im[ibase + kk] * mk;
    if (kk == 0 || kk == N / 2) xi = 0.f;   // irfft ignores the imaginary part of DC / Nyquist
    if (k > N / 2) xi = -xi;
    d[__brev((unsigned)k) >> (32 - p.logN)] = make_float2(xr, xi);
  }
  for (int k = tx; k < N / 2; k += nthr) {
    float sn, cs;
    sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  __syncthreads();
  fft_radix2<true>(d, tw, N, tx, nthr);
  float* dst = p.frames_buf + ((((long long)b * p.S + s) * p.frames) + fr) * N;
  const float inv = 1.0f / (float)N;
  for (int n = tx; n < N; n += nthr) dst[n] = d[n].x * inv * hann_periodic(n, N);
}

__global__ void istft_ola_kernel(const IstftParams p) {
  const int bs = blockIdx.y;  // b*S + s
  const int N = p.N;
  const long long total_len = (long long)N + (long long)p.hop * (p.frames - 1);
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < p.nsample; n += gridDim.x * blockDim.x) {
    const long long pos = (long long)n + N / 2;
    float acc = 0.f, wss = 0.f;
    if (pos < total_len) {
      // frames i with 0 <= pos - i*hop < N
      long long i_hi = pos / p.hop;
      if (i_hi > p.frames - 1) i_hi = p.frames - 1;
      long long i_lo = (pos - N + p.hop) / p.hop;   // ceil((pos-N+1)/hop)
      if (pos - N + 1 <= 0) i_lo = 0;
      const float* fb = p.frames_buf + (long long)bs * p.frames * N;
      for (long long i = i_lo; i <= i_hi; ++i) {
        const int off = (int)(pos - i * p.hop);
        const float w = hann_periodic(off, N);
        acc += fb[i * N + off];
        wss += w * w;
      }
    }
    // librosa: divide where the window sum-square exceeds tiny
    p.out[(long long)bs * p.nsample + n] = (wss > 1.17549435e-38f) ? acc / wss : acc;
  }
}

inline int ilog2(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" int onssen_stft_features(const float* wav_mix, const float* wav_s1, const float* wav_s2, int B,
                                    int nsample, int n_fft, int hop, const int32_t* crop_start, int T,
                                    float* feature, float* mag_mix, float* mag_s1, float* mag_s2, float* cos_s1,
                                    float* cos_s2, float* ph_mix, float* ph_s1, float* ph_s2, float* feat_max,
                                    const int32_t* nsample_per_utt, void* stream) {
  if (!wav_mix || !crop_start || B <= 0 || T <= 0 || hop <= 0) return ONSSEN_ERR_ARG;
  if (n_fft < 64 || n_fft > 2048 || (n_fft & (n_fft - 1))) return ONSSEN_ERR_UNSUPPORTED;
  if (nsample <= n_fft / 2) return ONSSEN_ERR_ARG;
  const bool need_s = mag_s1 || mag_s2 || cos_s1 || cos_s2 || ph_s1 || ph_s2;
  if (need_s && (!wav_s1 || !wav_s2)) return ONSSEN_ERR_ARG;
  StftParams p;
  p.wav[0] = wav_mix; p.wav[1] = wav_s1; p.wav[2] = wav_s2;
  p.crop_start = crop_start;
  p.ns_per_utt = nsample_per_utt;
  p.B = B; p.ns = nsample; p.N = n_fft; p.logN = ilog2(n_fft); p.hop = hop; p.T = T; p.F = n_fft / 2 + 1;
  p.nsig = need_s ? 3 : 1;
  p.feature = feature;
  p.mag[0] = mag_mix; p.mag[1] = mag_s1; p.mag[2] = mag_s2;
  p.cosd[0] = cos_s1; p.cosd[1] = cos_s2;
  p.ph[0] = ph_mix; p.ph[1] = ph_s1; p.ph[2] = ph_s2;
  p.feat_max = feat_max;
  cudaStream_t s = (cudaStream_t)stream;
  if (feat_max) fill_kernel<<<(B + 255) / 256, 256, 0, s>>>(feat_max, B, -INFINITY);
  const size_t smem = (size_t)(STFT_WARPS + 1) * (n_fft / 2 + 2) * sizeof(float2);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(stft_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * (1024 + 2) * 8) !=
        cudaSuccess)
      return ONSSEN_ERR_CUDA;
    attr_set = true;
  }
  const int items = B * T, slots = STFT_WARPS / p.nsig;
  int per_sm = (int)((200u * 1024u) / smem);
  if (per_sm > 5) per_sm = 5;
  if (per_sm < 1) per_sm = 1;
  int grid = (items + slots - 1) / slots;
  if (grid > num_sms() * per_sm) grid = num_sms() * per_sm;
  stft_feat_kernel<<<grid, STFT_WARPS * 32, smem, s>>>(p, items);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_one_hot_vad(const float* feature, const float* mag_s1, const float* mag_s2,
                                  const float* feat_max, float db_threshold, int B, int T, int F, void* one_hot,
                                  int out_dtype, void* stream) {
  if (!feature || !mag_s1 || !mag_s2 || !feat_max || !one_hot || B <= 0 || T <= 0 || F <= 0)
    return ONSSEN_ERR_ARG;
  const long long per = (long long)T * F, total = per * B;
  const float off = db_threshold / 20.0f;
  long long g = (total + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  cudaStream_t s = (cudaStream_t)stream;
  switch (out_dtype) {
    case ONSSEN_DT_F32:
      one_hot_kernel<float><<<(int)g, 256, 0, s>>>(feature, mag_s1, mag_s2, feat_max, off, per, total,
                                                  (float*)one_hot);
      break;
    case ONSSEN_DT_F64:
      one_hot_kernel<double><<<(int)g, 256, 0, s>>>(feature, mag_s1, mag_s2, feat_max, off, per, total,
                                                   (double*)one_hot);
      break;
    case ONSSEN_DT_U8:
      one_hot_kernel<uint8_t><<<(int)g, 256, 0, s>>>(feature, mag_s1, mag_s2, feat_max, off, per, total,
                                                    (uint8_t*)one_hot);
      break;
    default: return ONSSEN_ERR_ARG;
  }
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" size_t onssen_istft_scratch_bytes(int B, int S, int frames, int n_fft) {
  return (size_t)B * S * frames * n_fft * sizeof(float);
}

extern "C" int onssen_istft_masked(const float* stft_re, const float* stft_im, const float* mask, int B, int S,
                                   int frames, int n_fft, int hop, int nsample, float* out, void* scratch,
                                   void* stream) {
  if (!stft_re || !stft_im || !out || !scratch || B <= 0 || S <= 0 || frames <= 0 || hop <= 0 || nsample <= 0)
    return ONSSEN_ERR_ARG;
  if (n_fft < 64 || n_fft > 2048 || (n_fft & (n_fft - 1))) return ONSSEN_ERR_UNSUPPORTED;
  IstftParams p;
  p.re = stft_re; p.im = stft_im; p.mask = mask;
  p.B = B; p.S = S; p.frames = frames; p.N = n_fft; p.logN = ilog2(n_fft); p.hop = hop; p.F = n_fft / 2 + 1;
  p.nsample = nsample; p.frames_buf = (float*)scratch; p.out = out;
  cudaStream_t s = (cudaStream_t)stream;
  const int tps = n_fft / 2 < 128 ? n_fft / 2 : 128;
  const size_t smem = (size_t)(n_fft + n_fft / 2) * sizeof(float2);
  istft_frames_kernel<<<dim3(frames, S, B), tps, smem, s>>>(p);
  int gx = (nsample + 255) / 256;
  if (gx > 1024) gx = 1024;
  istft_ola_kernel<<<dim3(gx, B * S), 256, 0, s>>>(p);
  return ONSSEN_CHECK_LAUNCH();
}
