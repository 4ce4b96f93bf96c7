// PCM decode and polyphase resampling on the device: the numeric part of
//   librosa.load(fn, sr=None)            onssen/data/feature_utils.py:15   (int16 -> float32 / 32768, channel mean)
//   librosa.core.resample(sig, fs, sr)   onssen/data/feature_utils.py:19   (here: the polyphase FIR of
//                                        scipy.signal.resample_poly; librosa's resampy kernel is a different
//                                        low-pass -> resampled audio is "parity unpinned", see data/wavio.py)
// Both are streaming kernels: bytes = input once + output once.
#include "common.cuh"

namespace onssen {
namespace {

inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// pcm [R][pitch*ch] int16 (interleaved channels), frames[r] valid frames of row r -> out [R][pitch] float32 (0 beyond)
__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ pcm, int R, int pitch, int ch,
                                    const int32_t* __restrict__ frames, float* __restrict__ out) {
  const long long total = (long long)R * pitch;
  const float inv_ch = 1.0f / (float)ch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / pitch), n = (int)(i % pitch);
    float v = 0.f;
    if (n < frames[r]) {
      const int16_t* src = pcm + ((long long)r * pitch + n) * ch;
      if (ch == 1) {
        v = (float)src[0] * (1.0f / 32768.0f);
      } else {
        float s = 0.f;
        for (int c = 0; c < ch; ++c) s += (float)src[c] * (1.0f / 32768.0f);
        v = s * inv_ch;
      }
    }
    out[i] = v;
  }
}

// y[r][j] = sum_i x[r][i] * h[(j + pre) * down - i * up]   (upfirdn of scipy.signal.resample_poly), j < ceil(n_in*up/down)
__global__ void resample_poly_kernel(const float* __restrict__ x, int R, int pitch_in, const int32_t* __restrict__ n_in,
                                     int up, int down, const float* __restrict__ h, int hlen, int pre,
                                     float* __restrict__ y, int pitch_out) {
  const long long total = (long long)R * pitch_out;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / pitch_out), j = (int)(idx % pitch_out);
    const int n = n_in[r];
    const long long n_out = ((long long)n * up + down - 1) / down;
    float acc = 0.f;
    if (j < n_out) {
      const long long t = (long long)(j + pre) * down;
      long long i_hi = t / up;
      if (i_hi > n - 1) i_hi = n - 1;
      long long i_lo = t - hlen + 1 <= 0 ? 0 : (t - hlen + 1 + up - 1) / up;
      const float* xr = x + (long long)r * pitch_in;
      for (long long i = i_lo; i <= i_hi; ++i) acc = fmaf(xr[i], __ldg(h + (t - i * up)), acc);
    }
    y[idx] = acc;
  }
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" int onssen_pcm16_to_f32(const void* pcm_i16, int R, int pitch, int channels, const int32_t* frames,
                                   float* out, void* stream) {
  if (!pcm_i16 || !frames || !out || R <= 0 || pitch <= 0 || channels <= 0) return ONSSEN_ERR_ARG;
  pcm16_to_f32_kernel<<<grid_for((long long)R * pitch, 256), 256, 0, (cudaStream_t)stream>>>(
      (const int16_t*)pcm_i16, R, pitch, channels, frames, out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_resample_poly(const float* x, int R, int pitch_in, const int32_t* n_in, int up, int down,
                                    const float* h, int hlen, int n_pre_remove, float* y, int pitch_out, void* stream) {
  if (!x || !n_in || !h || !y || R <= 0 || pitch_in <= 0 || pitch_out <= 0 || up <= 0 || down <= 0 || hlen <= 0 ||
      n_pre_remove < 0)
    return ONSSEN_ERR_ARG;
  resample_poly_kernel<<<grid_for((long long)R * pitch_out, 256), 256, 0, (cudaStream_t)stream>>>(
      x, R, pitch_in, n_in, up, down, h, hlen, n_pre_remove, y, pitch_out);
  return ONSSEN_CHECK_LAUNCH();
}
