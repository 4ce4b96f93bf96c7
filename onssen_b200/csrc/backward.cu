// Backward-pass helpers of the training path (autograd of deep_clustering.py:34-42 + train.py:81-84):
//   * amax -> power-of-two loss scale so fp32 gradients can feed the fp16 tensor-core GEMMs without underflow
//   * scaled fp32 -> fp16 cast with an optional transposed (and time-shifted) copy: wgrad GEMMs contract over
//     the row dimension M, so both operands are needed M-contiguous
//   * F.normalize backward, BatchNorm1d backward, column sums (bias grads), gradient un-permutation / un-padding
// All HBM-bound streaming / column-reduction kernels.
#include "common.cuh"

namespace onssen {
namespace {

inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

__global__ void amax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));   // positive floats order as uints
}

// scale[0] = 2^k with amax * 2^k ~= target (k clamped), scale[1] = 2^-k
__global__ void make_scale_kernel(const unsigned int* __restrict__ amax_bits, float target, float* __restrict__ scale) {
  const float a = __uint_as_float(*amax_bits);
  float k = 0.f;
  if (a > 0.f && isfinite(a)) k = floorf(log2f(target / a));
  k = fminf(fmaxf(k, -60.f), 60.f);
  scale[0] = exp2f(k);
  scale[1] = exp2f(-k);
}

// src fp32 [R][C] (row pitch ld) -> out_n fp16 [R][Cp] (zero padded) and/or out_t fp16 [C][Rp] (zero padded),
// both multiplied by scale[0] (or 1).  32x32 smem tile transpose.
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ src, int R, int C, long long ld, const float* __restrict__ scale,
                      __half* __restrict__ out_n, int Cp, __half* __restrict__ out_t, int Rp) {
  __shared__ float tile[32][33];
  const float sc = scale ? scale[0] : 1.0f;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    const float v = (r < R && c < C) ? src[(long long)r * ld + c] * sc : 0.f;
    tile[i][tx] = v;
    if (out_n != nullptr && r < R && c < Cp) out_n[(long long)r * Cp + c] = to_half_sat(v);
  }
  if (out_t == nullptr) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < Rp) out_t[(long long)c * Rp + r] = to_half_sat(tile[tx][i]);
  }
}

// fp16 src [R][C] (pitch ld) -> out_t fp16 [C][Rp] with the rows shifted: out_t[c][r] = src[r - shift][c]
// (zero outside [0,R)).  Only columns [col0, col0+ncol) are transposed (row c - col0 of out_t).
__global__ void __launch_bounds__(256)
transpose_shift_f16_kernel(const __half* __restrict__ src, int R, long long ld, int col0, int ncol, int shift,
                           __half* __restrict__ out_t, int Rp) {
  __shared__ __half tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;     // r0 indexes OUTPUT rows-of-src (after shift)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i - shift, c = c0 + tx;
    tile[i][tx] = (r >= 0 && r < R && c < ncol) ? src[(long long)r * ld + col0 + c] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < ncol && r < Rp) out_t[(long long)c * Rp + r] = tile[tx][i];
  }
}

// column sums of fp32 [R][C] (pitch ld) scaled by `mult`: two-stage, deterministic.
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, int R, int C, long long ld, double* __restrict__ part) {
  __shared__ double sp[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(R, r0 + CS_ROWS);
  double s = 0.0;
  if (c < C)
    for (int r = r0 + warp; r < r1; r += 8) s += (double)x[(long long)r * ld + c];
  sp[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < C) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sp[w][lane];
    part[(long long)blockIdx.y * C + c] = t;
  }
}
// same, 4 columns per lane (16-byte loads: a warp reads 512 contiguous bytes of a row); needs C % 4 == ld % 4 == 0
__global__ void __launch_bounds__(256)
colsum_partial4_kernel(const float* __restrict__ x, int R, int C, long long ld, double* __restrict__ part) {
  __shared__ double sp[8][32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * 4;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(R, r0 + CS_ROWS);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (c < C) {
    // fp32 inside a warp's run of <= 32 rows, fp64 across warps and chunks
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = r0 + warp; r < r1; r += 8) {
      const float4 v = *reinterpret_cast<const float4*>(x + (long long)r * ld + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    s0 = a.x; s1 = a.y; s2 = a.z; s3 = a.w;
  }
  sp[warp][lane][0] = s0; sp[warp][lane][1] = s1; sp[warp][lane][2] = s2; sp[warp][lane][3] = s3;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int l = threadIdx.x >> 2, j = threadIdx.x & 3;
    const int cc = (blockIdx.x * 32 + l) * 4 + j;
    if (cc < C) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sp[w][l][j];
      part[(long long)blockIdx.y * C + cc] = t;
    }
  }
}
__global__ void colsum_final_kernel(const double* __restrict__ part, int nchunk, int C, float mult,
                                    float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t = 0.0;
  for (int k = 0; k < nchunk; ++k) t += part[(long long)k * C + c];
  out[c] = (float)t * mult;
}

// F.normalize backward + layout change: d_emb / emb are batch-first (B,T,F,D); dz is written time-major
// [T*B][F*D] fp32:  dz = (de - e * <e,de>) * inv_norm.   One thread per (row, group); the group is read and
// written as D/4 float4 (160 contiguous bytes per thread at D=40, neighbouring threads neighbouring groups).
template <int D>
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const float* __restrict__ d_emb, const float* __restrict__ emb,
                     const float* __restrict__ inv_norm, int B, int T, int F, float* __restrict__ dz,
                     unsigned int* __restrict__ amax_bits) {
  const long long total = (long long)B * T * F;
  float am = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long bt = i / F;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const float4* e4 = reinterpret_cast<const float4*>(emb + i * D);
    const float4* g4 = reinterpret_cast<const float4*>(d_emb + i * D);
    float4 e[D / 4], g[D / 4];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < D / 4; ++k) {
      e[k] = e4[k];
      g[k] = g4[k];
      dot = fmaf(e[k].x, g[k].x, fmaf(e[k].y, g[k].y, fmaf(e[k].z, g[k].z, fmaf(e[k].w, g[k].w, dot))));
    }
    const float inv = inv_norm[i];
    float4* o = reinterpret_cast<float4*>(dz + ((long long)t * B + b) * ((long long)F * D) + (long long)f * D);
#pragma unroll
    for (int k = 0; k < D / 4; ++k) {
      float4 v;
      v.x = (g[k].x - e[k].x * dot) * inv;
      v.y = (g[k].y - e[k].y * dot) * inv;
      v.z = (g[k].z - e[k].z * dot) * inv;
      v.w = (g[k].w - e[k].w * dot) * inv;
      o[k] = v;
      am = fmaxf(am, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  }
  am = warp_max(am);
  if ((threadIdx.x & 31) == 0 && am > 0.f) atomicMax(amax_bits, __float_as_uint(am));
}

// BatchNorm1d backward (train mode, batch statistics) on the padded channel layout.
//   partial: per row chunk, sums of dA and dA*yhat per channel
__global__ void __launch_bounds__(256)
bn_bwd_partial_kernel(const float* __restrict__ d_out, const float* __restrict__ y, int M, int H, int Hp,
                      const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ part) {
  __shared__ double s0[8][32], s1[8][32];
  const int C = 2 * Hp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(M, r0 + CS_ROWS);
  double a = 0.0, bq = 0.0;
  if (c < C && (c % Hp) < H) {
    const int j = (c / Hp) * H + (c % Hp);
    const float mu = mean[j], is = invstd[j];
    for (int r = r0 + warp; r < r1; r += 8) {
      const float g = d_out[(long long)r * C + c];
      a += (double)g;
      bq += (double)g * (double)((y[(long long)r * C + c] - mu) * is);
    }
  }
  s0[warp][lane] = a;
  s1[warp][lane] = bq;
  __syncthreads();
  if (warp == 0 && c < C) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { t0 += s0[w][lane]; t1 += s1[w][lane]; }
    const int nchunk = gridDim.y;
    part[(long long)blockIdx.y * C + c] = t0;
    part[(long long)(nchunk + blockIdx.y) * C + c] = t1;
  }
}
__global__ void bn_bwd_final_kernel(const double* __restrict__ part, int nchunk, const double* __restrict__ total,
                                    int M_total, int H, int Hp, const float* __restrict__ gamma,
                                    const float* __restrict__ invstd, float* __restrict__ d_gamma,
                                    float* __restrict__ d_beta,
                                    float* __restrict__ coef /* [3][C]: a, b, c of dy = a*g + b*yhat + c */) {
  // part: this rank's per-chunk sums (-> d_gamma, d_beta); total (nullable): sums over ALL ranks [2][C] used with
  // M_total for the input gradient (cross-rank batch statistics); without it the local sums are the totals
  const int C = 2 * Hp;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int u = c % Hp;
  if (u >= H) { coef[c] = coef[C + c] = coef[2 * C + c] = 0.f; return; }
  const int j = (c / Hp) * H + u;
  double sg = 0.0, sgy = 0.0;
  for (int k = 0; k < nchunk; ++k) {
    sg += part[(long long)k * C + c];
    sgy += part[(long long)(nchunk + k) * C + c];
  }
  d_beta[j] = (float)sg;
  d_gamma[j] = (float)sgy;
  if (total != nullptr) {
    sg = total[c];
    sgy = total[C + c];
  }
  const float a = gamma[j] * invstd[j];
  coef[c] = a;
  coef[C + c] = -a * (float)(sgy / M_total);
  coef[2 * C + c] = -a * (float)(sg / M_total);
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ d_out, const float* __restrict__ y, long long M, int H,
                                    int Hp, const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ coef, float* __restrict__ d_y) {
  const int C = 2 * Hp;
  const long long total = M * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int u = c % Hp;
    float v = 0.f;
    if (u < H) {
      const int j = (c / Hp) * H + u;
      const float yh = (y[i] - mean[j]) * invstd[j];
      v = fmaf(coef[c], d_out[i], fmaf(coef[C + c], yh, coef[2 * C + c]));
    }
    d_y[i] = v;
  }
}

// padded/permuted gradient buffers -> PyTorch parameter layout (fp32), inverse of pack_* in pack.cu
__global__ void unpack_linear_grad_kernel(const float* __restrict__ gp, int N, int K, int blstm, int Hin, int Hinp,
                                          int Kp, float* __restrict__ g) {
  const long long total = (long long)N * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % K);
    const long long n = idx / K;
    const int k = blstm ? (j / Hin) * Hinp + (j % Hin) : j;
    g[idx] = gp[n * Kp + k];
  }
}
__global__ void unpack_lstm_grad_kernel(const float* __restrict__ gp, int H, int Hp, int K, int blstm, int Hin,
                                        int Hinp, int Kp, int dir, float* __restrict__ g) {
  // g [4H][K] <- gp [2*4Hp][Kp] rows of `dir`;  source row gate*H+u  <->  packed row dir*4Hp + rb*128 + 4*ul + gate
  const long long total = 4LL * H * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % K);
    const int row = (int)(idx / K);
    const int gate = row / H, u = row % H;
    const int n = dir * 4 * Hp + (u >> 5) * 128 + 4 * (u & 31) + gate;
    const int k = blstm ? (j / Hin) * Hinp + (j % Hin) : j;
    g[idx] = gp[(long long)n * Kp + k];
  }
}

// d/d mask of the PIT-L1 mask loss (loss_chimera.py:25-29,53-57): perm[b]==0 pairs (A,s1),(B,s2), else swapped.
__global__ void pit_l1_bwd_kernel(const float* __restrict__ mask_a, const float* __restrict__ mask_b, long long mstride,
                                  const float* __restrict__ mix, const float* __restrict__ s1,
                                  const float* __restrict__ s2, const float* __restrict__ c1,
                                  const float* __restrict__ c2, const int32_t* __restrict__ perm,
                                  const float* __restrict__ g, int B, long long N, float* __restrict__ d_a,
                                  float* __restrict__ d_b) {
  const long long total = (long long)B * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / N);
    const float m = mix[i];
    float t1 = s1[i], t2 = s2[i];
    if (c1 != nullptr) {
      t1 = fminf(m, fmaxf(t1 * c1[i], 0.f));
      t2 = fminf(m, fmaxf(t2 * c2[i], 0.f));
    }
    const bool sw = perm[b] != 0;
    const float ea = mask_a[i * mstride] * m - (sw ? t2 : t1);
    const float eb = mask_b[i * mstride] * m - (sw ? t1 : t2);
    const float gb = g[b] * m;
    d_a[i] = ea > 0.f ? gb : (ea < 0.f ? -gb : 0.f);      // torch.abs backward: sign(x), 0 at 0
    d_b[i] = eb > 0.f ? gb : (eb < 0.f ? -gb : 0.f);
  }
}

// sigmoid backward + layout change: d_out / out batch-first [B][T][C]; dz time-major [T*B][C]
__global__ void sigmoid_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ out, int B, int T, int C,
                                   int relu, float* __restrict__ dz, unsigned int* __restrict__ amax_bits) {
  const long long total = (long long)B * T * C;
  float am = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / C;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const float o = out[i];
    const float v = relu ? (o > 0.f ? d_out[i] : 0.f) : d_out[i] * o * (1.0f - o);
    dz[((long long)t * B + b) * C + c] = v;
    am = fmaxf(am, fabsf(v));
  }
  am = warp_max(am);
  if ((threadIdx.x & 31) == 0 && am > 0.f) atomicMax(amax_bits, __float_as_uint(am));
}

// ---- phase network (repaired phase_network.py:54-64 / loss_phase.py:26-35) ---------------------------------
// d/d(est) of  -sum_n mag * (cos(X,s1) + cos(Y,s2)),  (X,Y) = (A,B) if perm==0 else (B,A);  cos with
// F.cosine_similarity's eps clamp on each norm;  g = upstream gradient per utterance
__global__ void phase_cos_bwd_kernel(const float* __restrict__ pa, const float* __restrict__ pb,
                                     const float* __restrict__ s1, const float* __restrict__ s2,
                                     const float* __restrict__ mag, const int32_t* __restrict__ perm,
                                     const float* __restrict__ g, int B, int N, float* __restrict__ d_pa,
                                     float* __restrict__ d_pb) {
  const long long total = (long long)B * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / N);
    const bool swap = perm[b] != 0;
    const float w = -mag[i] * g[b];
    auto grad = [&](float2 x, float2 y) {   // d cos(x,y) / dx
      const float rx = sqrtf(x.x * x.x + x.y * x.y), ry = sqrtf(y.x * y.x + y.y * y.y);
      const float nx = fmaxf(rx, 1e-8f), ny = fmaxf(ry, 1e-8f);
      const float inv = 1.0f / (nx * ny);
      float2 d = make_float2(y.x * inv, y.y * inv);
      if (rx > 1e-8f) {                     // the clamp is inactive: the norm depends on x
        const float c = (x.x * y.x + x.y * y.y) * inv / (nx * nx);
        d.x -= c * x.x;
        d.y -= c * x.y;
      }
      return d;
    };
    const float2 A = reinterpret_cast<const float2*>(pa)[i], Bv = reinterpret_cast<const float2*>(pb)[i];
    const float2 S1 = reinterpret_cast<const float2*>(s1)[i], S2 = reinterpret_cast<const float2*>(s2)[i];
    const float2 da = grad(A, swap ? S2 : S1), db = grad(Bv, swap ? S1 : S2);
    reinterpret_cast<float2*>(d_pa)[i] = make_float2(w * da.x, w * da.y);
    reinterpret_cast<float2*>(d_pb)[i] = make_float2(w * db.x, w * db.y);
  }
}

// y = v / max(|v|, 1e-12), v = ph + x_phase over (re, im) pairs; d_y / ph / x_phase batch-first [B][T][F][2];
// dz (gradient w.r.t. ph, the fc_phase output) time-major [T*B][2F]
__global__ void l2norm_pairs_bwd_kernel(const float* __restrict__ d_y, const float* __restrict__ ph,
                                        const float* __restrict__ xp, int B, int T, int F, float* __restrict__ dz,
                                        unsigned int* __restrict__ amax_bits) {
  const long long total = (long long)B * T * F;
  float am = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long bt = i / F;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const float2 a = reinterpret_cast<const float2*>(ph)[i], r = reinterpret_cast<const float2*>(xp)[i];
    const float2 dy = reinterpret_cast<const float2*>(d_y)[i];
    const float re = a.x + r.x, im = a.y + r.y;
    const float nrm = sqrtf(re * re + im * im);
    const float inv = 1.0f / fmaxf(nrm, 1e-12f);
    float2 d = make_float2(dy.x * inv, dy.y * inv);
    if (nrm > 1e-12f) {
      const float yx = re * inv, yy = im * inv;
      const float dot = (yx * dy.x + yy * dy.y) * inv;
      d.x -= yx * dot;
      d.y -= yy * dot;
    }
    reinterpret_cast<float2*>(dz)[((long long)t * B + b) * F + f] = d;
    am = fmaxf(am, fmaxf(fabsf(d.x), fabsf(d.y)));
  }
  am = warp_max(am);
  if ((threadIdx.x & 31) == 0 && am > 0.f) atomicMax(amax_bits, __float_as_uint(am));
}

// second-BLSTM input = cat(x_mag * mask_s, x_phase): d_masks[b][t][f][s] += d_xin[t*B+b][f] * x_mag[b][t][f]
__global__ void phase_input_bwd_kernel(const float* __restrict__ d_xin, long long ld, const float* __restrict__ x_mag,
                                       int B, int T, int F, int S, int s_idx, float* __restrict__ d_masks) {
  const long long total = (long long)B * T * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long bt = i / F;
    const int t = (int)(bt % T), b = (int)(bt / T);
    d_masks[i * S + s_idx] += d_xin[((long long)t * B + b) * ld + f] * x_mag[i];
  }
}

// loss_mask_psa backward (loss_mask.py:25-40): d/dmask of sum_n |mask*noisy - min(noisy, relu(clean*cos))| per utterance
__global__ void l1_psa_bwd_kernel(const float* __restrict__ mask, const float* __restrict__ noisy,
                                  const float* __restrict__ clean, const float* __restrict__ cosd,
                                  const float* __restrict__ g, int B, int N, float* __restrict__ d_mask) {
  const long long total = (long long)B * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float m = noisy[i];
    const float d = mask[i] * m - fminf(m, fmaxf(clean[i] * cosd[i], 0.f));
    d_mask[i] = (d > 0.f ? m : (d < 0.f ? -m : 0.f)) * g[i / N];
  }
}

// middle of the enhancement model (enhancement.py:49-50), time-major [M][F]:  est = pre * mask, pre = relu(.),
// mask = sigmoid(.):  dz_pre = d_est * mask * (pre > 0),  dz_mi = d_est * pre * mask * (1 - mask)
__global__ void enhance_mid_bwd_kernel(const float* __restrict__ d_est, long long ld, const float* __restrict__ pre,
                                       const float* __restrict__ mask, long long M, int F, float* __restrict__ dz_pre,
                                       float* __restrict__ dz_mi, unsigned int* __restrict__ amax2) {
  const long long total = M * F;
  float a0 = 0.f, a1 = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / F;
    const int f = (int)(i % F);
    const float g = d_est[m * ld + f], pv = pre[i], mk = mask[i];
    const float v0 = pv > 0.f ? g * mk : 0.f;
    const float v1 = g * pv * mk * (1.0f - mk);
    dz_pre[i] = v0;
    dz_mi[i] = v1;
    a0 = fmaxf(a0, fabsf(v0));
    a1 = fmaxf(a1, fabsf(v1));
  }
  a0 = warp_max(a0); a1 = warp_max(a1);
  if ((threadIdx.x & 31) == 0) {
    if (a0 > 0.f) atomicMax(amax2, __float_as_uint(a0));
    if (a1 > 0.f) atomicMax(amax2 + 1, __float_as_uint(a1));
  }
}

// nn.MSELoss backward: d_a = 2 (a - b) / n * g[0]
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                               const float* __restrict__ g, float* __restrict__ d_a) {
  const float k = 2.0f / (float)n * g[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d_a[i] = (a[i] - b[i]) * k;
}

__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    a[i] += b[i];
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" int onssen_loss_pit_l1_bwd(const float* mask_a, const float* mask_b, long long mask_stride,
                                      const float* mag_mix, const float* mag_s1, const float* mag_s2,
                                      const float* cos_s1, const float* cos_s2, const int32_t* perm, const float* g,
                                      int B, int N, float* d_mask_a, float* d_mask_b, void* stream) {
  if (!mask_a || !mask_b || !mag_mix || !mag_s1 || !mag_s2 || !perm || !g || !d_mask_a || !d_mask_b) return ONSSEN_ERR_ARG;
  if ((cos_s1 == nullptr) != (cos_s2 == nullptr)) return ONSSEN_ERR_ARG;
  pit_l1_bwd_kernel<<<grid_for((long long)B * N, 256), 256, 0, (cudaStream_t)stream>>>(
      mask_a, mask_b, mask_stride, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2, perm, g, B, N, d_mask_a, d_mask_b);
  return ONSSEN_CHECK_LAUNCH();
}

static int act_bwd(const float* d_out, const float* out, int B, int T, int C, int relu, float* dz,
                   void* amax_bits_u32, void* stream) {
  if (!d_out || !out || !dz || !amax_bits_u32) return ONSSEN_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(amax_bits_u32, 0, 4, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  sigmoid_bwd_kernel<<<grid_for((long long)B * T * C, 256), 256, 0, s>>>(d_out, out, B, T, C, relu, dz,
                                                                        (unsigned int*)amax_bits_u32);
  return ONSSEN_CHECK_LAUNCH();
}
extern "C" int onssen_sigmoid_bwd(const float* d_out, const float* out, int B, int T, int C, float* dz,
                                  void* amax_bits_u32, void* stream) {
  return act_bwd(d_out, out, B, T, C, 0, dz, amax_bits_u32, stream);
}
extern "C" int onssen_relu_bwd(const float* d_out, const float* out, int B, int T, int C, float* dz,
                               void* amax_bits_u32, void* stream) {
  return act_bwd(d_out, out, B, T, C, 1, dz, amax_bits_u32, stream);
}
extern "C" int onssen_enhance_mid_bwd(const float* d_est, long long ld, const float* pre, const float* mask,
                                      long long M, int F, float* dz_pre, float* dz_mi, void* amax_bits_2xu32,
                                      void* stream) {
  if (!d_est || !pre || !mask || !dz_pre || !dz_mi || !amax_bits_2xu32 || M <= 0 || F <= 0) return ONSSEN_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(amax_bits_2xu32, 0, 8, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  enhance_mid_bwd_kernel<<<grid_for(M * F, 256), 256, 0, s>>>(d_est, ld, pre, mask, M, F, dz_pre, dz_mi,
                                                             (unsigned int*)amax_bits_2xu32);
  return ONSSEN_CHECK_LAUNCH();
}
extern "C" int onssen_loss_mse_bwd(const float* a, const float* b, long long n, const float* g, float* d_a,
                                   void* stream) {
  if (!a || !b || !g || !d_a || n <= 0) return ONSSEN_ERR_ARG;
  mse_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, g, d_a);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_loss_phase_cos_bwd(const float* phase_a, const float* phase_b, const float* phase_s1,
                                         const float* phase_s2, const float* mag_mix, const int32_t* perm,
                                         const float* g, int B, int N, float* d_phase_a, float* d_phase_b,
                                         void* stream) {
  if (!phase_a || !phase_b || !phase_s1 || !phase_s2 || !mag_mix || !perm || !g || !d_phase_a || !d_phase_b ||
      B <= 0 || N <= 0)
    return ONSSEN_ERR_ARG;
  phase_cos_bwd_kernel<<<grid_for((long long)B * N, 256), 256, 0, (cudaStream_t)stream>>>(
      phase_a, phase_b, phase_s1, phase_s2, mag_mix, perm, g, B, N, d_phase_a, d_phase_b);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_l2norm_pairs_bwd(const float* d_y, const float* x, const float* residual, int B, int T, int F,
                                       float* dz, void* amax_bits_u32, void* stream) {
  if (!d_y || !x || !residual || !dz || !amax_bits_u32 || B <= 0 || T <= 0 || F <= 0) return ONSSEN_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(amax_bits_u32, 0, 4, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  l2norm_pairs_bwd_kernel<<<grid_for((long long)B * T * F, 256), 256, 0, s>>>(d_y, x, residual, B, T, F, dz,
                                                                             (unsigned int*)amax_bits_u32);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_phase_input_bwd(const float* d_xin, long long ld, const float* x_mag, int B, int T, int F,
                                      int S, int s_idx, float* d_masks, void* stream) {
  if (!d_xin || !x_mag || !d_masks || B <= 0 || T <= 0 || F <= 0 || S <= 0 || s_idx < 0 || s_idx >= S || ld < F)
    return ONSSEN_ERR_ARG;
  phase_input_bwd_kernel<<<grid_for((long long)B * T * F, 256), 256, 0, (cudaStream_t)stream>>>(
      d_xin, ld, x_mag, B, T, F, S, s_idx, d_masks);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_loss_l1_psa_bwd(const float* mask, const float* mag_noisy, const float* mag_clean,
                                      const float* cos_diff, const float* g, int B, int N, float* d_mask,
                                      void* stream) {
  if (!mask || !mag_noisy || !mag_clean || !cos_diff || !g || !d_mask || B <= 0 || N <= 0) return ONSSEN_ERR_ARG;
  l1_psa_bwd_kernel<<<grid_for((long long)B * N, 256), 256, 0, (cudaStream_t)stream>>>(mask, mag_noisy, mag_clean,
                                                                                       cos_diff, g, B, N, d_mask);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_add_inplace(float* a, const float* b, long long n, void* stream) {
  if (!a || !b || n <= 0) return ONSSEN_ERR_ARG;
  add_inplace_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_amax_scale(const float* x, long long n, float target, void* scratch_u32, float* scale2,
                                 void* stream) {
  if (!x || !scratch_u32 || !scale2 || n <= 0) return ONSSEN_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(scratch_u32, 0, 4, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  amax_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, n, (unsigned int*)scratch_u32);
  make_scale_kernel<<<1, 1, 0, s>>>((const unsigned int*)scratch_u32, target, scale2);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_scale_from_amax_bits(const void* amax_bits_u32, float target, float* scale2, void* stream) {
  if (!amax_bits_u32 || !scale2) return ONSSEN_ERR_ARG;
  make_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned int*)amax_bits_u32, target, scale2);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_cast_transpose_f16(const float* src, int R, int C, long long ld, const float* scale2,
                                         void* out_n, int Cp, void* out_t, int Rp, void* stream) {
  if (!src || R <= 0 || C <= 0 || (!out_n && !out_t)) return ONSSEN_ERR_ARG;
  if ((out_n && Cp < C) || (out_t && Rp < R)) return ONSSEN_ERR_ARG;
  const int cmax = out_n ? (Cp > C ? Cp : C) : C;
  const int rmax = out_t ? (Rp > R ? Rp : R) : R;
  dim3 grid((cmax + 31) / 32, (rmax + 31) / 32);
  cast_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, R, C, ld, scale2, (__half*)out_n, Cp,
                                                                (__half*)out_t, Rp);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_transpose_shift_f16(const void* src, int R, long long ld, int col0, int ncol, int shift,
                                          void* out_t, int Rp, void* stream) {
  if (!src || !out_t || R <= 0 || ncol <= 0 || Rp < R) return ONSSEN_ERR_ARG;
  dim3 grid((ncol + 31) / 32, (Rp + 31) / 32);
  transpose_shift_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)src, R, ld, col0, ncol, shift,
                                                                     (__half*)out_t, Rp);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" size_t onssen_colsum_scratch_bytes(int R, int C) {
  return (size_t)((R + CS_ROWS - 1) / CS_ROWS) * C * sizeof(double) * 2;
}

extern "C" int onssen_colsum(const float* x, int R, int C, long long ld, float mult, float* out, void* scratch,
                             void* stream) {
  if (!x || !out || !scratch || R <= 0 || C <= 0) return ONSSEN_ERR_ARG;
  const int nchunk = (R + CS_ROWS - 1) / CS_ROWS;
  cudaStream_t s = (cudaStream_t)stream;
  if ((C & 3) == 0 && (ld & 3) == 0 && ((uintptr_t)x & 15) == 0)
    colsum_partial4_kernel<<<dim3((C / 4 + 31) / 32, nchunk), 256, 0, s>>>(x, R, C, ld, (double*)scratch);
  else
    colsum_partial_kernel<<<dim3((C + 31) / 32, nchunk), 256, 0, s>>>(x, R, C, ld, (double*)scratch);
  colsum_final_kernel<<<(C + 127) / 128, 128, 0, s>>>((const double*)scratch, nchunk, C, mult, out);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_normalize_bwd(const float* d_emb, const float* emb, const float* inv_norm, int B, int T,
                                    int F, int D, float* dz, void* amax_bits_u32, void* stream) {
  if (!d_emb || !emb || !inv_norm || !dz || !amax_bits_u32) return ONSSEN_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(amax_bits_u32, 0, 4, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  if ((D & 3) || ((uintptr_t)d_emb & 15) || ((uintptr_t)emb & 15) || ((uintptr_t)dz & 15)) return ONSSEN_ERR_UNSUPPORTED;
  const int grid = grid_for((long long)B * T * F, 256);
#define ONSSEN_NB(DD) \
  case DD: normalize_bwd_kernel<DD><<<grid, 256, 0, s>>>(d_emb, emb, inv_norm, B, T, F, dz, (unsigned int*)amax_bits_u32); break;
  switch (D) {
    ONSSEN_NB(4) ONSSEN_NB(8) ONSSEN_NB(12) ONSSEN_NB(16) ONSSEN_NB(20) ONSSEN_NB(24) ONSSEN_NB(32) ONSSEN_NB(40)
    default: return ONSSEN_ERR_UNSUPPORTED;
  }
#undef ONSSEN_NB
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_bn_backward(const float* d_out, const float* y, int M, int H, const float* gamma,
                                  const float* save_mean, const float* save_invstd, float* d_y, float* d_gamma,
                                  float* d_beta, void* scratch, void* stream) {
  if (!d_out || !y || !gamma || !save_mean || !save_invstd || !d_y || !d_gamma || !d_beta || !scratch)
    return ONSSEN_ERR_ARG;
  const int Hp = hp_of(H), C = 2 * Hp;
  const int nchunk = (M + CS_ROWS - 1) / CS_ROWS;
  double* part = (double*)scratch;
  float* coef = (float*)(part + 2LL * nchunk * C);
  cudaStream_t s = (cudaStream_t)stream;
  bn_bwd_partial_kernel<<<dim3((C + 31) / 32, nchunk), 256, 0, s>>>(d_out, y, M, H, Hp, save_mean, save_invstd, part);
  bn_bwd_final_kernel<<<(C + 127) / 128, 128, 0, s>>>(part, nchunk, nullptr, M, H, Hp, gamma, save_invstd, d_gamma,
                                                      d_beta, coef);
  bn_bwd_apply_kernel<<<grid_for((long long)M * C, 256), 256, 0, s>>>(d_out, y, M, H, Hp, save_mean, save_invstd,
                                                                      coef, d_y);
  return ONSSEN_CHECK_LAUNCH();
}

// cross-rank variant: onssen_bn_backward_stats -> [caller all-reduces 2*C doubles] -> onssen_bn_backward_apply
__global__ void bwd_chunk_sum_kernel(const double* __restrict__ part, int nchunk, int C2, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C2) return;
  const int C = C2 / 2;
  const int half = c / C, col = c % C;
  double t = 0.0;
  for (int k = 0; k < nchunk; ++k) t += part[(long long)(half * nchunk + k) * C + col];
  sums[c] = t;
}

extern "C" int onssen_bn_backward_stats(const float* d_out, const float* y, int M, int H, const float* save_mean,
                                        const float* save_invstd, void* sums_f64, void* scratch, void* stream) {
  if (!d_out || !y || !save_mean || !save_invstd || !sums_f64 || !scratch || M <= 0 || H <= 0) return ONSSEN_ERR_ARG;
  const int Hp = hp_of(H), C = 2 * Hp;
  const int nchunk = (M + CS_ROWS - 1) / CS_ROWS;
  cudaStream_t s = (cudaStream_t)stream;
  bn_bwd_partial_kernel<<<dim3((C + 31) / 32, nchunk), 256, 0, s>>>(d_out, y, M, H, Hp, save_mean, save_invstd,
                                                                    (double*)scratch);
  bwd_chunk_sum_kernel<<<(2 * C + 127) / 128, 128, 0, s>>>((const double*)scratch, nchunk, 2 * C, (double*)sums_f64);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_bn_backward_apply(const float* d_out, const float* y, int M, int M_total, int H,
                                        const float* gamma, const float* save_mean, const float* save_invstd,
                                        const void* sums_total_f64, float* d_y, float* d_gamma, float* d_beta,
                                        void* scratch, void* stream) {
  if (!d_out || !y || !gamma || !save_mean || !save_invstd || !sums_total_f64 || !d_y || !d_gamma || !d_beta ||
      !scratch || M <= 0 || M_total < M)
    return ONSSEN_ERR_ARG;
  const int Hp = hp_of(H), C = 2 * Hp;
  const int nchunk = (M + CS_ROWS - 1) / CS_ROWS;
  double* part = (double*)scratch;           // still holds this rank's per-chunk sums from onssen_bn_backward_stats
  float* coef = (float*)(part + 2LL * nchunk * C);
  cudaStream_t s = (cudaStream_t)stream;
  bn_bwd_final_kernel<<<(C + 127) / 128, 128, 0, s>>>(part, nchunk, (const double*)sums_total_f64, M_total, H, Hp,
                                                      gamma, save_invstd, d_gamma, d_beta, coef);
  bn_bwd_apply_kernel<<<grid_for((long long)M * C, 256), 256, 0, s>>>(d_out, y, M, H, Hp, save_mean, save_invstd,
                                                                      coef, d_y);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" size_t onssen_bn_backward_scratch_bytes(int M, int H) {
  const int C = 2 * hp_of(H);
  return (size_t)2 * ((M + CS_ROWS - 1) / CS_ROWS) * C * sizeof(double) + (size_t)3 * C * sizeof(float);
}

extern "C" int onssen_unpack_linear_grad(const float* gp, int N, int K, int in_is_blstm, int Hin, int Kp, float* g,
                                         void* stream) {
  if (!gp || !g || N <= 0 || K <= 0) return ONSSEN_ERR_ARG;
  unpack_linear_grad_kernel<<<grid_for((long long)N * K, 256), 256, 0, (cudaStream_t)stream>>>(
      gp, N, K, in_is_blstm, Hin, in_is_blstm ? hp_of(Hin) : 0, Kp, g);
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" int onssen_unpack_lstm_grad(const float* gp, int H, int K, int in_is_blstm, int Hin, int Kp, int dir,
                                       float* g, void* stream) {
  if (!gp || !g || H <= 0 || K <= 0 || dir < 0 || dir > 1) return ONSSEN_ERR_ARG;
  unpack_lstm_grad_kernel<<<grid_for(4LL * H * K, 256), 256, 0, (cudaStream_t)stream>>>(
      gp, H, hp_of(H), K, in_is_blstm, Hin, in_is_blstm ? hp_of(Hin) : 0, Kp, dir, g);
  return ONSSEN_CHECK_LAUNCH();
}
