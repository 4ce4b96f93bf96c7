// Loss reductions of the STFT-mask separation path.
//
// loss_dc (onssen/loss/loss_dc.py:6-44): with a_n = (sum_k y_nk) * e_n and per-point weight
// w_n^2 = m_n / sum(m), the reference forms three weighted Gram matrices by bmm and takes UN-squared
// Frobenius norms:  l = ||A^T A|| - 2||A^T Y|| + ||Y^T Y||, returned as the (B,B) outer product sum(m)_i*l_j.
// Here one streaming pass over the embedding accumulates sum_n m_n a a^T (DxD), sum_n m_n a y^T (DxS),
// sum_n m_n y y^T (SxS) and sum_n m_n in registers (the 1/sum(m) factor is linear and applied at the end),
// so the 264 MB embedding of BASELINE cfg2 is read exactly once.  fp32 FMA on CUDA cores: at D=40 the
// kernel sits at the fp32/HBM ridge (20 flop/B), tensor cores would need tf32 rounding of the parity-critical
// loss operand, see DESIGN.md section 6.
//
// PIT L1 mask loss (onssen/loss/loss_chimera.py:25-29,53-57): four per-utterance L1 sums, min over the
// two speaker permutations.
#include "common.cuh"

namespace onssen {
namespace {

constexpr int DC_THREADS = 256;
constexpr int DC_PTS_PER_CHUNK = 1024;

template <typename LT>
__device__ __forceinline__ float lab_to_f(LT v) { return (float)v; }

// ---- fast path: S == 2, D = TS * TG with TG*TG | 256 ----
template <int D, int TS, typename LT>
__global__ void __launch_bounds__(DC_THREADS)
loss_dc_partial_kernel(const float* __restrict__ emb, const LT* __restrict__ label,
                       const float* __restrict__ mag, int N, float* __restrict__ scratch) {
  // only the tiles on or above the diagonal of the symmetric Gram are accumulated (tj >= ti); the final kernel
  // counts off-diagonal tiles twice.  GT threads per group, NG groups split the points of a tile.
  constexpr int TG = D / TS;
  constexpr int GT = TG * (TG + 1) / 2;
  constexpr int NG = DC_THREADS / GT;
  constexpr int P = NG > 64 ? NG : 64;
  constexpr int R = D * D + D * 2 + 4;
  static_assert(TG * TS == D && NG >= 1, "unsupported D");
  extern __shared__ __align__(16) float sm[];
  float* a_s = sm;                 // [P][D]   s_n * e_n
  float* ma_s = a_s + P * D;       // [P][D]   m_n * s_n * e_n
  float* y_s = ma_s + P * D;       // [P][2]
  float* m_s = y_s + P * 2;        // [P]
  float* red = sm;                 // reused after the main loop: [NG][R]

  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int nchunk = gridDim.x;
  const int n_begin = chunk * DC_PTS_PER_CHUNK;
  const int n_end = min(N, n_begin + DC_PTS_PER_CHUNK);
  const int tid = threadIdx.x;
  const int g = tid / GT;          // g >= NG: idle in the accumulation (still stages data / syncs)
  const int lt = tid % GT;
  int ti = 0, tj;
  {
    int rem = lt;
    while (rem >= TG - ti) { rem -= TG - ti; ++ti; }
    tj = ti + rem;
  }
  const bool diag = (ti == tj);

  float acc[TS][TS];
  float accy[TS][2];
  float accs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < TS; ++i) {
    accy[i][0] = accy[i][1] = 0.f;
#pragma unroll
    for (int j = 0; j < TS; ++j) acc[i][j] = 0.f;
  }
  const float* eb = emb + (long long)b * N * D;
  const LT* lb = label + (long long)b * N * 2;
  const float* mb = mag + (long long)b * N;

  for (int n0 = n_begin; n0 < n_end; n0 += P) {
    const int np = min(P, n_end - n0);
    for (int i = tid; i < np; i += DC_THREADS) {
      y_s[2 * i] = lab_to_f(lb[2 * (long long)(n0 + i)]);
      y_s[2 * i + 1] = lab_to_f(lb[2 * (long long)(n0 + i) + 1]);
      m_s[i] = mb[n0 + i];
    }
    __syncthreads();
    for (int i = tid; i < np * D / 4; i += DC_THREADS) {
      const int pnt = (i * 4) / D;
      const float4 e = reinterpret_cast<const float4*>(eb + (long long)n0 * D)[i];
      const float s = y_s[2 * pnt] + y_s[2 * pnt + 1];
      reinterpret_cast<float4*>(a_s)[i] = make_float4(e.x * s, e.y * s, e.z * s, e.w * s);
      // m * (s*e): keep the reference's association (mask first, then weight)
      reinterpret_cast<float4*>(ma_s)[i] = make_float4(e.x * s * m_s[pnt], e.y * s * m_s[pnt],
                                                       e.z * s * m_s[pnt], e.w * s * m_s[pnt]);
    }
    __syncthreads();
    for (int pnt = g; pnt < np && g < NG; pnt += NG) {
      float rv[TS], cv[TS];
      if constexpr ((TS & 1) == 0) {   // D*4 and TS*4 bytes are multiples of 8: 64-bit smem loads
#pragma unroll
        for (int i = 0; i < TS; i += 2) {
          const float2 r2 = *reinterpret_cast<const float2*>(ma_s + pnt * D + ti * TS + i);
          const float2 c2 = *reinterpret_cast<const float2*>(a_s + pnt * D + tj * TS + i);
          rv[i] = r2.x; rv[i + 1] = r2.y;
          cv[i] = c2.x; cv[i + 1] = c2.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TS; ++i) {
          rv[i] = ma_s[pnt * D + ti * TS + i];
          cv[i] = a_s[pnt * D + tj * TS + i];
        }
      }
#pragma unroll
      for (int i = 0; i < TS; ++i)
#pragma unroll
        for (int j = 0; j < TS; ++j) acc[i][j] = fmaf(rv[i], cv[j], acc[i][j]);
      if (diag) {
        const float y0 = y_s[2 * pnt], y1 = y_s[2 * pnt + 1];
#pragma unroll
        for (int i = 0; i < TS; ++i) {
          accy[i][0] = fmaf(rv[i], y0, accy[i][0]);
          accy[i][1] = fmaf(rv[i], y1, accy[i][1]);
        }
        if (ti == 0) {
          const float m = m_s[pnt];
          accs[0] = fmaf(m * y0, y0, accs[0]);
          accs[1] = fmaf(m * y0, y1, accs[1]);
          accs[2] = fmaf(m * y1, y1, accs[2]);
          accs[3] += m;
        }
      }
    }
    __syncthreads();
  }
  // cross-group reduction in a fixed order (deterministic); entries below the diagonal tiles stay zero
  for (int i = tid; i < NG * R; i += DC_THREADS) red[i] = 0.f;
  __syncthreads();
  float* mine = red + (g < NG ? g : 0) * R;
  if (g < NG) {
#pragma unroll
  for (int i = 0; i < TS; ++i)
#pragma unroll
    for (int j = 0; j < TS; ++j) mine[(ti * TS + i) * D + tj * TS + j] = acc[i][j];
  }
  if (diag && g < NG) {
#pragma unroll
    for (int i = 0; i < TS; ++i) {
      mine[D * D + (ti * TS + i) * 2] = accy[i][0];
      mine[D * D + (ti * TS + i) * 2 + 1] = accy[i][1];
    }
    if (ti == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mine[D * D + D * 2 + i] = accs[i];
    }
  }
  __syncthreads();
  float* out = scratch + ((long long)b * nchunk + chunk) * R;
  for (int i = tid; i < R; i += DC_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) s += red[gg * R + i];
    out[i] = s;
  }
}

template <int D, int TS>
constexpr size_t dc_smem_bytes() {
  constexpr int TG = D / TS, GT = TG * (TG + 1) / 2, NG = DC_THREADS / GT, P = NG > 64 ? NG : 64;
  constexpr size_t tile = (size_t)(2 * P * D + 3 * P) * 4;
  constexpr size_t red = (size_t)NG * (D * D + D * 2 + 4) * 4;
  return tile > red ? tile : red;
}

// ---- generic path: any D <= 128, S <= 4 (one thread per output entry, slow but exact same maths) ----
template <typename LT>
__global__ void __launch_bounds__(DC_THREADS)
loss_dc_partial_generic(const float* __restrict__ emb, const LT* __restrict__ label,
                        const float* __restrict__ mag, int N, int D, int S, float* __restrict__ scratch) {
  constexpr int P = 32;
  extern __shared__ __align__(16) float sm[];
  float* a_s = sm;              // [P][D]
  float* y_s = a_s + P * D;     // [P][S]
  float* m_s = y_s + P * S;     // [P]
  const int R = D * D + D * S + S * S + 1;
  const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int n_begin = chunk * DC_PTS_PER_CHUNK;
  const int n_end = min(N, n_begin + DC_PTS_PER_CHUNK);
  const float* eb = emb + (long long)b * N * D;
  const LT* lb = label + (long long)b * N * S;
  const float* mb = mag + (long long)b * N;
  float* out = scratch + ((long long)b * nchunk + chunk) * R;
  for (int i = threadIdx.x; i < R; i += DC_THREADS) out[i] = 0.f;
  for (int n0 = n_begin; n0 < n_end; n0 += P) {
    const int np = min(P, n_end - n0);
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += DC_THREADS) {
      float s = 0.f;
      for (int k = 0; k < S; ++k) {
        const float y = lab_to_f(lb[(long long)(n0 + i) * S + k]);
        y_s[i * S + k] = y;
        s += y;
      }
      m_s[i] = mb[n0 + i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < np * D; i += DC_THREADS) {
      const int pnt = i / D;
      float s = 0.f;
      for (int k = 0; k < S; ++k) s += y_s[pnt * S + k];
      a_s[i] = eb[(long long)n0 * D + i] * s;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < R; o += DC_THREADS) {
      float accv = out[o];
      if (o < D * D) {
        const int i = o / D, j = o % D;
        for (int pnt = 0; pnt < np; ++pnt) accv = fmaf(a_s[pnt * D + i] * m_s[pnt], a_s[pnt * D + j], accv);
      } else if (o < D * D + D * S) {
        const int i = (o - D * D) / S, k = (o - D * D) % S;
        for (int pnt = 0; pnt < np; ++pnt) accv = fmaf(a_s[pnt * D + i] * m_s[pnt], y_s[pnt * S + k], accv);
      } else if (o < D * D + D * S + S * S) {
        const int k = (o - D * D - D * S) / S, l = (o - D * D - D * S) % S;
        for (int pnt = 0; pnt < np; ++pnt) accv = fmaf(m_s[pnt] * y_s[pnt * S + k], y_s[pnt * S + l], accv);
      } else {
        for (int pnt = 0; pnt < np; ++pnt) accv += m_s[pnt];
      }
      out[o] = accv;
    }
  }
}

// ordered (deterministic) sum of the per-chunk partial records: grid (ceil(R/256), B) -> summed[b][R]
__global__ void __launch_bounds__(256) loss_dc_reduce_kernel(const float* __restrict__ scratch, int nchunk, int R,
                                                             float* __restrict__ summed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= R) return;
  const float* base = scratch + (long long)b * nchunk * R + i;
  float v = 0.f;
  for (int c = 0; c < nchunk; ++c) v += base[(long long)c * R];
  summed[(long long)b * R + i] = v;
}

// per-utterance: norms, l_b and sum(m)_b from the summed record (nchunk == 1 layout)
__global__ void __launch_bounds__(256) loss_dc_final_kernel(const float* __restrict__ scratch, int nchunk, int D,
                                                            int S, int fast_layout, int sym_ts,
                                                            float* __restrict__ l_out,
                                                            float* __restrict__ msum_out) {
  const int b = blockIdx.x;
  // layout of one partial record
  const int nG = D * D, nC = D * S;
  const int nY = fast_layout ? 3 : S * S;
  const int R = fast_layout ? (D * D + D * 2 + 4) : (D * D + D * S + S * S + 1);
  __shared__ double s_part[3][8];
  __shared__ float s_msum;
  double q[3] = {0.0, 0.0, 0.0};
  const float* base = scratch + (long long)b * nchunk * R;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    float v = 0.f;
    for (int c = 0; c < nchunk; ++c) v += base[(long long)c * R + i];
    if (i < nG) {
      // symmetric fast layout: only tiles with tile(col) >= tile(row) are stored; off-diagonal tiles count twice
      const double w = (sym_ts > 0 && (i / D) / sym_ts < (i % D) / sym_ts) ? 2.0 : 1.0;
      q[0] += w * (double)v * v;
    } else if (i < nG + nC) {
      q[1] += (double)v * v;
    } else if (i < nG + nC + nY) {
      double w = 1.0;
      if (fast_layout && (i - nG - nC) == 1) w = 2.0;  // off-diagonal y0*y1 appears twice in Y^T Y
      q[2] += w * (double)v * v;
    } else {
      s_msum = v;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 3; ++k) {
    const double r = warp_sum(q[k]);
    if (lane == 0) s_part[k][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < 3; ++k)
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t[k] += s_part[k][w];
    const float msum = s_msum;
    const float ne = sqrtf((float)t[0]) / msum;
    const float ney = sqrtf((float)t[1]) / msum;
    const float ny = sqrtf((float)t[2]) / msum;
    const float l = ne - 2.0f * ney + ny;
    if (l_out) l_out[b] = l;
    if (msum_out) msum_out[b] = msum;
  }
}

__global__ void loss_dc_outer_kernel(const float* __restrict__ l, const float* __restrict__ msum, int B,
                                     float* __restrict__ loss_bb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * B) return;
  const int i = idx / B, j = idx % B;
  loss_bb[idx] = l[j] * msum[i];
}

template <int D, int TS, typename LT>
int launch_dc_fast(const float* emb, const void* label, const float* mag, int B, int N, float* scratch,
                   cudaStream_t s) {
  constexpr size_t smem = dc_smem_bytes<D, TS>();
  auto kern = loss_dc_partial_kernel<D, TS, LT>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return ONSSEN_ERR_CUDA;
  }
  dim3 grid((N + DC_PTS_PER_CHUNK - 1) / DC_PTS_PER_CHUNK, B);
  kern<<<grid, DC_THREADS, smem, s>>>(emb, (const LT*)label, mag, N, scratch);
  return ONSSEN_CHECK_LAUNCH();
}

// record layout of the partial / summed Gram records: "fast" = [D*D G][D*2 C][3 Y (y0y0,y0y1,y1y1)][msum],
// generic = [D*D][D*S][S*S][msum]
inline bool dc_fast_layout(const float* emb, int D, int S) {
  if (S != 2 || (reinterpret_cast<uintptr_t>(emb) & 15) != 0) return false;
  return D == 8 || D == 16 || D == 20 || D == 32 || D == 40 || D == 64;
}

inline int dc_tile_size(int D) { return (D == 20 || D == 40) ? 5 : 4; }   // TS of the fast kernels below

template <typename LT>
int dispatch_dc(const float* emb, const void* label, const float* mag, int B, int N, int D, int S,
                float* scratch, cudaStream_t s, int* fast) {
  *fast = 1;
  if (dc_fast_layout(emb, D, S)) {
    switch (D) {
      // 5x5 register tiles (measured on B200 at cfg2: 321 us; 10x10 tiles drop to 1 CTA/SM and take 636 us)
      case 20: return launch_dc_fast<20, 5, LT>(emb, label, mag, B, N, scratch, s);
      case 40: return launch_dc_fast<40, 5, LT>(emb, label, mag, B, N, scratch, s);
      case 8: return launch_dc_fast<8, 4, LT>(emb, label, mag, B, N, scratch, s);
      case 16: return launch_dc_fast<16, 4, LT>(emb, label, mag, B, N, scratch, s);
      case 32: return launch_dc_fast<32, 4, LT>(emb, label, mag, B, N, scratch, s);
      case 64: return launch_dc_fast<64, 4, LT>(emb, label, mag, B, N, scratch, s);
      default: break;
    }
  }
  *fast = 0;
  if (D > 128 || S > 4) return ONSSEN_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(32 * D + 32 * S + 32) * 4;
  dim3 grid((N + DC_PTS_PER_CHUNK - 1) / DC_PTS_PER_CHUNK, B);
  loss_dc_partial_generic<LT><<<grid, DC_THREADS, smem, s>>>(emb, (const LT*)label, mag, N, D, S, scratch);
  return ONSSEN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- loss_dc backward
// d mean(...)/d e_n for utterance j (autograd of loss_dc.py:24-44):  with gl_j = sum_i g_bb[i][j] * msum_i,
//   d e_n = gl_j / msum_j * ( 2 m_n s_n^2 (G e_n) / ||G||  -  2 m_n s_n (C y_n) / ||C|| ),
// G = sum m s^2 e e^T (DxD), C = sum m s e y^T (Dx2) from the forward's summed record.  Thread d of a group
// keeps row d of G in registers and streams points from a smem tile (float4 broadcast loads).
template <int D, typename LT>
__global__ void __launch_bounds__(256)
loss_dc_bwd_kernel(const float* __restrict__ emb, const LT* __restrict__ label, const float* __restrict__ mag,
                   const float* __restrict__ summed, int R, int msum_idx, int sym_ts,
                   const float* __restrict__ g_bb, int B, int N, float* __restrict__ d_emb) {
  constexpr int GSZ = D <= 32 ? 32 : 64;
  constexpr int NG = 256 / GSZ;
  constexpr int P = 64;
  __shared__ __align__(16) float e_s[P][D];
  __shared__ float y_s[P][2];
  __shared__ float m_s[P];
  __shared__ double s_red[2][8];
  __shared__ float s_coef[2];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int g = tid / GSZ;
  const int d = tid % GSZ;
  const float* rec = summed + (long long)b * R;
  // norms of G and C (fp64 block reduction, same for every block of this utterance)
  {
    double q0 = 0.0, q1 = 0.0;
    for (int i = tid; i < D * D; i += 256) {
      const double w = (sym_ts > 0 && (i / D) / sym_ts < (i % D) / sym_ts) ? 2.0 : 1.0;   // see loss_dc_final_kernel
      q0 += w * (double)rec[i] * rec[i];
    }
    for (int i = tid; i < D * 2; i += 256) q1 += (double)rec[D * D + i] * rec[D * D + i];
    q0 = warp_sum(q0); q1 = warp_sum(q1);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = q0; s_red[1][tid >> 5] = q1; }
    __syncthreads();
    if (tid == 0) {
      double t0 = 0.0, t1 = 0.0;
      for (int w = 0; w < 8; ++w) { t0 += s_red[0][w]; t1 += s_red[1][w]; }
      const float msum = rec[msum_idx];
      float gl = 0.f;
      for (int i = 0; i < B; ++i) gl += g_bb[(long long)i * B + b] * summed[(long long)i * R + msum_idx];
      const float ng = sqrtf((float)t0), nc = sqrtf((float)t1);
      s_coef[0] = ng > 0.f ? gl * 2.0f / (msum * ng) : 0.f;
      s_coef[1] = nc > 0.f ? gl * 2.0f / (msum * nc) : 0.f;
    }
    __syncthreads();
  }
  const float cg = s_coef[0], cc = s_coef[1];
  float grow[D];
  float c0 = 0.f, c1 = 0.f;
  if (d < D) {
#pragma unroll
    for (int k = 0; k < D; ++k)   // symmetric record: the entry lives in the tile on/above the diagonal
      grow[k] = (sym_ts > 0 && d / sym_ts > k / sym_ts) ? rec[k * D + d] : rec[d * D + k];
    c0 = rec[D * D + d * 2];
    c1 = rec[D * D + d * 2 + 1];
  }
  const int n_begin = blockIdx.x * DC_PTS_PER_CHUNK;
  const int n_end = min(N, n_begin + DC_PTS_PER_CHUNK);
  const float* eb = emb + (long long)b * N * D;
  const LT* lb = label + (long long)b * N * 2;
  const float* mb = mag + (long long)b * N;
  float* ob = d_emb + (long long)b * N * D;
  for (int n0 = n_begin; n0 < n_end; n0 += P) {
    const int np = min(P, n_end - n0);
    __syncthreads();
    for (int i = tid; i < np; i += 256) {
      y_s[i][0] = lab_to_f(lb[2 * (long long)(n0 + i)]);
      y_s[i][1] = lab_to_f(lb[2 * (long long)(n0 + i) + 1]);
      m_s[i] = mb[n0 + i];
    }
    for (int i = tid; i < np * D / 4; i += 256)
      reinterpret_cast<float4*>(&e_s[0][0])[i] = reinterpret_cast<const float4*>(eb + (long long)n0 * D)[i];
    __syncthreads();
    if (d < D) {
      for (int pnt = g; pnt < np; pnt += NG) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < D; k += 4) {
          const float4 e4 = *reinterpret_cast<const float4*>(&e_s[pnt][k]);
          acc = fmaf(grow[k], e4.x, acc);
          acc = fmaf(grow[k + 1], e4.y, acc);
          acc = fmaf(grow[k + 2], e4.z, acc);
          acc = fmaf(grow[k + 3], e4.w, acc);
        }
        const float y0 = y_s[pnt][0], y1 = y_s[pnt][1];
        const float sn = y0 + y1, m = m_s[pnt];
        const float v = m * sn * (cg * sn * acc - cc * (c0 * y0 + c1 * y1));
        ob[(long long)(n0 + pnt) * D + d] = v;
      }
    }
  }
}

template <typename LT>
int dispatch_dc_bwd(const float* emb, const void* label, const float* mag, const float* summed, int R, int msum_idx,
                    int sym_ts, const float* g_bb, int B, int N, int D, float* d_emb, cudaStream_t s) {
  dim3 grid((N + DC_PTS_PER_CHUNK - 1) / DC_PTS_PER_CHUNK, B);
#define ONSSEN_DC_BWD(DD)                                                                                  \
  case DD:                                                                                                  \
    loss_dc_bwd_kernel<DD, LT><<<grid, 256, 0, s>>>(emb, (const LT*)label, mag, summed, R, msum_idx, sym_ts, g_bb, \
                                                    B, N, d_emb);                                               \
    break;
  switch (D) {
    ONSSEN_DC_BWD(8) ONSSEN_DC_BWD(16) ONSSEN_DC_BWD(20) ONSSEN_DC_BWD(32) ONSSEN_DC_BWD(40) ONSSEN_DC_BWD(64)
    ONSSEN_DC_BWD(4) ONSSEN_DC_BWD(12) ONSSEN_DC_BWD(24)
    default: return ONSSEN_ERR_UNSUPPORTED;
  }
#undef ONSSEN_DC_BWD
  return ONSSEN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------- PIT L1
__global__ void __launch_bounds__(1024)
pit_l1_kernel(const float* __restrict__ mask_a, const float* __restrict__ mask_b, long long mstride,
              const float* __restrict__ mix, const float* __restrict__ s1, const float* __restrict__ s2,
              const float* __restrict__ c1, const float* __restrict__ c2, int N, float* __restrict__ out,
              int32_t* __restrict__ perm) {
  const int b = blockIdx.x;
  const long long off = (long long)b * N;
  double q[4] = {0.0, 0.0, 0.0, 0.0};
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float m = mix[off + n];
    float t1 = s1[off + n], t2 = s2[off + n];
    if (c1 != nullptr) {
      t1 = fminf(m, fmaxf(t1 * c1[off + n], 0.f));
      t2 = fminf(m, fmaxf(t2 * c2[off + n], 0.f));
    }
    const float ea = mask_a[(off + n) * mstride] * m;
    const float eb = mask_b[(off + n) * mstride] * m;
    q[0] += fabsf(ea - t1);
    q[1] += fabsf(eb - t2);
    q[2] += fabsf(eb - t1);
    q[3] += fabsf(ea - t2);
  }
  __shared__ double sp[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < 4; ++k) {
    const double r = warp_sum(q[k]);
    if (lane == 0) sp[k][warp] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < 4; ++k)
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t[k] += sp[k][w];
    const float l1 = (float)t[0] + (float)t[1];
    const float l2 = (float)t[2] + (float)t[3];
    out[b] = fminf(l1, l2);
    if (perm) perm[b] = (l1 < l2) ? 0 : 1;   // loss_phase.py:21 uses strict <
  }
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" int onssen_loss_dc_num_chunks(int N) { return (N + DC_PTS_PER_CHUNK - 1) / DC_PTS_PER_CHUNK; }

extern "C" int onssen_loss_dc_fwd(const float* emb, const void* label, int label_dtype, const float* mag, int B,
                                  int N, int D, int S, float* loss_bb, float* l, float* mag_sum, float* scratch,
                                  void* stream) {
  if (!emb || !label || !mag || !scratch || B <= 0 || N <= 0 || D <= 0 || S <= 0) return ONSSEN_ERR_ARG;
  if (!l || !mag_sum) return ONSSEN_ERR_ARG;  // needed as intermediates for the (B,B) product
  cudaStream_t s = (cudaStream_t)stream;
  int fast = 0, rc;
  switch (label_dtype) {
    case ONSSEN_DT_F32: rc = dispatch_dc<float>(emb, label, mag, B, N, D, S, scratch, s, &fast); break;
    case ONSSEN_DT_F64: rc = dispatch_dc<double>(emb, label, mag, B, N, D, S, scratch, s, &fast); break;
    case ONSSEN_DT_U8: rc = dispatch_dc<uint8_t>(emb, label, mag, B, N, D, S, scratch, s, &fast); break;
    default: return ONSSEN_ERR_ARG;
  }
  if (rc != ONSSEN_OK) return rc;
  const int nchunk = onssen_loss_dc_num_chunks(N);
  const int R = fast ? (D * D + D * 2 + 4) : (D * D + D * S + S * S + 1);
  float* summed = scratch + (long long)B * nchunk * (D * D + D * S + S * S + 4);   // behind the partial records
  loss_dc_reduce_kernel<<<dim3((R + 255) / 256, B), 256, 0, s>>>(scratch, nchunk, R, summed);
  loss_dc_final_kernel<<<B, 256, 0, s>>>(summed, 1, D, S, fast, fast ? dc_tile_size(D) : 0, l, mag_sum);
  if (loss_bb) loss_dc_outer_kernel<<<(B * B + 255) / 256, 256, 0, s>>>(l, mag_sum, B, loss_bb);
  return ONSSEN_CHECK_LAUNCH();
}


extern "C" int onssen_loss_dc_bwd(const float* emb, const void* label, int label_dtype, const float* mag,
                                  const float* summed_record, const float* g_bb, int B, int N, int D, int S,
                                  float* d_emb, void* stream) {
  if (!emb || !label || !mag || !summed_record || !g_bb || !d_emb || B <= 0 || N <= 0) return ONSSEN_ERR_ARG;
  // the fast forward layout (S == 2, D % 4 == 0) is the one the training path supports
  if (S != 2 || (D & 3) || (reinterpret_cast<uintptr_t>(emb) & 15)) return ONSSEN_ERR_UNSUPPORTED;
  const bool fast = dc_fast_layout(emb, D, S);
  const int R = fast ? D * D + D * 2 + 4 : D * D + D * S + S * S + 1;
  const int mi = R - 1;
  const int st = fast ? dc_tile_size(D) : 0;
  cudaStream_t s = (cudaStream_t)stream;
  switch (label_dtype) {
    case ONSSEN_DT_F32: return dispatch_dc_bwd<float>(emb, label, mag, summed_record, R, mi, st, g_bb, B, N, D, d_emb, s);
    case ONSSEN_DT_F64: return dispatch_dc_bwd<double>(emb, label, mag, summed_record, R, mi, st, g_bb, B, N, D, d_emb, s);
    case ONSSEN_DT_U8: return dispatch_dc_bwd<uint8_t>(emb, label, mag, summed_record, R, mi, st, g_bb, B, N, D, d_emb, s);
    default: return ONSSEN_ERR_ARG;
  }
}

extern "C" int onssen_loss_pit_l1_fwd(const float* mask_a, const float* mask_b, long long mask_stride,
                                      const float* mag_mix, const float* mag_s1, const float* mag_s2,
                                      const float* cos_s1, const float* cos_s2, int B, int N, float* out,
                                      int32_t* perm, void* stream) {
  if (!mask_a || !mask_b || !mag_mix || !mag_s1 || !mag_s2 || !out || B <= 0 || N <= 0 || mask_stride <= 0)
    return ONSSEN_ERR_ARG;
  if ((cos_s1 == nullptr) != (cos_s2 == nullptr)) return ONSSEN_ERR_ARG;
  pit_l1_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(mask_a, mask_b, mask_stride, mag_mix, mag_s1, mag_s2,
                                                      cos_s1, cos_s2, N, out, perm);
  return ONSSEN_CHECK_LAUNCH();
}
