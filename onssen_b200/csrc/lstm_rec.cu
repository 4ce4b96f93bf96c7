// Persistent bidirectional LSTM recurrence for one layer (both directions) on sm_100a.
//
// Replaces the time loop hidden inside torch.nn.LSTM / cuDNN at onssen/nn/deep_clustering.py:34-35,
// chimera.py:35-36, enhancement.py:43-44, phase_network.py:50,57:
//     g_t = (W_ih x_t + b_ih + b_hh) + W_hh h_{t-1};   i,f,o = sigmoid, g = tanh   (gate order i,f,g,o)
//     c_t = f*c_{t-1} + i*g;   h_t = o*tanh(c_t);   h_0 = c_0 = 0
// The bracketed term for all t is precomputed by the projection GEMM (gemm_tc05.cu) into `gates`.
//
// Decomposition: grid = 2 directions x S batch slices x nrb row blocks, all CTAs
// co-resident (cooperative launch, 1 CTA / SM).  CTA (dir, slice, rb) owns 32 hidden units = 128 gate rows
// (row r = 4*unit + gate) and keeps its 128 x Hp fp16 slice of W_hh resident in shared memory for the whole
// sequence (weights are read from HBM exactly once per layer).  Per step:
//   control thread : wait until all nrb CTAs of its (dir,slice) group published h_{t-1}  (global counter,
//                    acquire) -> cp.async.bulk h_{t-1} [NBP x Hp fp16, pre-laid-out as UMMA B operand]
//                    -> Hp/16 x tcgen05.mma (M=128 gate rows, N=NBP batch columns, K=16) -> commit
//   4 gate warps   : tcgen05.ld their 32 TMEM lanes (one gate row per thread, NB batch columns),
//                    add the prefetched input pre-activation, sigmoid/tanh (one transcendental chain per
//                    thread: the 4 gates of a unit sit in 4 adjacent lanes), exchange the activated gates
//                    through a warp-private smem tile, update c (registers) and h, publish h_t (fp16) to the
//                    group's exchange buffer and to the layer output, then release-increment the counter.
// The recurrent contraction runs on tensor cores (fp16 operands, fp32 accumulate); everything else is
// latency-bound control: the design minimises the serial chain per step, not bytes.
#include "tc05.cuh"
#include "common.cuh"

namespace onssen {
namespace {

using namespace tc05;

constexpr int REC_THREADS = 160;  // 4 gate warps + 1 control warp

struct RecParams {
  const float* gates;
  const __half* whh;
  __half* y_h;
  float* y_f;
  __half* hbuf;
  unsigned int* flags;
  int B, T, H, Hp, nrb, S, Bs;
  float dropout_p;
  unsigned long long seed, offset;
};

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
  // splitmix64 finaliser -> 24-bit uniform in [0,1)
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(unsigned int)(z >> 40) * (1.0f / 16777216.0f);
}

template <int NB, bool TC>
__global__ void __launch_bounds__(REC_THREADS, 1) blstm_rec_kernel(const RecParams p) {
  constexpr int NBP = NB <= 16 ? 16 : 32;  // MMA N / rows of the h operand tile
  constexpr int XP = NB + 1;               // exchange tile pitch (floats)
  extern __shared__ __align__(128) uint8_t smem[];
  const int Hp = p.Hp;
  const uint32_t w_bytes = (uint32_t)Hp * 128u * 2u;
  const uint32_t h_bytes = (uint32_t)Hp * NBP * 2u;
  uint8_t* w_s = smem;
  uint8_t* h_s = smem + w_bytes;
  float* xch = reinterpret_cast<float*>(h_s + h_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 128 * XP + ((128 * XP) & 1));
  uint64_t* wbar = bars;
  uint64_t* hbar = bars + 1;
  uint64_t* mbar = bars + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int rb = blockIdx.x % p.nrb;
  const int sl = (blockIdx.x / p.nrb) % p.S;
  const int dir = blockIdx.x / (p.nrb * p.S);
  const int b0 = sl * p.Bs;
  const int nb_valid = min(p.Bs, p.B - b0);
  const int T = p.T;
  unsigned int* flag = p.flags + dir * p.S + sl;
  const size_t hbuf_group = (size_t)Hp * NBP;  // halves per (parity,dir,slice)
  __half* hbuf0 = p.hbuf + ((size_t)(0 * 2 + dir) * p.S + sl) * hbuf_group;
  __half* hbuf1 = p.hbuf + ((size_t)(1 * 2 + dir) * p.S + sl) * hbuf_group;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(wbar, 1);
      mbar_init(hbar, 1);
      mbar_init(mbar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    if (TC) tmem_alloc(tmem_ptr, 32);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base = 0;
  if (TC) tmem_base = *tmem_ptr;

  if (warp == 4) {
    // ===================== control warp =====================
    if (lane == 0) {
      // resident recurrent weights: one contiguous slab per (dir, rb)
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.whh) + ((size_t)dir * p.nrb + rb) * w_bytes;
      mbar_arrive_expect_tx(wbar, w_bytes);
      constexpr uint32_t CH = 32768;
      for (uint32_t off = 0; off < w_bytes; off += CH) {
        const uint32_t n = (w_bytes - off) < CH ? (w_bytes - off) : CH;
        bulk_load(w_s + off, wsrc + off, n, wbar);
      }
      if (TC) mbar_wait(wbar, 0);
    }
    __syncwarp();
    const uint32_t idesc = make_idesc_f16(128, NBP);
    const uint32_t w_addr = smem_u32(w_s);
    const uint32_t h_addr = smem_u32(h_s);
    for (int s = 1; s < T; ++s) {
      if (lane == 0) {
        const unsigned int target = (unsigned int)p.nrb * (unsigned int)s;
        while (ld_acquire_u32(flag) < target) {
        }
        fence_proxy_async();
        const __half* hsrc = ((s - 1) & 1) ? hbuf1 : hbuf0;
        mbar_arrive_expect_tx(hbar, h_bytes);
        bulk_load(h_s, hsrc, h_bytes, hbar);
        if (TC) {
          mbar_wait(hbar, (s - 1) & 1);
          tc_fence_after_sync();
          const int ksteps = Hp / 16;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = make_smem_desc(w_addr + ks * 4096, 2048, 128, 0);
            const uint64_t db = make_smem_desc(h_addr + ks * (2 * NBP * 16), NBP * 16, 128, 0);
            umma_f16(tmem_base, da, db, idesc, ks != 0);
          }
          umma_commit(mbar);
        }
      }
      __syncwarp();
      // wait until the gate warps have drained TMEM / finished reading h_s for this step
      named_bar_sync(2, REC_THREADS);
      tc_fence_after_sync();
    }
  } else {
    // ===================== gate warps =====================
    const int r = tid;            // gate row inside the row block: r = 4*ul + gate
    const int gate = r & 3;
    const int ul = r >> 2;        // unit inside the row block (0..31)
    const int u = rb * 32 + ul;   // padded hidden unit index
    const float ak = (gate == 2) ? 2.0f : 1.0f;   // act(x) = ak*sigmoid(ak*x) + ab  (tanh for gate g)
    const float ab = (gate == 2) ? -1.0f : 0.0f;
    const long long ldg = 2LL * 4 * Hp;
    const long long ldy = 2LL * Hp;
    const float* gcol = p.gates + (long long)dir * 4 * Hp + rb * 128 + r;
    const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;

    float c_state[NB / 4];
#pragma unroll
    for (int i = 0; i < NB / 4; ++i) c_state[i] = 0.f;
    float gpre[NB];
    {
      const int t0 = dir ? T - 1 : 0;
#pragma unroll
      for (int j = 0; j < NB; ++j)
        gpre[j] = (j < nb_valid) ? __ldcs(gcol + ((long long)t0 * p.B + b0 + j) * ldg) : 0.f;
    }
    if (!TC) mbar_wait(wbar, 0);

    for (int s = 0; s < T; ++s) {
      const int t = dir ? T - 1 - s : s;
      float acc[NB];
      if (s == 0) {
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = 0.f;
      } else if (TC) {
        mbar_wait(mbar, (s - 1) & 1);
        tc_fence_after_sync();
        uint32_t v[NBP];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        if (NBP == 16) tmem_ld16(taddr, v); else tmem_ld32(taddr, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = __uint_as_float(v[j]);
        tc_fence_before_sync();
        asm volatile("bar.arrive 2, %0;" ::"r"(REC_THREADS) : "memory");
      } else {
        mbar_wait(hbar, (s - 1) & 1);
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = 0.f;
        const uint4* wv = reinterpret_cast<const uint4*>(w_s);
        const uint4* hv = reinterpret_cast<const uint4*>(h_s);
        for (int kc = 0; kc < Hp / 8; ++kc) {
          const uint4 w8 = wv[kc * 128 + r];
          const __half2* wh = reinterpret_cast<const __half2*>(&w8);
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const uint4 h8 = hv[kc * NBP + j];
            const __half2* hh = reinterpret_cast<const __half2*>(&h8);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = __half22float2(wh[e]);
              const float2 b = __half22float2(hh[e]);
              acc[j] = fmaf(a.x, b.x, acc[j]);
              acc[j] = fmaf(a.y, b.y, acc[j]);
            }
          }
        }
        asm volatile("bar.arrive 2, %0;" ::"r"(REC_THREADS) : "memory");
      }

      // activation of this thread's gate for all NB batch columns
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float pre = acc[j] + gpre[j];
        xch[r * XP + j] = fmaf(ak, sigmoid_f(ak * pre), ab);
      }
      // prefetch next step's input pre-activations (hidden behind this step's tail + next step's wait)
      if (s + 1 < T) {
        const int tn = dir ? t - 1 : t + 1;
#pragma unroll
        for (int j = 0; j < NB; ++j)
          gpre[j] = (j < nb_valid) ? __ldcs(gcol + ((long long)tn * p.B + b0 + j) * ldg) : 0.f;
      }
      __syncwarp();

      __half* hdst = (s & 1) ? hbuf1 : hbuf0;
#pragma unroll
      for (int ci = 0; ci < NB / 4; ++ci) {
        const int j = 4 * ci + gate;  // this lane finishes batch column j of unit ul
        const float* xr = xch + (4 * ul) * XP + j;
        const float gi = xr[0];
        const float gf = xr[XP];
        const float gg = xr[2 * XP];
        const float go = xr[3 * XP];
        const float c = fmaf(gf, c_state[ci], gi * gg);
        c_state[ci] = c;
        const float h = go * fmaf(2.0f, sigmoid_f(2.0f * c), -1.0f);
        // next step's B operand tile: [k/8][n][k%8]
        hdst[((size_t)(u >> 3) * NBP + j) * 8 + (u & 7)] = __float2half_rn(h);
        if (j < nb_valid) {
          const long long m = (long long)t * p.B + b0 + j;
          float hv = h;
          if (p.dropout_p > 0.f) {
            const float rnd = hash_uniform(p.seed, p.offset + (unsigned long long)(m * ldy + dir * Hp + u));
            hv = rnd < p.dropout_p ? 0.f : h * keep_scale;
          }
          if (p.y_h) p.y_h[m * ldy + dir * Hp + u] = __float2half_rn(hv);
          if (p.y_f) p.y_f[m * ldy + dir * Hp + u] = hv;
        }
      }
      // publish: all 128 gate threads' stores, then one release increment
      named_bar_sync(1, 128);
      if (tid == 0) {
        __threadfence();
        atomicAdd(flag, 1u);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (TC && warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 32);
  }
}

template <int NB>
size_t rec_smem_bytes(int Hp) {
  constexpr int NBP = NB <= 16 ? 16 : 32;
  return (size_t)Hp * 128 * 2 + (size_t)Hp * NBP * 2 + (size_t)(128 * (NB + 1) + 1) * 4 + 64;
}

template <int NB, bool TC>
int launch_rec(RecParams& p, int grid, cudaStream_t stream) {
  const size_t smem = rec_smem_bytes<NB>(p.Hp);
  auto kern = blstm_rec_kernel<NB, TC>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return ONSSEN_ERR_CUDA;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem) != cudaSuccess)
    return ONSSEN_ERR_CUDA;
  if (per_sm * num_sms() < grid) return ONSSEN_ERR_RESIDENCY;
  void* args[] = {(void*)&p};
  if (cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(REC_THREADS), args, smem, stream) !=
      cudaSuccess)
    return ONSSEN_ERR_CUDA;
  return ONSSEN_OK;
}

template <bool TC>
int dispatch_nb(int nb, RecParams& p, int grid, cudaStream_t stream) {
  switch (nb) {
    case 4: return launch_rec<4, TC>(p, grid, stream);
    case 8: return launch_rec<8, TC>(p, grid, stream);
    case 12: return launch_rec<12, TC>(p, grid, stream);
    case 16: return launch_rec<16, TC>(p, grid, stream);
    case 24: return launch_rec<24, TC>(p, grid, stream);
    case 32: return launch_rec<32, TC>(p, grid, stream);
    default: return ONSSEN_ERR_UNSUPPORTED;
  }
}

// slice plan shared by workspace sizing and launch
struct SlicePlan {
  int S, Bs, NB, NBP;
  bool ok;
};
SlicePlan plan_slices(int B, int H) {
  SlicePlan sp{};
  const int nrb = hp_of(H) / 32;
  int smax = num_sms() / (2 * nrb);
  if (smax < 1) { sp.ok = false; return sp; }
  if (smax > B) smax = B;
  int bs = (B + smax - 1) / smax;
  static const int opts[] = {4, 8, 12, 16, 24, 32};
  int nb = -1;
  for (int o : opts) if (o >= bs) { nb = o; break; }
  if (nb < 0) { sp.ok = false; return sp; }
  sp.Bs = bs;
  sp.S = (B + bs - 1) / bs;
  sp.NB = nb;
  sp.NBP = nb <= 16 ? 16 : 32;
  sp.ok = true;
  return sp;
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" size_t onssen_blstm_rec_workspace_bytes(int B, int H) {
  if (B <= 0 || H <= 0) return 0;
  const int Hp = hp_of(H);
  const int nrb = Hp / 32;
  int smax = num_sms() / (2 * nrb);
  if (smax < 1) smax = 1;
  // upper bound independent of the exact plan: NBP = 32
  return 256 + (size_t)2 * 2 * smax * Hp * 32 * 2;
}

extern "C" int onssen_blstm_rec_fwd(const float* gates, const void* whh_p, int B, int T, int H, void* y_h,
                                    float* y_f, float dropout_p, unsigned long long seed,
                                    unsigned long long offset, void* workspace, size_t workspace_bytes,
                                    int use_tensor_cores, void* stream) {
  if (!gates || !whh_p || !workspace || B <= 0 || T <= 0 || H <= 0) return ONSSEN_ERR_ARG;
  if (!y_h && !y_f) return ONSSEN_ERR_ARG;
  if (dropout_p < 0.f || dropout_p >= 1.f) return ONSSEN_ERR_ARG;
  const SlicePlan sp = plan_slices(B, H);
  if (!sp.ok) return ONSSEN_ERR_UNSUPPORTED;
  RecParams p;
  p.gates = gates;
  p.whh = (const __half*)whh_p;
  p.y_h = (__half*)y_h;
  p.y_f = y_f;
  p.B = B; p.T = T; p.H = H; p.Hp = hp_of(H); p.nrb = p.Hp / 32; p.S = sp.S; p.Bs = sp.Bs;
  p.dropout_p = dropout_p; p.seed = seed; p.offset = offset;
  const size_t hbuf_bytes = (size_t)2 * 2 * sp.S * p.Hp * sp.NBP * 2;
  if (workspace_bytes < 256 + hbuf_bytes) return ONSSEN_ERR_ARG;
  p.flags = (unsigned int*)workspace;
  p.hbuf = (__half*)((uint8_t*)workspace + 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(workspace, 0, 256 + hbuf_bytes, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  const int grid = 2 * sp.S * p.nrb;
  return use_tensor_cores ? dispatch_nb<true>(sp.NB, p, grid, s) : dispatch_nb<false>(sp.NB, p, grid, s);
}
