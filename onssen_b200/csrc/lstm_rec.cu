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
// (row r = 4*unit + gate).  Its 128 x Hp fp16 slice of W_hh is loaded ONCE into TENSOR MEMORY (lane = gate
// row, two fp16 per 32-bit column: 304 of the 512 TMEM columns at H=600) and stays there for the whole
// sequence as the A operand of tcgen05.mma (TS form): recurrent weights cross HBM once per layer and are
// never re-streamed through shared memory.  Per step:
//   MMA warp       : (converged, elected lane issues) waits on a named barrier until the gate warps have written the
//                    h_{t-1} tile [NBP x Hp fp16, UMMA B-operand layout] to smem -> Hp/16 x tcgen05.mma (A: TMEM,
//                    B: smem, D: 2 TMEM accumulators; M=128, N=NBP, K=16) -> tcgen05.commit to an mbarrier
//   gate warps     : (16 / 12 / 16 warps for 16 / 24 / 32-column slices) tcgen05.ld their TMEM lanes (thread = one gate
//                    row x 4 or 8 batch columns), add the input pre-activation prefetched two steps ahead, activate
//                    (inference: one tanh.approx per value; training forward: ex2 + rcp), exchange activated gates
//                    through a warp-private smem tile (the 4 gates of a unit are 4 adjacent lanes), update c
//                    (registers) and h, PUBLISH h_t: 16-byte relaxed stores into the group's L2-resident exchange
//                    tile with bit 14 of every fp16 (always 0 for |h| <= 1) carrying the step parity -- the data
//                    validate themselves, no counter, no fence; write the layer output and issue the prefetch;
//                    GATHER: spin on 16-byte relaxed loads of all nrb producers' chunks until their flag bits are
//                    this step's, strip the flag, write the next step's operand tile, fence.proxy.async, bar.arrive.
// Latency-bound by the per-step serial chain (MMA issue -> TMEM load + gate math -> publish -> gather), not by bytes;
// DESIGN.md section 4.1 and profiles/r02_exchange_experiments.md hold the measured anatomy and the rejected variants.
#include "tc05.cuh"
#include "common.cuh"

namespace onssen {
namespace {

using namespace tc05;

// Gate warps of the tensor-core path: thread = one gate row x 8 batch columns, so a slice of NB = 16 / 24 / 32 columns
// takes 8 / 12 / 16 warps (4 warps cover the 128 TMEM lanes).  (Round 1 ran 32-column slices on 8 warps, 16 columns per
// thread: 43 KB of SASS, more than the 32 KB instruction cache the step body streams through once per step.)
// The SIMT validation path keeps 8 warps.
#ifndef ONSSEN_REC_PREFETCH_AFTER
#define ONSSEN_REC_PREFETCH_AFTER 0   // A/B knob: issue the step-(s+2) gate prefetch after the gather instead of before it
#endif
#ifndef ONSSEN_REC_GW16
#define ONSSEN_REC_GW16 16    // gate warps of the 16-column instantiation: 16 = 4 columns per thread (measured at cfg2:
#endif                        // 8 warps 2.37 us/step, 16 warps 2.27 -- four warps per scheduler hide the dependent-issue
                              // latency of the gate math -- and 2.16 with the one-MUFU activations below)
__host__ __device__ constexpr int rec_gate_warps(int nb, bool tc) { return tc ? (nb == 16 ? ONSSEN_REC_GW16 : nb / 2) : 8; }
__host__ __device__ constexpr int rec_threads(int nb, bool tc) { return (rec_gate_warps(nb, tc) + 1) * 32; }  // + MMA warp
constexpr int TMEM_A_COL = 128;                         // first TMEM column of the resident W_hh slice
constexpr int NACC = 2;   // independent accumulators (columns a*NBP): back-to-back MMAs into ONE accumulator
                          // serialise on the ~70-cycle accumulate latency when N is this small
constexpr int TRACE_S0 = 100;
long long* g_trace_ptr = nullptr;
int g_poll_delay = 0;   // SM cycles to wait between publishing h_t and the first gather round (tuning knob)

struct RecParams {
  const float* gates;
  const __half* whh;
  __half* y_h;
  float* y_f;
  __half* hbuf;
  // column-chunk launches only.  (These two replace an unused 8-byte field of the round-1 struct: growing the
  // parameter struct changes ptxas' allocation and the NB=32 instantiation starts to spill.)
  int Bp;                  // row pitch of the time-major buffers in utterances (> B for a column chunk)
  unsigned int hash_off;   // element offset of the chunk's first column (dropout hash uses absolute indices)
  int B, T, H, Hp, nrb, S, Bs;
  float dropout_p;
  unsigned int seed_lo, seed_hi;
  long long* trace;  // debug: clock64 stamps of CTA 0, steps [TRACE_S0, TRACE_S0+4)
  // training: state saved for BPTT (all optional)
  int poll_delay;
  float* act_out;    // == gates: activated i,f,g,o written in place over the consumed pre-activations
  float* c_out;      // [T*B][2*Hp] cell state
  __half* h_raw;     // [T*B][2*Hp] h before dropout (operand of the W_hh weight gradient)
  const int* col_len; // optional (inference on padded batches): frames of every utterance; state and output of column b are
                      // held at zero for t >= col_len[b], so the reverse direction starts at the utterance's own last frame
};

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// One MUFU op (tanh.approx.f32, max rel. error 2^-11 -- the size of the fp16 rounding h gets anyway) instead of
// ex2 + rcp per activation: template parameter FAST.  With 8 gate warps it made no difference (round 1: 2.40 us/step
// either way); with 16 the MUFU pipe is the limiter of the gate phase (512 threads x 4 columns x 2 ops) and it is worth
// 5 % of the step (2.27 -> 2.16 us).  Used by the inference forward only (nothing saved for BPTT): the training forward
// keeps the ex2 + rcp activations, whose saved values the hand-written backward differentiates -- the ill-conditioned
// phase_net gradient check (tests/test_reference_gpu.py, cfg4) moved from 2.7e-2 to 9e-2 with approximate activations.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 32-bit counter hash -> uniform [0,1): inter-layer dropout mask (statistical parity only, like cuDNN's)
__device__ __forceinline__ float hash_uniform32(unsigned int seed_lo, unsigned int seed_hi, unsigned int idx) {
  unsigned int x = idx ^ seed_lo;
  x *= 0x9E3779B1u; x ^= x >> 15;
  x *= 0x85EBCA77u; x ^= x >> 13;
  x += seed_hi;
  x *= 0xC2B2AE3Du; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

// group trace: every CTA of (dir 0, slice 0) stamps the global timer (ns) at slot for step TRACE_S0+2 into
// trace[64 + rb*8 + slot]  (needs a 64 + nrb*8 int64 buffer)
__device__ __forceinline__ long long gtimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define REC_GTRACE(slot)                                                                        \
  do {                                                                                          \
    if (p.trace != nullptr && dir == 0 && sl == 0 && s == TRACE_S0 + 2)                         \
      p.trace[64 + rb * 8 + (slot)] = gtimer_ns();                                              \
  } while (0)
#define REC_TRACE(slot)                                                             \
  do {                                                                              \
    if (p.trace != nullptr && blockIdx.x == 0 && s >= TRACE_S0 && s < TRACE_S0 + 4) \
      p.trace[(s - TRACE_S0) * 16 + (slot)] = clock64();                            \
  } while (0)

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* v) {
  if constexpr (N == 4) tmem_ld4(taddr, v);
  if constexpr (N == 8) tmem_ld8(taddr, v);
  if constexpr (N == 12) { tmem_ld8(taddr, v); tmem_ld4(taddr + 8, v + 8); }
  if constexpr (N == 16) tmem_ld16(taddr, v);
}

// Self-validating exchange: |h_t| <= 1, so bit 14 (top exponent bit) of every fp16 h is always 0 and is used
// as a per-element step-parity flag.  A buffer (selected by s&1) is rewritten every 2 steps with the flag bit
// toggled, so a reader knows an element is fresh from the element itself: no fences, no counters, no extra
// bytes (the NCCL "LL" idea at 1 bit per element).  Stores are 4-byte relaxed (two units of one column).
__device__ __forceinline__ void st_relaxed_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// CHUNK: the launch covers a column chunk of a larger batch (row pitch p.Bp > p.B, absolute dropout indices); the
// whole-batch instantiation keeps the round-1 register allocation (two more live parameters spill at NB = 32).
template <int NB, bool TC, bool CHUNK, bool FAST>
__global__ void __launch_bounds__(rec_threads(NB, TC), 1) blstm_rec_kernel(const RecParams p) {
  constexpr int REC_GATE_WARPS = rec_gate_warps(NB, TC);
  constexpr int REC_THREADS = rec_threads(NB, TC);
  constexpr int NBP = NB <= 16 ? 16 : 32;  // MMA N / rows of the h operand tile
  constexpr int NBH = NB / (REC_GATE_WARPS / 4);   // batch columns per gate warp (4 warps cover the 128 TMEM lanes)
  constexpr int XP = NB + 1;               // exchange tile pitch (floats)
  constexpr int GT = REC_GATE_WARPS * 32;  // gate threads
  extern __shared__ __align__(128) uint8_t smem[];
  const int Hp = p.Hp;
  const uint32_t w_bytes = TC ? 0u : (uint32_t)Hp * 128u * 2u;   // SIMT validation path keeps W in smem
  const uint32_t h_bytes = (uint32_t)Hp * NBP * 2u;
  uint8_t* w_s = smem;
  uint8_t* h_s = smem + w_bytes;
  float* xch = reinterpret_cast<float*>(h_s + h_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 128 * XP + ((128 * XP) & 1));
  uint64_t* wbar = bars;
  uint64_t* mbar = bars + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int rb = blockIdx.x % p.nrb;
  const int sl = (blockIdx.x / p.nrb) % p.S;
  const int dir = blockIdx.x / (p.nrb * p.S);
  const int b0 = sl * p.Bs;
  const int nb_valid = min(p.Bs, p.B - b0);
  const int T = p.T;
  // exchange buffers: per (parity, dir, slice) one operand tile [k/8][n][k%8] fp16 with flag bits
  const size_t ll_group_bytes = (size_t)Hp * NBP * 2;
  uint8_t* ll0 = reinterpret_cast<uint8_t*>(p.hbuf) + ((size_t)(0 * 2 + dir) * p.S + sl) * ll_group_bytes;
  uint8_t* ll1 = reinterpret_cast<uint8_t*>(p.hbuf) + ((size_t)(1 * 2 + dir) * p.S + sl) * ll_group_bytes;
  const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.whh) + ((size_t)dir * p.nrb + rb) * ((size_t)Hp * 256);

  if (warp == REC_GATE_WARPS) {
    if (lane == 0) {
      mbar_init(wbar, 1);
      mbar_init(mbar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    if (TC) tmem_alloc(tmem_ptr, 512);
  }
  // zero the h operand tile once (rows n >= NB are never written afterwards and must read as zero)
  for (uint32_t i = tid; i < h_bytes / 16; i += REC_THREADS) reinterpret_cast<uint4*>(h_s)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base = 0;
  if (TC) tmem_base = *tmem_ptr;

  // ---- resident weights: TMEM (product path) or smem (SIMT validation path) ----
  if (TC) {
    if (warp < 4) {
      // thread = gate row r = 32*warp + lane: its Hp fp16 = Hp/2 words go to lane r, columns TMEM_A_COL..
      const uint4* src = reinterpret_cast<const uint4*>(wsrc + (size_t)(warp * 32 + lane) * Hp * 2);
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + TMEM_A_COL;
      const int words = Hp / 2;
      int c = 0;
      for (; c + 32 <= words; c += 32) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 q4 = __ldg(src + c / 4 + i);
          v[4 * i] = q4.x; v[4 * i + 1] = q4.y; v[4 * i + 2] = q4.z; v[4 * i + 3] = q4.w;
        }
        tmem_st32(trow + c, v);
      }
      if (c < words) {   // Hp multiple of 32 -> remainder is exactly 16 words
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 q4 = __ldg(src + c / 4 + i);
          v[4 * i] = q4.x; v[4 * i + 1] = q4.y; v[4 * i + 2] = q4.z; v[4 * i + 3] = q4.w;
        }
        tmem_st16(trow + c, v);
      }
      tmem_wait_st();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  } else if (warp == REC_GATE_WARPS && lane == 0) {
    mbar_arrive_expect_tx(wbar, w_bytes);
    constexpr uint32_t CH = 32768;
    for (uint32_t off = 0; off < w_bytes; off += CH) {
      const uint32_t n = (w_bytes - off) < CH ? (w_bytes - off) : CH;
      bulk_load(w_s + off, wsrc + off, n, wbar);
    }
  }

  if (warp == REC_GATE_WARPS) {
    // ===================== MMA warp =====================
    // The whole warp runs the loop converged so descriptors / TMEM addresses live in uniform registers;
    // only the tcgen05 instructions are predicated on the elected lane (a single divergent lane pays ~200
    // cycles of R2UR traffic per MMA, see scripts/microbench/mma_cost.cu).
    if (TC) {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_f16(128, NBP);
      const uint32_t h_addr = smem_u32(h_s);
      const int ksteps = Hp / 16;
      constexpr uint64_t DB_STEP = (2 * NBP * 16) >> 4;   // start-address field counts 16-byte units
      for (int s = 1; s < T; ++s) {
        // all gate threads have written + fenced their part of h_{s-1} and drained the accumulators
        named_bar_sync(3, REC_THREADS);
        tc_fence_after_sync();
        if (lane == 0) REC_TRACE(3);
        // blocks of 8 k-steps whose operands are computed independently from the block base: incremental
        // updates would chain every MMA behind ~10-cycle uniform-datapath adds
        uint64_t db0 = make_smem_desc(h_addr, NBP * 16, 128, 0);
        uint32_t ta0 = tmem_base + TMEM_A_COL;
        constexpr int UB = 19;   // k-steps per unrolled issue block (H=600: 38 = 2 blocks)
        for (int ks0 = 0; ks0 < ksteps; ks0 += UB) {
#pragma unroll
          for (int j = 0; j < UB; ++j) {
            if (ks0 + j < ksteps && leader)
              umma_f16_ts(tmem_base + (j % NACC) * NBP, ta0 + 8 * j, db0 + (uint64_t)j * DB_STEP, idesc,
                          (ks0 + j) >= NACC);
          }
          ta0 += 8 * UB;       // UB k-steps x (K=16 fp16 = 8 TMEM columns)
          db0 += UB * DB_STEP;
        }
        if (leader) umma_commit(mbar);
        __syncwarp();
        if (lane == 0) REC_TRACE(4);
      }
    }
  } else {
    // ===================== gate warps =====================
    const int q = warp & 3;         // TMEM lane quarter
    const int half = warp >> 2;     // which group of NBH batch columns this warp owns
    const int r = q * 32 + lane;    // gate row inside the row block: r = 4*ul + gate
    const int gate = r & 3;
    const int ul = r >> 2;          // unit inside the row block (0..31)
    const int u = rb * 32 + ul;     // padded hidden unit index
    const int jbase = half * NBH;
    const float ak = (gate == 2) ? 2.0f : 1.0f;   // act(x) = ak*sigmoid(ak*x) + ab  (tanh for gate g)
    const float ab = (gate == 2) ? -1.0f : 0.0f;
    (void)ak; (void)ab;
    const long long ldg = 2LL * 4 * Hp;
    const int ldy = 2 * Hp;
    const float* gcol = p.gates + (long long)dir * 4 * Hp + rb * 128 + r;
    const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
    // gather: 16-byte chunks (kc, n) = 8 consecutive units of one column; chunk c = kc*NBP + n at byte 16*c
    // in both the global exchange tile and the smem operand tile.  Thread handles c = tid + GT*i.
    const int nchunks = Hp * NBP / 8;
    constexpr int MAXCH = (768 * NBP / 8 + GT - 1) / GT;   // Hp <= 768
    const int my_chunks = (nchunks - tid + GT - 1) / GT;   // <= MAXCH (checked on the host)
    unsigned int want_mask = 0;                            // chunks of real batch columns only (pads stay 0)
    for (int i = 0; i < my_chunks; ++i)
      if (((tid + GT * i) % NBP) < nb_valid) want_mask |= 1u << i;

    float c_state[NBH / 4];
    int len_ci[NBH / 4];          // frames of the utterance each of this lane's (unit, column) items belongs to
#pragma unroll
    for (int i = 0; i < NBH / 4; ++i) {
      c_state[i] = 0.f;
      len_ci[i] = p.col_len != nullptr ? p.col_len[b0 + min(jbase + 4 * i + gate, nb_valid - 1)] : T;
    }
    // input pre-activations are prefetched TWO steps ahead (HBM latency under load exceeds the MMA phase)
    // All prefetch loads are UNCONDITIONAL (a `cond ? load : 0` select would make the warp wait for the load
    // at the select): pad columns / out-of-range steps read a valid neighbouring address and the value is
    // simply never used for anything that is kept (pad columns of h are neither gathered nor written out).
    float gpre[NBH], gnext[NBH];
    int jcl[NBH];   // clamped batch column per register slot
#pragma unroll
    for (int j = 0; j < NBH; ++j) jcl[j] = b0 + min(jbase + j, nb_valid - 1);
    {
      const int t0 = dir ? T - 1 : 0;
      const int t1 = T > 1 ? (dir ? T - 2 : 1) : t0;
#pragma unroll
      for (int j = 0; j < NBH; ++j) {
        gpre[j] = __ldcs(gcol + ((long long)t0 * (CHUNK ? p.Bp : p.B) + jcl[j]) * ldg);
        gnext[j] = __ldcs(gcol + ((long long)t1 * (CHUNK ? p.Bp : p.B) + jcl[j]) * ldg);
      }
    }
    if (!TC) mbar_wait(wbar, 0);

    for (int s = 0; s < T; ++s) {
      const int t = dir ? T - 1 - s : s;
      float acc[NBH];
      if (s == 0) {
#pragma unroll
        for (int j = 0; j < NBH; ++j) acc[j] = 0.f;
      } else if (TC) {
        mbar_wait(mbar, (s - 1) & 1);
        tc_fence_after_sync();
        if (tid == 0) { REC_TRACE(8); REC_GTRACE(0); }
        uint32_t v[NACC][NBH];
#pragma unroll
        for (int a = 0; a < NACC; ++a)
          tmem_ld_n<NBH>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * NBP + jbase, v[a]);
        tmem_wait_ld();
        const int nacc = min(NACC, Hp / 16);   // accumulators actually written (tiny Hp: fewer k-steps)
#pragma unroll
        for (int j = 0; j < NBH; ++j) {
          float a = __uint_as_float(v[0][j]);
#pragma unroll
          for (int q2 = 1; q2 < NACC; ++q2) a += (nacc > q2) ? __uint_as_float(v[q2][j]) : 0.f;
          acc[j] = a;
        }
        if (tid == 0) REC_TRACE(9);
      } else {
#pragma unroll
        for (int j = 0; j < NBH; ++j) acc[j] = 0.f;
        const uint4* wv = reinterpret_cast<const uint4*>(w_s + (size_t)r * Hp * 2);
        const uint4* hv = reinterpret_cast<const uint4*>(h_s);
        for (int kc = 0; kc < Hp / 8; ++kc) {
          const uint4 w8 = wv[kc];
          const __half2* wh = reinterpret_cast<const __half2*>(&w8);
#pragma unroll
          for (int j = 0; j < NBH; ++j) {
            const uint4 h8 = hv[kc * NBP + jbase + j];
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
        named_bar_sync(1, GT);   // everyone finished reading h_s before it is overwritten below
      }

      // activation of this thread's gate for its NBH batch columns
#pragma unroll
      for (int j = 0; j < NBH; ++j) {
        const float pre = acc[j] + gpre[j];
        float av;
        if constexpr (FAST) {
          // sigmoid(x) = 0.5*tanh(0.5x)+0.5 ; tanh(x) itself for gate g
          const float th = tanh_fast((gate == 2) ? pre : 0.5f * pre);
          av = (gate == 2) ? th : fmaf(0.5f, th, 0.5f);
        } else {
          av = fmaf(ak, sigmoid_f(ak * pre), ab);      // ak = 1 -> sigmoid, ak = 2 -> tanh
        }
        xch[r * XP + jbase + j] = av;
        // in place: this element was consumed (prefetched) two steps ago
        if (p.act_out != nullptr && jbase + j < nb_valid)
          p.act_out[((long long)t * (CHUNK ? p.Bp : p.B) + b0 + jbase + j) * ldg + (long long)dir * 4 * Hp + rb * 128 + r] = av;
      }
      __syncwarp();
      if (tid == 0) REC_TRACE(10);

      uint8_t* lldst = (s & 1) ? ll1 : ll0;
      const unsigned int fbit = (((unsigned int)s >> 1) & 1u) ^ 1u;   // buffers start zeroed -> first use expects 1
      const unsigned int fword = fbit ? 0x40004000u : 0u;
      float hval[NBH / 4];
#pragma unroll
      for (int ci = 0; ci < NBH / 4; ++ci) {
        const int j = jbase + 4 * ci + gate;  // this lane finishes batch column j of unit ul
        const float* xr = xch + (4 * ul) * XP + j;
        const float gi = xr[0];
        const float gf = xr[XP];
        const float gg = xr[2 * XP];
        const float go = xr[3 * XP];
#ifdef ONSSEN_REC_NO_LEN
        const bool live = true;
#else
        const bool live = t < len_ci[ci];            // padded frames of a shorter utterance: zero state, zero output
#endif
        const float c = live ? fmaf(gf, c_state[ci], gi * gg) : 0.f;
        c_state[ci] = c;
        const float h = live ? go * (FAST ? tanh_fast(c) : fmaf(2.0f, sigmoid_f(2.0f * c), -1.0f)) : 0.f;
        hval[ci] = h;
        // publish: the 8 units of this warp (one k-chunk) x column j form one 16-byte chunk of the operand
        // tile; they sit in the 8 lanes that share this gate index.  Assemble the chunk with a 3-level
        // butterfly and let the ul_w == 0 lane issue ONE 16-byte store (atomic w.r.t. the readers' 16-byte
        // loads: no half-updated chunks, 8x fewer L2 write transactions than per-pair 4-byte stores).
        const unsigned int ulw = (unsigned int)lane >> 2;      // unit within the warp, 0..7
        const float hp = __shfl_xor_sync(0xffffffffu, h, 4);
        const __half2 pk = (ulw & 1) ? __floats2half2_rn(hp, h) : __floats2half2_rn(h, hp);
        const unsigned int w1 = *reinterpret_cast<const unsigned int*>(&pk) | fword;
        const unsigned int w1o = __shfl_xor_sync(0xffffffffu, w1, 8);
        const unsigned int d0 = (ulw & 2) ? w1o : w1, d1 = (ulw & 2) ? w1 : w1o;
        const unsigned int e0 = __shfl_xor_sync(0xffffffffu, d0, 16), e1 = __shfl_xor_sync(0xffffffffu, d1, 16);
        if (ulw == 0)
          st_relaxed_v4(lldst + (size_t)((u >> 3) * NBP + j) * 16, make_uint4(d0, d1, e0, e1));
      }
      if (tid == 0) { REC_TRACE(11); REC_GTRACE(1); }
      // layer output (plain stores, not needed by the other CTAs) and the prefetch two steps ahead are issued
      // between publish and gather: their latency is absorbed by the wait for the other CTAs' h_t
#pragma unroll
      for (int ci = 0; ci < NBH / 4; ++ci) {
        const int j = jbase + 4 * ci + gate;
        if (j < nb_valid) {
          const int m = t * (CHUNK ? p.Bp : p.B) + b0 + j;
          const long long o = (long long)m * ldy + dir * Hp + u;
          float hv = hval[ci];
          if (p.dropout_p > 0.f) {
            const float rnd = hash_uniform32(p.seed_lo, p.seed_hi, (unsigned int)o + (CHUNK ? p.hash_off : 0u));
            hv = rnd < p.dropout_p ? 0.f : hv * keep_scale;
          }
          if (p.y_h) p.y_h[o] = __float2half_rn(hv);
          if (p.y_f) p.y_f[o] = hv;
          if (p.c_out) p.c_out[o] = c_state[ci];
          if (p.h_raw) p.h_raw[o] = __float2half_rn(hval[ci]);
        }
      }
      auto prefetch_gates = [&]() {
        // rotate the prefetch registers and issue the loads for step s+2
        const int tn = (s + 2 < T) ? (dir ? t - 2 : t + 2) : t;   // clamped: the last two loads are unused
#pragma unroll
        for (int j = 0; j < NBH; ++j) {
          gpre[j] = gnext[j];
          gnext[j] = __ldcs(gcol + ((long long)tn * (CHUNK ? p.Bp : p.B) + jcl[j]) * ldg);
        }
      };
      if (!ONSSEN_REC_PREFETCH_AFTER) prefetch_gates();
      if (s + 1 < T) {
        // gather h_t of the whole group (all nrb producers) into the smem operand tile: spin on the flag bits
        uint4 v[MAXCH];
        unsigned int pending = want_mask;
        if (p.poll_delay > 0) {   // a poll issued before the other CTAs' stores reach L2 costs a full round trip
          const long long t_go = clock64() + p.poll_delay;
          while (clock64() < t_go) {
          }
        }
        while (pending) {
#pragma unroll
          for (int i = 0; i < MAXCH; ++i)
            if (pending & (1u << i)) v[i] = ld_relaxed_v4(lldst + (size_t)(tid + GT * i) * 16);
#pragma unroll
          for (int i = 0; i < MAXCH; ++i)
            if (pending & (1u << i)) {
              const unsigned int m = 0x40004000u;
              const bool fresh = ((v[i].x & m) == fword) && ((v[i].y & m) == fword) && ((v[i].z & m) == fword) &&
                                 ((v[i].w & m) == fword);
              if (fresh) {
                *reinterpret_cast<uint4*>(h_s + (size_t)(tid + GT * i) * 16) =
                    make_uint4(v[i].x & ~m, v[i].y & ~m, v[i].z & ~m, v[i].w & ~m);
                pending &= ~(1u << i);
              }
            }
        }
        if (tid == 0) { REC_TRACE(12); REC_GTRACE(2); }
        if (TC) {
          fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
          tc_fence_before_sync();
          asm volatile("bar.arrive 3, %0;" ::"r"(REC_THREADS) : "memory");
        } else {
          named_bar_sync(1, GT);
        }
        if (tid == 0) REC_TRACE(13);
      }
      if (ONSSEN_REC_PREFETCH_AFTER) prefetch_gates();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (TC && warp == REC_GATE_WARPS) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NB, bool TC>
size_t rec_smem_bytes(int Hp) {
  constexpr int NBP = NB <= 16 ? 16 : 32;
  return (TC ? 0 : (size_t)Hp * 128 * 2) + (size_t)Hp * NBP * 2 + (size_t)(128 * (NB + 1) + 1) * 4 + 64;
}

template <int NB, bool TC>
int launch_rec(RecParams& p, int grid, cudaStream_t stream) {
  const size_t smem = rec_smem_bytes<NB, TC>(p.Hp);
  constexpr int REC_THREADS = rec_threads(NB, TC);
  // approximate activations only on the tensor-core inference path (no BPTT state saved)
  const bool fast = TC && p.act_out == nullptr;
  auto kern = (p.Bp != p.B) ? (fast ? blstm_rec_kernel<NB, TC, true, TC> : blstm_rec_kernel<NB, TC, true, false>)
                            : (fast ? blstm_rec_kernel<NB, TC, false, TC> : blstm_rec_kernel<NB, TC, false, false>);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return ONSSEN_ERR_CUDA;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, REC_THREADS, smem) != cudaSuccess)
    return ONSSEN_ERR_CUDA;
  if (per_sm > 1 && TC) per_sm = 1;   // each CTA allocates all 512 TMEM columns
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
    case 16: return launch_rec<16, TC>(p, grid, stream);
    case 24: if constexpr (TC) return launch_rec<24, TC>(p, grid, stream); else return launch_rec<32, TC>(p, grid, stream);
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
  const int Hp = hp_of(H);
  const int nrb = Hp / 32;
  int smax = num_sms() / (2 * nrb);
  // TMEM: NACC accumulators + Hp/2 weight columns; gather: Hp*NBP/8 chunks over 256 gate threads, at most
  // 6 (NBP=16) / 12 (NBP=32) per thread  <=>  Hp <= 768
  if (smax < 1 || Hp / 2 + TMEM_A_COL > 512 || Hp > 768) { sp.ok = false; return sp; }
  if (smax > B) smax = B;
  int bs = (B + smax - 1) / smax;
  static const int opts[] = {16, 24, 32};
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

extern "C" void onssen_blstm_rec_set_trace(void* device_buf_64_int64) {
  g_trace_ptr = (long long*)device_buf_64_int64;
}

extern "C" void onssen_blstm_rec_set_poll_delay(int cycles) { g_poll_delay = cycles < 0 ? 0 : cycles; }

extern "C" size_t onssen_blstm_rec_workspace_bytes(int B, int H) {
  if (B <= 0 || H <= 0) return 0;
  const int Hp = hp_of(H);
  const int nrb = Hp / 32;
  int smax = num_sms() / (2 * nrb);
  if (smax < 1) smax = 1;
  // upper bound independent of the exact plan: NBP = 32
  return 256 + (size_t)2 * 2 * smax * Hp * 32 * 2;
}

static int rec_fwd_impl(const float* gates, const void* whh_p, int B, int T, int H, void* y_h, float* y_f,
                        float dropout_p, unsigned long long seed, unsigned long long offset, void* workspace,
                        size_t workspace_bytes, int use_tensor_cores, void* stream, float* act_out, float* c_out,
                        void* h_raw, const int* col_len = nullptr);

extern "C" int onssen_blstm_rec_fwd(const float* gates, const void* whh_p, int B, int T, int H, void* y_h,
                                    float* y_f, float dropout_p, unsigned long long seed,
                                    unsigned long long offset, void* workspace, size_t workspace_bytes,
                                    int use_tensor_cores, void* stream) {
  return rec_fwd_impl(gates, whh_p, B, T, H, y_h, y_f, dropout_p, seed, offset, workspace, workspace_bytes,
                      use_tensor_cores, stream, nullptr, nullptr, nullptr);
}

extern "C" int onssen_blstm_rec_fwd_var(const float* gates, const void* whh_p, int B, int T, int H, void* y_h,
                                        float* y_f, const int32_t* frames_per_utt, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  if (!frames_per_utt) return ONSSEN_ERR_ARG;
  return rec_fwd_impl(gates, whh_p, B, T, H, y_h, y_f, 0.f, 0, 0, workspace, workspace_bytes, 1, stream, nullptr, nullptr,
                      nullptr, frames_per_utt);
}

extern "C" int onssen_blstm_rec_fwd_train(float* gates_inout, const void* whh_p, int B, int T, int H, void* y_h,
                                          float* y_f, float* c_out, void* h_raw, float dropout_p,
                                          unsigned long long seed, unsigned long long offset, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  if (!c_out) return ONSSEN_ERR_ARG;
  return rec_fwd_impl(gates_inout, whh_p, B, T, H, y_h, y_f, dropout_p, seed, offset, workspace, workspace_bytes, 1,
                      stream, gates_inout, c_out, h_raw);
}

static int rec_fwd_impl(const float* gates, const void* whh_p, int B, int T, int H, void* y_h, float* y_f,
                        float dropout_p, unsigned long long seed, unsigned long long offset, void* workspace,
                        size_t workspace_bytes, int use_tensor_cores, void* stream, float* act_out, float* c_out,
                        void* h_raw, const int* col_len) {
  if (!gates || !whh_p || !workspace || B <= 0 || T <= 0 || H <= 0) return ONSSEN_ERR_ARG;
  if (!y_h && !y_f) return ONSSEN_ERR_ARG;
  if (dropout_p < 0.f || dropout_p >= 1.f) return ONSSEN_ERR_ARG;
  // batches above one launch's capacity (32 columns x co-resident CTAs per row block: 96 utterances at H=600 on 148
  // SMs) run as column chunks of nearly equal size, back to back on the stream (torch.nn.LSTM has no batch limit:
  // deep_clustering.py:34-35)
  const int Hp = hp_of(H);
  const int smax = num_sms() / (2 * (Hp / 32));
  if (smax < 1) return ONSSEN_ERR_UNSUPPORTED;
  const int bmax = smax * 32;
  const int nchunk = (B + bmax - 1) / bmax;
  const int per = (B + nchunk - 1) / nchunk;
  RecParams p;
  p.whh = (const __half*)whh_p;
  p.Bp = B; p.T = T; p.H = H; p.Hp = Hp; p.nrb = p.Hp / 32;
  p.dropout_p = dropout_p;
  const unsigned long long mix = seed * 0x9E3779B97F4A7C15ull + offset * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull;
  p.seed_lo = (unsigned int)mix;
  p.seed_hi = (unsigned int)(mix >> 32);
  p.trace = g_trace_ptr;
  p.poll_delay = g_poll_delay;
  p.hbuf = (__half*)((uint8_t*)workspace + 256);
  cudaStream_t s = (cudaStream_t)stream;
  for (int c0 = 0; c0 < B; c0 += per) {
    const int bc = (B - c0) < per ? (B - c0) : per;
    const SlicePlan sp = plan_slices(bc, H);
    if (!sp.ok) return ONSSEN_ERR_UNSUPPORTED;
    // the chunk's first column is folded into the base pointers (row m = t*Bp + b: column offset = b rows)
    const size_t og = (size_t)c0 * 8 * Hp, oy = (size_t)c0 * 2 * Hp;
    p.gates = gates + og;
    p.y_h = y_h ? (__half*)y_h + oy : nullptr;
    p.y_f = y_f ? y_f + oy : nullptr;
    p.act_out = act_out ? act_out + og : nullptr;
    p.c_out = c_out ? c_out + oy : nullptr;
    p.h_raw = h_raw ? (__half*)h_raw + oy : nullptr;
    p.col_len = col_len ? col_len + c0 : nullptr;
    p.hash_off = (unsigned int)oy;
    p.B = bc; p.S = sp.S; p.Bs = sp.Bs;
    const size_t hbuf_bytes = (size_t)2 * 2 * sp.S * p.Hp * sp.NBP * 2;
    if (workspace_bytes < 256 + hbuf_bytes) return ONSSEN_ERR_ARG;
    if (cudaMemsetAsync(workspace, 0, 256 + hbuf_bytes, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
    const int grid = 2 * sp.S * p.nrb;
    const int rc = use_tensor_cores ? dispatch_nb<true>(sp.NB, p, grid, s) : dispatch_nb<false>(sp.NB, p, grid, s);
    if (rc != ONSSEN_OK) return rc;
  }
  return ONSSEN_OK;
}
