// BPTT for one BLSTM layer (autograd of torch.nn.LSTM as used at onssen/nn/deep_clustering.py:34-35; the
// reference gets it from cuDNN through loss.backward(), onssen/utils/train.py:82).
//
// This file: the mma.sync kernels -- the per-step kernel (validation partner of the differential tests) and the
// persistent kernel of round 1 (fallback for shapes the tcgen05 cluster kernel of lstm_bwd_tc.cu does not take) -- the
// W_hh^T packing for both implementations, and the C entry point that dispatches between the three.
//
// Per-step kernel: one launch per time step (both directions), no inter-CTA synchronisation inside a launch: CTA (unit block,
// dir, batch block) owns 32 hidden units.  It first forms dh_rec[u][b] = sum_r W_hh[r][u] * dG_{prev}[r][b]
// over ALL 4H gate rows of the previously processed step (mma.sync m16n8k16, fp16 operands, W_hh^T slice and
// dG read from L2, K split over the 8 warps), then does the gate math for ITS units at this step and writes
// dG (fp32, in place over the saved activations, and as scaled fp16 for the next launch / the wgrad GEMMs).
// dh_rec and the cell-gradient carry never leave the CTA's unit block.  L2-bound: 2 x 4Hp x (32+B) fp16 per CTA.
#include "common.cuh"
#include "lstm_bwd.cuh"

namespace onssen {
namespace {

long long* g_bwd_trace = nullptr;
#define BWD_TRACE_S0 100
#define BWD_TRACE(slot)                                                                                          \
  do {                                                                                                           \
    if (p.trace != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (threadIdx.x & 31) == 0 &&          \
        s >= BWD_TRACE_S0 && s < BWD_TRACE_S0 + 8)                                                               \
      p.trace[(s - BWD_TRACE_S0) * 64 + (slot) * 8 + (threadIdx.x >> 5)] = clock64();                            \
  } while (0)

__device__ __forceinline__ float hash_uniform32(unsigned int seed_lo, unsigned int seed_hi, unsigned int idx) {
  unsigned int x = idx ^ seed_lo;
  x *= 0x9E3779B1u; x ^= x >> 15;
  x *= 0x85EBCA77u; x ^= x >> 13;
  x += seed_hi;
  x *= 0xC2B2AE3Du; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

int g_bwd_persistent = 2;   // 2: tcgen05 cluster kernel, 1: mma.sync persistent kernel, 0: one launch per step (validation)

__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256) lstm_bwd_step_kernel(const BwdParams p) {
  __shared__ float red[8][32][33];   // per-warp partial dh_rec[unit][batch]
  const int Hp = p.Hp, B = p.B, T = p.T;
  const int G4 = 4 * Hp;
  const int ub = blockIdx.x, dir = blockIdx.y, bb = blockIdx.z;
  const int b0 = bb * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int t = dir == 0 ? T - 1 - p.s : p.s;          // forward-time index handled by this launch
  const int t_pp = dir == 0 ? t + 1 : t - 1;           // step processed by the previous launch
  const int t_fp = dir == 0 ? t - 1 : t + 1;           // forward-time predecessor (c_{t-1} of this direction)
  const float inv_scale = p.scale2[1], scale = p.scale2[0];

  if (p.s > 0) {
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
    const int ksteps = G4 / 16;
    const int nub = Hp / 32, nbb = gridDim.z;
    const uint4* wfr = reinterpret_cast<const uint4*>(p.wt) + ((size_t)(dir * nub + ub) * ksteps) * 2 * 32 + lane;
    const uint2* bfr = reinterpret_cast<const uint2*>(p.frag) +
                       ((((size_t)((p.s - 1) & 1) * 2 + dir) * nbb + bb) * ksteps) * 4 * 32 + lane;
    // each warp owns a contiguous range of k-steps; operands are pre-laid-out in fragment order, so every load is
    // one fully coalesced 16-byte (A) / 8-byte (B) access; U k-steps of loads are issued before their MMAs
    const int per_warp = (ksteps + 7) / 8;
    const int ks_begin = warp * per_warp;
    const int ks_end = min(ksteps, ks_begin + per_warp);
    constexpr int U = 4;    // measured: U=10 makes ptxas serialise the loads again (64 regs) and is slower
    for (int ksb = ks_begin; ksb < ks_end; ksb += U) {
      uint4 a[U][2];
      uint2 bf[U][4];
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {
        const int ks = min(ksb + uu, ks_end - 1);          // clamped: duplicates are skipped below
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) a[uu][mt] = __ldg(wfr + ((size_t)ks * 2 + mt) * 32);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bf[uu][nt] = bfr[((size_t)ks * 4 + nt) * 32];
      }
#pragma unroll
      for (int uu = 0; uu < U; ++uu) {
        if (ksb + uu < ks_end) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
              mma_16816(acc[mt][nt], reinterpret_cast<const uint32_t*>(&a[uu][mt]),
                        reinterpret_cast<const uint32_t*>(&bf[uu][nt]));
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        red[warp][mt * 16 + g][nt * 8 + 2 * tq] = acc[mt][nt][0];
        red[warp][mt * 16 + g][nt * 8 + 2 * tq + 1] = acc[mt][nt][1];
        red[warp][mt * 16 + g + 8][nt * 8 + 2 * tq] = acc[mt][nt][2];
        red[warp][mt * 16 + g + 8][nt * 8 + 2 * tq + 1] = acc[mt][nt][3];
      }
  }
  __syncthreads();

  const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  for (int idx = tid; idx < 32 * 32; idx += 256) {
    const int ul = idx & 31, bl = idx >> 5;
    const int b = b0 + bl;
    if (b >= B) continue;
    const int u = ub * 32 + ul;
    float dh_rec = 0.f;
    if (p.s > 0) {
#pragma unroll
      for (int w = 0; w < 8; ++w) dh_rec += red[w][ul][bl];
      dh_rec *= inv_scale;
    }
    const long long m = (long long)t * B + b;
    const long long oy = m * (2 * Hp) + dir * Hp + u;
    float dyv = p.dy[oy];
    if (p.dropout_p > 0.f) {
      const float rnd = hash_uniform32(p.seed_lo, p.seed_hi, (unsigned int)oy);
      dyv = rnd < p.dropout_p ? 0.f : dyv * keep_scale;
    }
    const float dh = dyv + dh_rec;
    float4* gp = reinterpret_cast<float4*>(p.actg + m * (2 * G4) + dir * G4 + ub * 128 + 4 * ul);
    const float4 a4 = *gp;                              // i, f, g, o (activated)
    const float ct = p.c[oy];
    float cprev = 0.f;
    if (t_fp >= 0 && t_fp < T) cprev = p.c[((long long)t_fp * B + b) * (2 * Hp) + dir * Hp + u];
    const float tc = tanhf(ct);
    float* dcp = p.dc + ((size_t)dir * B + b) * Hp + u;
    const float dct = dh * a4.w * (1.0f - tc * tc) + (p.s > 0 ? *dcp : 0.f);
    float4 d4;
    d4.x = dct * a4.z * a4.x * (1.0f - a4.x);          // di (pre-activation)
    d4.y = dct * cprev * a4.y * (1.0f - a4.y);         // df
    d4.z = dct * a4.x * (1.0f - a4.z * a4.z);          // dg
    d4.w = dh * tc * a4.w * (1.0f - a4.w);             // do
    *dcp = dct * a4.y;
    *gp = d4;
    __half2 lo = __halves2half2(to_half_sat(d4.x * scale), to_half_sat(d4.y * scale));
    __half2 hi = __halves2half2(to_half_sat(d4.z * scale), to_half_sat(d4.w * scale));
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p.dg16 + m * (2 * G4) + dir * G4 + ub * 128 + 4 * ul) = o;
    // the same values in the next launch's B-fragment order: row r = ub*128 + 4*ul + gate is k-index
    // kk = (4*ul + gate) % 16 of k-step ub*8 + ul/4; words pair (kk, kk+1); column n = bl
    {
      const int ks = ub * 8 + (ul >> 2);
      const int j = ul & 3;                       // kk = 4*j + gate
      const int reg = j >> 1;                     // kk >= 8 -> second B register
      const int tq0 = (j & 1) * 2;                // (kk % 8) / 2 for gate pair (0,1); +1 for gates (2,3)
      uint32_t* fb = p.frag + (((((size_t)(p.s & 1) * 2 + dir) * gridDim.z + bb) * (G4 / 16) + ks) * 4 + (bl >> 3)) * 64;
      fb[((bl & 7) * 4 + tq0) * 2 + reg] = o.x;
      fb[((bl & 7) * 4 + tq0 + 1) * 2 + reg] = o.y;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent variant: ONE cooperative launch per layer.  CTA (unit block, dir, batch block) keeps its W_hh^T
// fragments in shared memory for all T steps; the per-step exchange of dG between the CTAs of a direction uses
// the forward kernel's flag-bit protocol: the caller picks the loss scale so that |dG * scale| < 2, which leaves
// bit 14 of every fp16 free as a step-parity flag -> self-validating data, no fences, no grid barrier.
__device__ __forceinline__ uint2 ld_relaxed_v2(const void* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(void* p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ __half to_half_flag_range(float x) {   // clamp to the largest fp16 below 2.0
  x = fminf(fmaxf(x, -1.9990234375f), 1.9990234375f);
  return __float2half_rn(x);
}

// NT = 8-column n-tiles per batch block (2 -> 16 utterances per CTA, 4 -> 32): the launcher takes the smallest NT whose
// grid still fits one CTA per SM, so that at B=32, H=600 76 CTAs share the work instead of 38.
//
// Step anatomy (measured, scripts/bwd_trace.py, profiles/r01_bwd_step_trace.txt): the exchange wait dominates --
// every CTA pulls the whole dG of its (direction, batch block), 4Hp x NB fp16 = 78 KB at NT=2, through its own
// L2->SM port each step.  So (1) a burst that comes back stale wastes a full transfer: one fragment per producer
// CTA is probed first and the burst (KPW k-steps x NT fragments per lane, ONE L2 round trip) is only issued
// when the probes carry the step's flag bits; (2) everything of the gate math that does not depend on dh_rec
// (tanh, dropout, the products of saved activations) is computed while the burst is in flight, which leaves
// ~10 dependent instructions between the partial-sum barrier and the publish stores; (3) the publish stores go
// before the dG stores that only the later GEMMs read.
// shared-memory accesses by 32-bit shared address (no generic->shared conversion on the critical path)
__device__ __forceinline__ uint32_t smem_addr32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void sts_v4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float2 lds_v2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
// pin a value in a register: without this ptxas re-derives loop invariants from special registers (S2R / S2UR of
// %ctaid, the shared window base) and constant-bank loads INSIDE the step loop, on the critical path after the barrier
template <typename Tp>
__device__ __forceinline__ void pin(Tp*& x) { asm volatile("" : "+l"(x)); }
__device__ __forceinline__ void pin(uint32_t& x) { asm volatile("" : "+r"(x)); }

template <int NT>
__global__ void __launch_bounds__(256, 1) lstm_bwd_persistent_kernel(const BwdParams p) {
  constexpr int KPW = NT == 2 ? 19 : 8;      // fragment k-steps held in registers at once (NT*KPW uint2 per lane)
  constexpr int NB = 8 * NT;                  // batch columns per CTA
  constexpr int ITEMS = NT;                   // (unit, batch) pairs per thread: 32 * NB / 256
  extern __shared__ __align__(16) uint8_t bsm[];
  const int Hp = p.Hp, B = p.B, T = p.T;
  const int G4 = 4 * Hp;
  const int ksteps = G4 / 16;
  uint4* a_s = reinterpret_cast<uint4*>(bsm);                                   // [ksteps][2][32] uint4
  // partial sums of the 8 K-split warps in accumulator-fragment order, double-buffered over steps:
  // red[step parity][source warp][tile = mt*NT + nt][lane] = that lane's 4 accumulators (one STS.128 per tile;
  // a __syncthreads drains pending shared stores at ~36 cycles each with 8 warps, so their COUNT is what matters)
  float4* red = reinterpret_cast<float4*>(bsm + (size_t)ksteps * 2 * 32 * 16);
  constexpr int RED_STEP = 8 * 2 * NT * 32;   // float4 per parity
  const int ub = blockIdx.x, dir = blockIdx.y, bb = blockIdx.z;
  const int nub = Hp / 32, nbb = gridDim.z;
  const int b0 = bb * NB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const float inv_scale = p.scale2[1], scale = p.scale2[0];
  const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  // resident W_hh^T fragments
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wt) + ((size_t)(dir * nub + ub) * ksteps) * 2 * 32;
    for (int i = tid; i < ksteps * 64; i += 256) a_s[i] = __ldg(src + i);
  }
  __syncthreads();
  const int per_warp = (ksteps + 7) / 8;
  const int ks_begin = warp * per_warp;
  const int ks_end = min(ksteps, ks_begin + per_warp);
  const size_t frag_group = (size_t)ksteps * NT * 32;   // uint2 per (parity, dir, bb)
  float dc_carry[ITEMS];                                 // cell-gradient carry of this thread's (unit, batch) pairs
  // this thread's publish slots (word offsets inside one (parity, dir, bb) fragment group) and output offsets
  unsigned int pub_off[ITEMS];
  long long row_off[ITEMS];
  // phase-B ownership follows the accumulator fragments: warp -> tile (mt, nt) and which of the lane's 4 values
  // (NT=2: two warps share a tile, rows g / g+8; NT=4: one warp per tile, all 4 values)
  const int my_tile = warp / (4 / NT), my_v0 = (warp % (4 / NT)) * NT;
  int item_ul[ITEMS], item_bl[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    dc_carry[i] = 0.f;
    const int v = my_v0 + i;                      // value index inside the m16n8 accumulator fragment
    const int ul = (my_tile / NT) * 16 + g + 8 * (v >> 1), bl = (my_tile % NT) * 8 + 2 * tq + (v & 1);
    item_ul[i] = ul;
    item_bl[i] = bl;
    const int ks = ub * 8 + (ul >> 2);
    const int j = ul & 3;
    // row r = ub*128 + 4*ul + gate is k-index 4*j + gate of k-step ks; word pair (2 gates) -> register j>>1
    pub_off[i] = (unsigned int)((((size_t)ks * NT + (bl >> 3)) * 32 + (bl & 7) * 4 + (j & 1) * 2) * 2 + (j >> 1));
    row_off[i] = (long long)min(b0 + bl, B - 1) * (2 * G4) + dir * G4 + ub * 128 + 4 * ul;
  }
  // loop invariants of the latency-critical parts, pinned in registers (by step parity)
  const uint2* bfr_par[2];    // this CTA's (dir, bb) fragment group, + lane: what it consumes
  uint32_t* fb_par[2];        // ... and what it publishes into
  uint32_t red_wr32[2], red_rd32[2];
  uint32_t valid_items = 0;
#pragma unroll
  for (int par = 0; par < 2; ++par) {
    bfr_par[par] = reinterpret_cast<const uint2*>(p.frag) + (((size_t)par * 2 + dir) * nbb + bb) * frag_group + lane;
    fb_par[par] = p.frag + (((size_t)par * 2 + dir) * nbb + bb) * frag_group * 2;
    red_wr32[par] = smem_addr32(red + par * RED_STEP + (warp * 2 * NT) * 32 + lane);
    red_rd32[par] = smem_addr32(reinterpret_cast<const float*>(red + par * RED_STEP + my_tile * 32 + lane) + my_v0);
    pin(bfr_par[par]); pin(fb_par[par]); pin(red_wr32[par]); pin(red_rd32[par]);
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (b0 + item_bl[i] < B) valid_items |= 1u << i;
  pin(valid_items);

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? T - 1 - s : s;
    const int t_fp = dir == 0 ? t - 1 : t + 1;
    BWD_TRACE(0);
    // ---- per-step inputs that do not depend on the recurrence: issue their loads before the exchange wait
    float4 a4[ITEMS];
    float ct[ITEMS], cprev[ITEMS], dyv[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int ul = item_ul[i], bl = item_bl[i];
      const int b = min(b0 + bl, B - 1);               // clamped: results of pad columns are replaced by zeros
      const int u = ub * 32 + ul;
      const long long oy = ((long long)t * B + b) * (2 * Hp) + dir * Hp + u;
      a4[i] = *reinterpret_cast<const float4*>(p.actg + (long long)t * B * (2 * G4) + row_off[i]);
      ct[i] = p.c[oy];
      cprev[i] = (t_fp >= 0 && t_fp < T) ? p.c[((long long)t_fp * B + b) * (2 * Hp) + dir * Hp + u] : 0.f;
      dyv[i] = p.dy[oy];
      if (p.dropout_p > 0.f) {   // the dropout decision only needs the index: fold it into a multiplier now
        const float rnd = hash_uniform32(p.seed_lo, p.seed_hi, (unsigned int)oy);
        dyv[i] *= rnd < p.dropout_p ? 0.f : keep_scale;
      }
    }
    // factors of the gate derivatives that do not depend on dh_rec:
    //   dh = dy + dh_rec;  dct = dh*k_c + carry;  (di, df, dg) = dct*(k_i, k_f, k_g);  do = dh*k_o;  carry' = dct*f
    float k_c[ITEMS], k_i[ITEMS], k_f[ITEMS], k_g[ITEMS], k_o[ITEMS], f_g[ITEMS];
    auto gate_factors = [&]() {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const float tc = tanhf(ct[i]);
        k_c[i] = a4[i].w * (1.0f - tc * tc);
        k_i[i] = a4[i].z * a4[i].x * (1.0f - a4[i].x) * scale;
        k_f[i] = cprev[i] * a4[i].y * (1.0f - a4[i].y) * scale;
        k_g[i] = a4[i].x * (1.0f - a4[i].z * a4[i].z) * scale;
        k_o[i] = tc * a4[i].w * (1.0f - a4[i].w) * scale;
        f_g[i] = a4[i].y;
      }
    };
    // ---- phase A: dh_rec = W_hh^T slice x dG_{previous step} (all gate rows of this direction)
    if (s > 0) {
      float acc[2][NT][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
      const unsigned int fbit = ((((unsigned int)(s - 1)) >> 1) & 1u) ^ 1u;
      const unsigned int fw = fbit ? 0x40004000u : 0u;
      const uint2* bfr = ((s - 1) & 1) ? bfr_par[1] : bfr_par[0];
      for (int ksb = ks_begin; ksb < ks_end; ksb += KPW) {
        const int kse = min(ksb + KPW, ks_end);
        // Fragment bookkeeping is per PRODUCER CTA (8 k-steps each, <= 4 producers in a warp's k-range), as bit masks
        // over the k-step index uu: one loop body holds the probes, the (re)load of the requested k-steps with
        // immediate offsets and the validation -- a few hundred instructions instead of several thousand, which
        // matters because the whole step body streams through the instruction cache once per step.
        const int p0 = ksb >> 3;
        unsigned int seg[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int lo = max(ksb, (p0 + q) * 8) - ksb, hi = min(kse, (p0 + q + 1) * 8) - ksb;
          seg[q] = hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
        }
        const unsigned int valid = seg[0] | seg[1] | seg[2] | seg[3];
        unsigned int waiting = (seg[0] ? 1u : 0u) | (seg[1] ? 2u : 0u) | (seg[2] ? 4u : 0u) | (seg[3] ? 8u : 0u);
        const uint2* base = bfr + (size_t)ksb * NT * 32;
        uint2 bf[KPW][NT];
        bool factors_done = ksb != ks_begin;
        while (true) {
          unsigned int issue = 0;
          if (waiting) {
            // probe the last-written n-tile of the first k-step of every producer still waited for; a producer's
            // k-steps are requested the moment ITS probe carries the step's flag bits (early transfers overlap the
            // wait for late producers)
            uint2 pv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (waiting & (1u << q))
                pv[q] = ld_relaxed_v2(bfr + ((size_t)max(ksb, (p0 + q) * 8) * NT + (NT - 1)) * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((waiting & (1u << q)) && (pv[q].x & 0x40004000u) == fw && (pv[q].y & 0x40004000u) == fw) {
                issue |= seg[q];
                waiting &= ~(1u << q);
              }
          } else {
            if (!factors_done) {   // overlaps the round trip of the last requests
              // keep the compiler from hoisting this block above the probes: it consumes the step's prefetched
              // activations (HBM latency) and would stall the warp before its first poll
#pragma unroll
              for (int i = 0; i < ITEMS; ++i)
                asm volatile("" : "+f"(a4[i].x), "+f"(a4[i].y), "+f"(a4[i].z), "+f"(a4[i].w), "+f"(ct[i]), "+f"(cprev[i]));
              gate_factors();
              factors_done = true;
            }
            // every k-step has been requested at least once: validate, re-request the stale ones
            unsigned int stale = 0;
#pragma unroll
            for (int uu = 0; uu < KPW; ++uu) {
              unsigned int t = 0;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) t |= (bf[uu][nt].x ^ fw) | (bf[uu][nt].y ^ fw);
              if (t & 0x40004000u) stale |= 1u << uu;
            }
            stale &= valid;
            if (!stale) break;
            issue = stale;
          }
#pragma unroll
          for (int uu = 0; uu < KPW; ++uu)
            if (issue & (1u << uu)) {
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) bf[uu][nt] = ld_relaxed_v2(base + (uu * NT + nt) * 32);
            }
        }
        BWD_TRACE(1);
        // A fragments double-buffered in registers: the shared-memory load of k-step uu+1 is in flight during the
        // MMAs of k-step uu
        uint4 an0 = a_s[((size_t)ksb * 2 + 0) * 32 + lane];
        uint4 an1 = a_s[((size_t)ksb * 2 + 1) * 32 + lane];
#pragma unroll
        for (int uu = 0; uu < KPW; ++uu) {
          if (ksb + uu < kse) {
            const uint4 a0 = an0, a1 = an1;
            const int kn = min(ksb + uu + 1, ks_end - 1);
            an0 = a_s[((size_t)kn * 2 + 0) * 32 + lane];
            an1 = a_s[((size_t)kn * 2 + 1) * 32 + lane];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const uint2 bv = make_uint2(bf[uu][nt].x & ~0x40004000u, bf[uu][nt].y & ~0x40004000u);   // strip the flags
              mma_16816(acc[0][nt], reinterpret_cast<const uint32_t*>(&a0), reinterpret_cast<const uint32_t*>(&bv));
              mma_16816(acc[1][nt], reinterpret_cast<const uint32_t*>(&a1), reinterpret_cast<const uint32_t*>(&bv));
            }
          }
        }
      }
      BWD_TRACE(5);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          sts_v4(((s & 1) ? red_wr32[1] : red_wr32[0]) + (mt * NT + nt) * 32 * 16,
                 make_float4(acc[mt][nt][0], acc[mt][nt][1], acc[mt][nt][2], acc[mt][nt][3]));
    } else {
      gate_factors();
    }
    BWD_TRACE(2);
    __syncthreads();
    BWD_TRACE(3);
    // ---- phase B: gate derivatives of this CTA's units at step t; publish dG in B-fragment order
    const unsigned int fbit_w = ((((unsigned int)s) >> 1) & 1u) ^ 1u;
    const unsigned int fww = fbit_w ? 0x40004000u : 0u;
    uint32_t* fb = (s & 1) ? fb_par[1] : fb_par[0];
    float4 d4v[ITEMS];
    uint2 ov[ITEMS];
    float rsum[ITEMS];
    if (s > 0) {
      const uint32_t rp = (s & 1) ? red_rd32[1] : red_rd32[0];
      float r[8][ITEMS];
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        if (NT == 2) {
          const float2 v = lds_v2(rp + w * (2 * NT * 32 * 16));
          r[w][0] = v.x; r[w][1] = v.y;
        } else {
          const float4 v = lds_v4(rp + w * (2 * NT * 32 * 16));
          r[w][0] = v.x; r[w][1] = v.y; r[w][ITEMS - 2] = v.z; r[w][ITEMS - 1] = v.w;
        }
      }
#pragma unroll
      for (int i = 0; i < ITEMS; ++i)
        rsum[i] = (((r[0][i] + r[1][i]) + (r[2][i] + r[3][i])) + ((r[4][i] + r[5][i]) + (r[6][i] + r[7][i]))) * inv_scale;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int bl = item_bl[i];
      float dh = dyv[i];
      if (s > 0) dh += rsum[i];
      const float dct = dh * k_c[i] + dc_carry[i];
      dc_carry[i] = dct * f_g[i];
      // scaled values (the k_* carry the loss scale); the fp32 copy is unscaled again below
      const float s_i = dct * k_i[i], s_f = dct * k_f[i], s_g = dct * k_g[i], s_o = dh * k_o[i];
      uint2 o = make_uint2(0u, 0u);
      if (valid_items & (1u << i)) {   // pad batch columns publish zeros so that every fragment entry turns fresh
        __half2 lo = __halves2half2(to_half_flag_range(s_i), to_half_flag_range(s_f));
        __half2 hi = __halves2half2(to_half_flag_range(s_g), to_half_flag_range(s_o));
        // |x| >= 2 (or NaN: the comparison is false) cannot be represented next to the flag bit: the clamped value
        // is published and the event is counted so that the host can surface it (never silent)
        if (!(fmaxf(fmaxf(fabsf(s_i), fabsf(s_f)), fmaxf(fabsf(s_g), fabsf(s_o))) < 1.9995f) && p.sat != nullptr)
          atomicAdd(p.sat, 1u);
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
      }
      if (s + 1 < T) {       // the exchange first: it is on the critical path of every CTA of this direction
        st_relaxed_u32(fb + pub_off[i], o.x | fww);
        st_relaxed_u32(fb + pub_off[i] + 2, o.y | fww);
      }
      d4v[i] = make_float4(s_i * inv_scale, s_f * inv_scale, s_g * inv_scale, s_o * inv_scale);
      ov[i] = o;
    }
    BWD_TRACE(6);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if (valid_items & (1u << i)) {
        const long long off = (long long)t * B * (2 * G4) + row_off[i];
        *reinterpret_cast<float4*>(p.actg + off) = d4v[i];
        *reinterpret_cast<uint2*>(p.dg16 + off) = ov[i];
      }
    }
    BWD_TRACE(4);
    // no second barrier: red[] is double-buffered, and a warp can only reach its write of step s+2 after the
    // barrier of step s+1, which every warp passes after its reads of step s
  }
}

// W_hh [4H][H] fp32 (both directions) -> W_hh^T in mma.m16n8k16 A-fragment order (fp16):
// [dir][ub][kstep][mtile][lane][word w][2 halves];  word w covers row = mtile*16 + lane/4 + 8*(w&1) (hidden unit
// ub*32 + row) and k = 16*kstep + 2*(lane%4) + 8*(w>>1) + {0,1} (permuted gate row index)
__global__ void pack_whh_t_kernel(const float* __restrict__ w_f, const float* __restrict__ w_r, int H, int Hp,
                                  __half* __restrict__ out) {
  const int nub = Hp / 32, ksteps = 4 * Hp / 16;
  const long long total = 2LL * nub * ksteps * 2 * 32 * 4 * 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long tt = idx;
    const int e = (int)(tt & 1); tt >>= 1;
    const int w = (int)(tt & 3); tt >>= 2;
    const int lane = (int)(tt & 31); tt >>= 5;
    const int mt = (int)(tt & 1); tt >>= 1;
    const int ks = (int)(tt % ksteps); tt /= ksteps;
    const int ub = (int)(tt % nub);
    const int dir = (int)(tt / nub);
    const int u = ub * 32 + mt * 16 + (lane >> 2) + 8 * (w & 1);
    const int r = 16 * ks + 2 * (lane & 3) + 8 * (w >> 1) + e;
    const int rb = r >> 7, ul = (r & 127) >> 2, gate = r & 3;
    const int ur = rb * 32 + ul;
    float v = 0.f;
    if (ur < H && u < H) v = (dir ? w_r : w_f)[(long long)(gate * H + ur) * H + u];
    out[idx] = to_half_sat(v);
  }
}

// W_hh [4H][H] fp32 (both directions) -> the tcgen05 BPTT kernel's TMEM slabs (fp16):
// [dir][unit block of 128][K quarter kq][unit row r][k]: gate row n = kq*Hp + k (n = 4*unit' + gate), hidden unit
// u = 128*ub + r; zero outside [0,H).  One row = the Hp/2 TMEM words of one lane.
__global__ void pack_whh_slab_kernel(const float* __restrict__ w_f, const float* __restrict__ w_r, int H, int Hp,
                                     __half* __restrict__ out) {
  const int nub = bwd_tc_nub(Hp);
  const long long total = 2LL * nub * 4 * 128 * Hp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long tt = idx;
    const int k = (int)(tt % Hp); tt /= Hp;
    const int r = (int)(tt & 127); tt >>= 7;
    const int kq = (int)(tt & 3); tt >>= 2;
    const int ub = (int)(tt % nub);
    const int dir = (int)(tt / nub);
    const int n = kq * Hp + k;
    const int ur = n >> 2, gate = n & 3;
    const int u = ub * 128 + r;
    float v = 0.f;
    if (ur < H && u < H) v = (dir ? w_r : w_f)[(long long)(gate * H + ur) * H + u];
    out[idx] = to_half_sat(v);
  }
}

}  // namespace
}  // namespace onssen

using namespace onssen;

extern "C" size_t onssen_lstm_pack_whh_t_elems(int H) {
  if (H <= 0) return 0;
  const int Hp = hp_of(H);
  return whh_t_frag_elems(Hp) + whh_t_slab_elems(Hp);
}

extern "C" int onssen_lstm_pack_whh_t(const float* w_hh_f, const float* w_hh_r, int H, void* out, void* stream) {
  if (!w_hh_f || !w_hh_r || !out || H <= 0) return ONSSEN_ERR_ARG;
  const int Hp = hp_of(H);
  long long g = (2LL * Hp * 4 * Hp + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  pack_whh_t_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(w_hh_f, w_hh_r, H, Hp, (__half*)out);
  pack_whh_slab_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(w_hh_f, w_hh_r, H, Hp,
                                                                 (__half*)out + whh_t_frag_elems(Hp));
  return ONSSEN_CHECK_LAUNCH();
}

extern "C" void onssen_blstm_rec_bwd_set_persistent(int mode) { g_bwd_persistent = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
extern "C" void onssen_blstm_rec_bwd_set_trace(void* device_buf_512_int64) {
  g_bwd_trace = (long long*)device_buf_512_int64;
  bwd_tc_set_trace((long long*)device_buf_512_int64);
}

extern "C" void onssen_blstm_rec_bwd_set_sm_reserve(int sms) { bwd_tc_set_sm_reserve(sms); }

static size_t bwd_frag_bytes(int B, int Hp) { return (size_t)2 * 2 * ((B + 31) / 32) * (4 * Hp / 16) * 4 * 32 * 2 * 4; }

extern "C" size_t onssen_blstm_rec_bwd_scratch_bytes(int B, int H) {
  const int Hp = hp_of(H);
  // dc carry [2][B][Hp] fp32 + fragment exchange [2 parity][2 dir][nbb][4Hp/16][4][32][2] u32
  // (+ the tcgen05 kernel's exchange tiles; only one of the two exchange areas is used by a given launch)
  return (size_t)2 * B * Hp * 4 + bwd_frag_bytes(B, Hp) + bwd_tc_xbuf_bytes(B, Hp);
}

extern "C" int onssen_blstm_rec_bwd(float* act_gates, void* dg16, const float* c, const float* dy, const void* whh_t,
                                    void* scratch, const float* scale2, void* sat_count_u32, int B, int T, int H,
                                    float dropout_p,
                                    unsigned long long seed, unsigned long long offset, void* stream) {
  if (!act_gates || !dg16 || !c || !dy || !whh_t || !scratch || !scale2 || B <= 0 || T <= 0 || H <= 0)
    return ONSSEN_ERR_ARG;
  float* dc_carry = (float*)scratch;
  BwdParams p;
  p.actg = act_gates; p.dg16 = (__half*)dg16; p.c = c; p.dy = dy; p.wt = (const uint32_t*)whh_t; p.dc = dc_carry;
  p.frag = (uint32_t*)((uint8_t*)scratch + (size_t)2 * B * hp_of(H) * 4);
  p.xbuf = (uint8_t*)p.frag + bwd_frag_bytes(B, hp_of(H));
  p.wslab = (const __half*)whh_t + whh_t_frag_elems(hp_of(H));
  p.scale2 = scale2; p.sat = (unsigned int*)sat_count_u32; p.B = B; p.T = T; p.H = H; p.Hp = hp_of(H);
  p.dropout_p = dropout_p;
  const unsigned long long mix = seed * 0x9E3779B97F4A7C15ull + offset * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull;
  p.seed_lo = (unsigned int)mix;
  p.seed_hi = (unsigned int)(mix >> 32);
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(scratch, 0, onssen_blstm_rec_bwd_scratch_bytes(B, H), s) != cudaSuccess) return ONSSEN_ERR_CUDA;
  p.trace = g_bwd_trace;
  // tcgen05 path: W_hh^T resident in tensor memory, K split over 4-CTA clusters (lstm_bwd_tc.cu)
  if (g_bwd_persistent == 2) {
    const int rc = launch_bwd_tc(p, s);
    if (rc != ONSSEN_ERR_UNSUPPORTED) return rc;
  }
  // mma.sync persistent path: W_hh^T fragments resident in smem, one cooperative launch for all T steps
  if (g_bwd_persistent) {
    for (int nt = 2; nt <= 4; nt += 2) {
      const int nb = 8 * nt;
      dim3 pgrid(p.Hp / 32, 2, (B + nb - 1) / nb);
      const size_t smem = (size_t)(4 * p.Hp / 16) * 2 * 32 * 16 + (size_t)2 * 8 * 2 * nt * 32 * 16;
      const int nblocks = (int)(pgrid.x * pgrid.y * pgrid.z);
      if (smem > 225 * 1024 || nblocks > num_sms()) continue;
      const void* fn = nt == 2 ? (const void*)lstm_bwd_persistent_kernel<2> : (const void*)lstm_bwd_persistent_kernel<4>;
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return ONSSEN_ERR_CUDA;
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, smem) != cudaSuccess ||
          per_sm * num_sms() < nblocks)
        continue;
      void* args[] = {(void*)&p};
      if (cudaLaunchCooperativeKernel(fn, pgrid, dim3(256), args, smem, s) != cudaSuccess) return ONSSEN_ERR_CUDA;
      return ONSSEN_OK;
    }
  }
  dim3 grid(p.Hp / 32, 2, (B + 31) / 32);
  for (int step = 0; step < T; ++step) {
    p.s = step;
    lstm_bwd_step_kernel<<<grid, 256, 0, s>>>(p);
  }
  return ONSSEN_CHECK_LAUNCH();
}
