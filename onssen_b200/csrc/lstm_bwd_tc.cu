// BPTT for one BLSTM layer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as the mma.sync kernels in lstm_bwd.cu (autograd of torch.nn.LSTM as used at
// onssen/nn/deep_clustering.py:34-35, reached through loss.backward() at onssen/utils/train.py:82):
//     dh_t   = dY_t + W_hh^T dG_{t'}            (t' = the step processed before t in this direction)
//     dc_t   = dh_t * o * (1 - tanh(c_t)^2) + f_{t'} * dc_{t'}
//     dG_t   = (dc*g*i(1-i), dc*c_prev*f(1-f), dc*i*(1-g^2), dh*tanh(c)*o(1-o))
// The recurrent product has K = 4Hp gate rows for every hidden unit, four times the forward's K, so a
// 128-unit slab of W_hh^T does not fit one SM's tensor memory.  Decomposition: a CLUSTER OF 4 CTAs owns 128
// hidden units of one (direction, batch slice); CTA kq of the cluster holds the K quarter
// [kq*Hp, (kq+1)*Hp) of the slab in TENSOR MEMORY (lane = unit, Hp/2 columns: the forward kernel's footprint)
// for the whole sequence.  Per step:
//   8 worker warps : gather the K quarter of dG_{t'} (fp16, flag bit 14 = step parity, the forward kernel's
//                    self-validating exchange through L2) into the smem B tile -> fence -> bar.arrive
//   MMA warp       : Hp/16 x tcgen05.mma (A: TMEM slab, B: smem tile, D: TMEM, M=128, N=16/32, K=16) -> commit
//   8 worker warps : tcgen05.ld the partial dh (lane quarter q = the 32 units finalised by CTA q of the
//                    cluster) and REDUCE-SCATTER it through distributed shared memory: st.async into CTA q's
//                    receive tile, completion counted in bytes on CTA q's mbarrier (no flags, no fences);
//                    CTA q adds the four partials, does the gate-derivative math of ITS 32 units, publishes
//                    dG_t (scaled fp16 + flag) for the next step's gathers and writes dG (fp32 + fp16) for the
//                    weight-gradient GEMMs.
// The cell-gradient carry lives in registers; the factors of the derivatives that do not depend on dh are
// computed while the exchange is in flight.
#include "tc05.cuh"
#include "lstm_bwd.cuh"

namespace onssen {
namespace {

using namespace tc05;

#ifndef BT_WORKERS_N
#define BT_WORKERS_N 8    // worker warps.  16 (one column per finalising thread) helped the forward kernel but not this
#endif                    // one: 2.93 -> 3.05 us/step at cfg2 (the gather's last warp arrives later, 96 registers spill)
constexpr int BT_WORKERS = BT_WORKERS_N;
constexpr int BT_THREADS = (BT_WORKERS + 1) * 32;   // + 1 MMA warp
constexpr int BT_A_COL = 128;                       // first TMEM column of the resident slab
constexpr int BT_NACC = 2;                          // independent accumulators (see lstm_rec.cu)
constexpr int BT_TRACE_S0 = 100;
constexpr size_t BT_MIN_SMEM = 120 * 1024;          // > half of an SM's shared memory: one CTA per SM (512 TMEM columns each)

long long* g_tc_trace = nullptr;
int g_tc_sm_reserve = 0;   // SMs left free for concurrent kernels (NCCL all-reduce of the layer above), see bwd_tc_set_sm_reserve

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes into another CTA's shared memory; the receiver's mbarrier counts the bytes when they have landed
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, float a, float b, float c, float d, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   dst_cluster),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
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
__device__ __forceinline__ float hash_uniform32(unsigned int seed_lo, unsigned int seed_hi, unsigned int idx) {
  unsigned int x = idx ^ seed_lo;
  x *= 0x9E3779B1u; x ^= x >> 15;
  x *= 0x85EBCA77u; x ^= x >> 13;
  x += seed_hi;
  x *= 0xC2B2AE3Du; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ __half to_half_flag_range(float x) {   // clamp to the largest fp16 below 2.0
  x = fminf(fmaxf(x, -1.9990234375f), 1.9990234375f);
  return __float2half_rn(x);
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
  if constexpr (N == 4) tmem_ld4(taddr, v);
  if constexpr (N == 8) tmem_ld8(taddr, v);
  if constexpr (N == 16) tmem_ld16(taddr, v);
}

#define BT_TRACE(slot)                                                                    \
  do {                                                                                    \
    if (p.trace != nullptr && blockIdx.x == 0 && s >= BT_TRACE_S0 && s < BT_TRACE_S0 + 4) \
      p.trace[(s - BT_TRACE_S0) * 16 + (slot)] = clock64();                               \
  } while (0)

template <int NBP>
__global__ void __launch_bounds__(BT_THREADS, 1) lstm_bwd_tc_kernel(const BwdParams p, const int S, const int Bs) {
  constexpr int ITEMS = NBP / BT_WORKERS;          // batch columns per finalising thread (1, 2 or 4)
  constexpr int NBH = NBP / (BT_WORKERS / 4);      // accumulator columns per worker warp (4 warps cover the 128 lanes)
  constexpr int PITCH = NBP + 4;     // floats per (source CTA, unit) row of the receive tile (16-byte rows, few bank conflicts)
  constexpr int GT = BT_WORKERS * 32;
  constexpr int RX_PAR = 4 * 32 * PITCH;
  extern __shared__ __align__(128) uint8_t smem[];
  const int Hp = p.Hp, B = p.B, T = p.T;
  const int G4 = 4 * Hp;
  const uint32_t bt_bytes = (uint32_t)Hp * NBP * 2u;
  uint8_t* bt = smem;                                          // B operand: [Hp/8][NBP][8] fp16
  float* rx = reinterpret_cast<float*>(smem + bt_bytes);       // [parity][source CTA][unit][PITCH]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rx + 2 * RX_PAR);
  uint64_t* mma_bar = bars;
  uint64_t* rx_bar = bars + 1;                                 // [parity]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();                     // K quarter held here == unit group finalised here
  const int cid = blockIdx.x >> 2;
  const int nub = bwd_tc_nub(Hp);
  const int ub = cid % nub;
  const int sl = (cid / nub) % S;
  const int dir = cid / (nub * S);
  const int b0 = sl * Bs;
  const int nb_valid = min(Bs, B - b0);
  const int ngroups = (nb_valid + 3) >> 2;                     // 4-column groups that carry real utterances
  const bool my_has = ub * 128 + 32 * (int)rank < Hp;          // the last unit block may be partly padding
  const uint32_t rx_expect = 4u * 32u * 16u * (uint32_t)ngroups;
  const size_t xt_bytes = (size_t)G4 * NBP * 2;
  uint8_t* xt0 = p.xbuf + ((size_t)(0 * 2 + dir) * S + sl) * xt_bytes;
  uint8_t* xt1 = p.xbuf + ((size_t)(1 * 2 + dir) * S + sl) * xt_bytes;
  const size_t seg_off = (size_t)rank * bt_bytes;              // this CTA's K quarter inside an exchange tile

  if (warp == BT_WORKERS) {
    if (lane == 0) {
      mbar_init(mma_bar, 1);
      mbar_init(rx_bar, 1);
      mbar_init(rx_bar + 1, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 512);
  }
  for (uint32_t i = tid; i < bt_bytes / 16; i += BT_THREADS) reinterpret_cast<uint4*>(bt)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  if (tid == 0 && my_has) {   // both parities armed for their first use
    mbar_arrive_expect_tx(rx_bar, rx_expect);
    mbar_arrive_expect_tx(rx_bar + 1, rx_expect);
  }
  // resident slab: thread = unit row; its Hp fp16 (this K quarter's gate rows) go to TMEM lane r
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(
        p.wslab + ((((size_t)dir * nub + ub) * 4 + rank) * 128 + r) * (size_t)Hp);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + BT_A_COL;
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
  cluster_sync_all();   // every peer's mbarriers are initialised and armed before anything is sent to them

  if (warp == BT_WORKERS) {
    // ===================== MMA warp (converged; the elected lane issues) =====================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, NBP);
    const uint32_t bt_addr = smem_u32(bt);
    const int ksteps = Hp / 16;
    constexpr uint64_t DB_STEP = (2 * NBP * 16) >> 4;
    for (int s = 1; s < T; ++s) {
      named_bar_sync(3, BT_THREADS);   // B tile of step s written + fenced, accumulators drained
      tc_fence_after_sync();
      if (lane == 0) BT_TRACE(3);
      uint64_t db0 = make_smem_desc(bt_addr, NBP * 16, 128, 0);
      uint32_t ta0 = tmem_base + BT_A_COL;
      constexpr int UB = 19;
      for (int ks0 = 0; ks0 < ksteps; ks0 += UB) {
#pragma unroll
        for (int j = 0; j < UB; ++j) {
          if (ks0 + j < ksteps && leader)
            umma_f16_ts(tmem_base + (j % BT_NACC) * NBP, ta0 + 8 * j, db0 + (uint64_t)j * DB_STEP, idesc,
                        (ks0 + j) >= BT_NACC);
        }
        ta0 += 8 * UB;
        db0 += UB * DB_STEP;
      }
      if (leader) umma_commit(mma_bar);
      __syncwarp();
      if (lane == 0) BT_TRACE(4);
    }
  } else {
    // ===================== worker warps =====================
    const float inv_scale = p.scale2[1], scale = p.scale2[0];
    const float keep_scale = p.dropout_p > 0.f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
    // TMEM role: lane quarter q (= destination CTA of the reduce-scatter), column group `half` of NBH columns
    const int q = warp & 3, half = warp >> 2;
    const bool dst_has = ub * 128 + 32 * q < Hp;
    // finalise role: unit ul of this CTA's 32, columns cg*ITEMS .. +ITEMS
    const int ul = lane, cg = warp;
    const int u = min(ub * 128 + 32 * (int)rank + ul, Hp - 1);   // clamped when this CTA finalises nothing
    unsigned int valid_items = 0;
    int bcol[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int col = cg * ITEMS + i;
      if (my_has && col < nb_valid) valid_items |= 1u << i;
      bcol[i] = b0 + min(col, nb_valid - 1);
    }
    // gather bookkeeping: 16-byte chunk c = kc*NBP + n of the K quarter, thread handles c = tid + GT*i
    const int nchunks = Hp * NBP / 8;
    constexpr int MAXCH = (768 * NBP / 8 + GT - 1) / GT;   // Hp <= 768
    const int my_chunks = (nchunks - tid + GT - 1) / GT;
    unsigned int want_mask = 0;
    for (int i = 0; i < my_chunks; ++i)
      if (((tid + GT * i) % NBP) < nb_valid) want_mask |= 1u << i;
    // receive-tile addresses
    const uint32_t rx_send_local = smem_u32(rx + ((int)rank * 32 + lane) * PITCH + half * NBH);
    const uint32_t rx_send0 = mapa_u32(rx_send_local, (uint32_t)q);
    const uint32_t rx_bar0 = mapa_u32(smem_u32(rx_bar), (uint32_t)q);
    const float* rx_read = rx + ul * PITCH + cg * ITEMS;
    float dc_carry[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) dc_carry[i] = 0.f;

    for (int s = 0; s < T; ++s) {
      const int t = dir == 0 ? T - 1 - s : s;
      const int t_fp = dir == 0 ? t - 1 : t + 1;
      if (tid == 0) BT_TRACE(0);
      // ---- this step's saved state: loads issued before the exchange wait
      float4 a4[ITEMS];
      float ct[ITEMS], cprev[ITEMS], dyv[ITEMS];
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const long long m = (long long)t * B + bcol[i];
        const long long oy = m * (2 * Hp) + dir * Hp + u;
        a4[i] = *reinterpret_cast<const float4*>(p.actg + m * (2 * G4) + dir * G4 + 4 * u);
        ct[i] = p.c[oy];
        cprev[i] = (t_fp >= 0 && t_fp < T) ? p.c[((long long)t_fp * B + bcol[i]) * (2 * Hp) + dir * Hp + u] : 0.f;
        dyv[i] = p.dy[oy];
        if (p.dropout_p > 0.f) {
          const float rnd = hash_uniform32(p.seed_lo, p.seed_hi, (unsigned int)oy);
          dyv[i] *= rnd < p.dropout_p ? 0.f : keep_scale;
        }
      }
      float rsum[ITEMS];
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) rsum[i] = 0.f;
      if (s > 0) {
        // ---- gather the K quarter of dG_{previous step}: spin on the flag bits
        const uint8_t* src = ((s - 1) & 1 ? xt1 : xt0) + seg_off;
        const unsigned int fbit = ((((unsigned int)(s - 1)) >> 1) & 1u) ^ 1u;
        const unsigned int fword = fbit ? 0x40004000u : 0u;
        uint4 v[MAXCH];
        unsigned int pending = want_mask;
        while (pending) {
#pragma unroll
          for (int i = 0; i < MAXCH; ++i)
            if (pending & (1u << i)) v[i] = ld_relaxed_v4(src + (size_t)(tid + GT * i) * 16);
#pragma unroll
          for (int i = 0; i < MAXCH; ++i)
            if (pending & (1u << i)) {
              const unsigned int m = 0x40004000u;
              const bool fresh = ((v[i].x & m) == fword) && ((v[i].y & m) == fword) && ((v[i].z & m) == fword) &&
                                 ((v[i].w & m) == fword);
              if (fresh) {
                *reinterpret_cast<uint4*>(bt + (size_t)(tid + GT * i) * 16) =
                    make_uint4(v[i].x & ~m, v[i].y & ~m, v[i].z & ~m, v[i].w & ~m);
                pending &= ~(1u << i);
              }
            }
        }
        if (tid == 0) BT_TRACE(1);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        asm volatile("bar.arrive 3, %0;" ::"r"(BT_THREADS) : "memory");
        if (tid == 0) BT_TRACE(2);
      }
      // ---- factors of the gate derivatives that do not depend on dh (overlap the MMA phase)
      float k_c[ITEMS], k_i[ITEMS], k_f[ITEMS], k_g[ITEMS], k_o[ITEMS], f_g[ITEMS];
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
      if (s > 0) {
        const int par = s & 1;
        // ---- partial dh of this K quarter: TMEM -> registers -> the finalising CTA's receive tile
        mbar_wait(mma_bar, (s - 1) & 1);
        tc_fence_after_sync();
        if (tid == 0) BT_TRACE(5);
        uint32_t acc[BT_NACC][NBH];
#pragma unroll
        for (int a = 0; a < BT_NACC; ++a)
          tmem_ld_cols<NBH>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * NBP + half * NBH, acc[a]);
        tmem_wait_ld();
        if (dst_has) {
          const uint32_t dst = rx_send0 + (uint32_t)par * RX_PAR * 4u;
          const uint32_t dbar = rx_bar0 + (uint32_t)par * 8u;
#pragma unroll
          for (int g = 0; g < NBH / 4; ++g) {
            if (half * NBH + 4 * g < nb_valid) {
              float f[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = __uint_as_float(acc[0][4 * g + e]);
#pragma unroll
                for (int q2 = 1; q2 < BT_NACC; ++q2) a += __uint_as_float(acc[q2][4 * g + e]);
                f[e] = a;
              }
              st_async_v4(dst + 16u * g, f[0], f[1], f[2], f[3], dbar);
            }
          }
        }
        if (tid == 0) BT_TRACE(6);
        // ---- reduce: the four partials of this CTA's 32 units
        if (my_has) {
          mbar_wait_cluster(rx_bar + par, ((unsigned int)(s - 1) >> 1) & 1u);
          if (tid == 0) {
            mbar_arrive_expect_tx(rx_bar + par, rx_expect);   // next use: step s+2
            BT_TRACE(7);
          }
          const float* rp = rx_read + par * RX_PAR;
#pragma unroll
          for (int src_cta = 0; src_cta < 4; ++src_cta) {
            if constexpr (ITEMS == 1) {
              rsum[0] += rp[src_cta * 32 * PITCH];
            } else if constexpr (ITEMS == 2) {
              const float2 w2 = *reinterpret_cast<const float2*>(rp + src_cta * 32 * PITCH);
              rsum[0] += w2.x; rsum[1] += w2.y;
            } else {
              const float4 w4 = *reinterpret_cast<const float4*>(rp + src_cta * 32 * PITCH);
              rsum[0] += w4.x; rsum[1] += w4.y; rsum[ITEMS - 2] += w4.z; rsum[ITEMS - 1] += w4.w;
            }
          }
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) rsum[i] *= inv_scale;
        }
      }
      // ---- gate derivatives of this CTA's units at step t; publish dG for the next step's gathers
      const unsigned int fbit_w = ((((unsigned int)s) >> 1) & 1u) ^ 1u;
      const unsigned int fww = fbit_w ? 0x40004000u : 0u;
      uint8_t* xt_w = (s & 1) ? xt1 : xt0;
      float4 d4v[ITEMS];
      uint2 ov[ITEMS];
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const float dh = dyv[i] + rsum[i];
        const float dct = dh * k_c[i] + dc_carry[i];
        dc_carry[i] = dct * f_g[i];
        const float s_i = dct * k_i[i], s_f = dct * k_f[i], s_g = dct * k_g[i], s_o = dh * k_o[i];
        __half2 lo = __halves2half2(to_half_flag_range(s_i), to_half_flag_range(s_f));
        __half2 hi = __halves2half2(to_half_flag_range(s_g), to_half_flag_range(s_o));
        if ((valid_items & (1u << i)) && p.sat != nullptr &&
            !(fmaxf(fmaxf(fabsf(s_i), fabsf(s_f)), fmaxf(fabsf(s_g), fabsf(s_o))) < 1.9995f))
          atomicAdd(p.sat, 1u);
        ov[i].x = *reinterpret_cast<uint32_t*>(&lo);
        ov[i].y = *reinterpret_cast<uint32_t*>(&hi);
        d4v[i] = make_float4(s_i * inv_scale, s_f * inv_scale, s_g * inv_scale, s_o * inv_scale);
      }
      // rows 4u..4u+3 of column n are half of the 16-byte chunk (kc = u/2, n): lanes (ul, ul^1) swap one column
      // each so that every lane issues ONE 16-byte store per column pair
      if constexpr (ITEMS == 1) {
        // one column per thread: the even lane of a unit pair stores the pair's chunk
        uint2 got;
        got.x = __shfl_xor_sync(0xffffffffu, ov[0].x, 1);
        got.y = __shfl_xor_sync(0xffffffffu, ov[0].y, 1);
        if (s + 1 < T && (valid_items & 1u) && !(lane & 1))
          st_relaxed_v4(xt_w + ((size_t)(u >> 1) * NBP + cg) * 16,
                        make_uint4(ov[0].x | fww, ov[0].y | fww, got.x | fww, got.y | fww));
      } else {
#pragma unroll
        for (int i = 0; i < ITEMS; i += 2) {
          const uint2 mine_keep = (lane & 1) ? ov[i + 1] : ov[i];
          const uint2 mine_send = (lane & 1) ? ov[i] : ov[i + 1];
          uint2 got;
          got.x = __shfl_xor_sync(0xffffffffu, mine_send.x, 1);
          got.y = __shfl_xor_sync(0xffffffffu, mine_send.y, 1);
          const int ii = i + (lane & 1);
          if (s + 1 < T && (valid_items & (1u << ii))) {
            const int col = cg * ITEMS + ii;
            const uint4 chunk = (lane & 1) ? make_uint4(got.x | fww, got.y | fww, mine_keep.x | fww, mine_keep.y | fww)
                                           : make_uint4(mine_keep.x | fww, mine_keep.y | fww, got.x | fww, got.y | fww);
            st_relaxed_v4(xt_w + ((size_t)(u >> 1) * NBP + col) * 16, chunk);
          }
        }
      }
      if (tid == 0) BT_TRACE(8);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (valid_items & (1u << i)) {
          const long long off = ((long long)t * B + bcol[i]) * (2 * G4) + dir * G4 + 4 * u;
          *reinterpret_cast<float4*>(p.actg + off) = d4v[i];
          *reinterpret_cast<uint2*>(p.dg16 + off) = ov[i];
        }
      }
      if (tid == 0) BT_TRACE(9);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // nobody exits while a peer may still write into its receive tile
  if (warp == BT_WORKERS) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NBP>
size_t bt_smem_bytes(int Hp) {
  const size_t need = (size_t)Hp * NBP * 2 + (size_t)2 * 4 * 32 * (NBP + 4) * 4 + 64;
  return need > BT_MIN_SMEM ? need : BT_MIN_SMEM;
}

template <int NBP>
int max_clusters(int Hp) {
  auto kern = lstm_bwd_tc_kernel<NBP>;
  const size_t smem = bt_smem_bytes<NBP>(Hp);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4);
  cfg.blockDim = dim3(BT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

template <int NBP>
int launch_tc(const BwdParams& p, int S, int Bs, cudaStream_t stream) {
  auto kern = lstm_bwd_tc_kernel<NBP>;
  const size_t smem = bt_smem_bytes<NBP>(p.Hp);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return ONSSEN_ERR_CUDA;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * 2 * S * bwd_tc_nub(p.Hp));
  cfg.blockDim = dim3(BT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the exchange spins on its peers
  at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  if (cudaLaunchKernelEx(&cfg, kern, p, S, Bs) != cudaSuccess) return ONSSEN_ERR_CUDA;
  return ONSSEN_OK;
}

}  // namespace

void bwd_tc_set_trace(long long* buf) { g_tc_trace = buf; }
void bwd_tc_set_sm_reserve(int sms) { g_tc_sm_reserve = sms < 0 ? 0 : sms; }

namespace {
struct TcPlan { int nbp, S, Bs; };
// slice plan shared by scratch sizing and launch: (direction, slice) groups of nub clusters each; as many slices as
// fit (fewer columns per CTA = shorter step), 16 columns per MMA when that covers the batch, else 32
bool plan_tc(int B, int Hp, TcPlan& pl) {
  if (B <= 0 || Hp / 2 + BT_A_COL > 512 || Hp > 768) return false;
  const int nub = bwd_tc_nub(Hp);
  static int maxc16 = -1, maxc32 = -1, maxc_hp = -1;
  if (maxc_hp != Hp) {
    maxc16 = max_clusters<16>(Hp);
    maxc32 = max_clusters<32>(Hp);
    maxc_hp = Hp;
  }
  for (int nbp = 16; nbp <= 32; nbp += 16) {
    int maxc = nbp == 16 ? maxc16 : maxc32;
    const int avail = (num_sms() - g_tc_sm_reserve) / 4;
    if (maxc > avail) maxc = avail;
    const int smax = maxc / (2 * nub);
    if (smax < 1 || (B + nbp - 1) / nbp > smax) continue;
    int S = smax < B ? smax : B;
    const int Bs = (B + S - 1) / S;
    S = (B + Bs - 1) / Bs;
    pl.nbp = nbp; pl.S = S; pl.Bs = Bs;
    return true;
  }
  return false;
}
}  // namespace

size_t bwd_tc_xbuf_bytes(int B, int Hp) {
  TcPlan pl;
  if (!plan_tc(B, Hp, pl)) return 0;
  return (size_t)2 * 2 * pl.S * 4 * Hp * pl.nbp * 2;
}

int launch_bwd_tc(const BwdParams& p_in, cudaStream_t stream) {
  BwdParams p = p_in;
  TcPlan pl;
  if (!p.wslab || !p.xbuf || !plan_tc(p.B, p.Hp, pl)) return ONSSEN_ERR_UNSUPPORTED;
  p.trace = g_tc_trace;
  return pl.nbp == 16 ? launch_tc<16>(p, pl.S, pl.Bs, stream) : launch_tc<32>(p, pl.S, pl.Bs, stream);
}

}  // namespace onssen
