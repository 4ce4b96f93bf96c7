// Dense projection GEMM for the BLSTM path on sm_100a:
//   out[m][n] = epilogue( sum_k A[m][k] * W[n][k] + bias[n] )
// A (activations) and W (weights) are fp16, K-major, fed by TMA (128B swizzle) into a 4-stage
// shared-memory ring; tcgen05.mma (kind::f16, M=128, N<=256, K=16) accumulates fp32 in TMEM with two
// accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1. Persistent: one CTA / SM.
//
// Used for (reference call sites it replaces):
//   * LSTM input projections W_ih x_t + b_ih + b_hh for all t at once   (torch nn.LSTM inside
//     onssen/nn/deep_clustering.py:34-35, chimera.py:35-36, enhancement.py:43-44)
//   * embedding head Linear + view + F.normalize                         (deep_clustering.py:39-42,
//     chimera.py:37,39-41)   -> EPI_L2NORM epilogue, rows remapped from time-major to (B,T)
//   * mask head Linear + sigmoid                                         (chimera.py:38,42)
//   * enhancement restoration layers Linear + relu                       (enhancement.py:49,51)
#include "tc05.cuh"
#include "common.cuh"
#include <cudaTypedefs.h>
#include <cstdlib>

namespace {

using namespace tc05;

constexpr int BM = 128;
constexpr int BK = 64;        // fp16 elements = 128 B = one swizzle row
constexpr int MAX_BN = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;       // 16 KB
constexpr int B_STAGE_BYTES = MAX_BN * BK * 2;   // 32 KB
constexpr int EPI_PITCH_MAX = 44;                // floats per staged row (D<=40 + 4 pad, or 32+4)
constexpr int EPI_WARP_FLOATS = 32 * EPI_PITCH_MAX;
// Two shapes of the same kernel, chosen by the host from K:
//   EW = 4 epilogue warps, 4 stages of 48 KB: long-K products (weight gradients, dgrad) whose tile time is MMA / L2 time;
//   EW = 8 epilogue warps (two per TMEM lane quarter, alternating 32-column chunks / D-groups), 3 stages: K <= 2048
//          (forward projections, heads).  With one epilogue warp per scheduler every dependent instruction waited out
//          its full latency (ncu: 17 % of the issue slots used, stall reason 'wait') and a 128x256 tile took ~19 k
//          cycles to drain -- longer than its MMAs at K = 1216.
__host__ __device__ constexpr int gemm_stages(int ew) { return ew == 8 ? 3 : 4; }
__host__ __device__ constexpr int gemm_threads(int ew) { return (2 + ew) * 32; }   // warp0 TMA, warp1 MMA+TMEM, epilogue
__host__ __device__ constexpr size_t gemm_smem_bytes(int ew) {
  return 1024 /*align slack*/ + (size_t)gemm_stages(ew) * (A_STAGE_BYTES + B_STAGE_BYTES) + (size_t)ew * EPI_WARP_FLOATS * 4 +
         (size_t)ew * MAX_BN * 4 /*per-warp bias tile*/ + 256;
}
constexpr int TMEM_COLS = 512;

struct GemmParams {
  int M, N, K;
  const float* bias;
  float* out;
  long long ld_out;
  int epi;          // 0 plain, 1 sigmoid, 2 relu, 3 l2norm over groups of `group` columns
  int group;
  int remap_inner;  // if > 0: row m = o * remap_inner + i is written to output row i * remap_outer + o
  int remap_outer;
  int block_n;
  int num_m_blocks, num_n_blocks;
  const float* out_scale;  // optional device scalar: accumulators are multiplied by *out_scale (before bias)
  float* inv_norm;         // optional (epi 3): 1/max(||z||,eps) per (output row, group) -> [rows][N/group]
  // "rows" mode (weight gradients): both operands are stored [k][m] / [k][n] (the contraction runs over their ROWS),
  // read in place as MN-major UMMA operands: TMA boxes of 64 k-rows x 64 columns (128 B, 128B swizzle), 8 KB each.
  int mn_major;
  int b_row_shift;         // rows mode: B row k + b_row_shift pairs with A row k (time shift of the W_hh gradient)
};

__device__ __forceinline__ float act_apply(float x, int epi) {
  if (epi == 1) return 1.0f / (1.0f + __expf(-x));
  if (epi == 2) return fmaxf(x, 0.0f);
  return x;
}

template <int D>
__device__ __forceinline__ void tmem_ld_group(uint32_t taddr, uint32_t* v) {
  // binary decomposition of D (multiple of 4, <= 40) into power-of-two TMEM loads
  int o = 0;
  if constexpr (D >= 32) { tmem_ld32(taddr + o, v + o); o += 32; }
  if constexpr ((D % 32) >= 16) { tmem_ld16(taddr + o, v + o); o += 16; }
  if constexpr ((D % 16) >= 8) { tmem_ld8(taddr + o, v + o); o += 8; }
  if constexpr ((D % 8) >= 4) { tmem_ld4(taddr + o, v + o); o += 4; }
}

template <int D, int EPI_WARPS>  // D = 0: no l2norm path compiled in
__global__ void __launch_bounds__(gemm_threads(EPI_WARPS), 1)
gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const GemmParams p) {
  constexpr int STAGES = gemm_stages(EPI_WARPS);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  float* bias_stage = epi_stage + EPI_WARPS * EPI_WARP_FLOATS;   // [warp][MAX_BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_stage + EPI_WARPS * MAX_BN);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BN = p.block_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], EPI_WARPS);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, TMEM_COLS);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = A_STAGE_BYTES + (uint32_t)BN * BK * 2;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.num_m_blocks;
        const int n_blk = tile / p.num_m_blocks;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.mn_major) {
            const int nbox_b = (BN + 63) / 64;
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(2 + nbox_b) * 8192u);
            for (int j = 0; j < 2; ++j)
              tma_load_2d(smem_a + stage * A_STAGE_BYTES + j * 8192, &tmap_a, &full_bar[stage], m_blk * BM + 64 * j,
                          kb * BK);
            for (int j = 0; j < nbox_b; ++j)   // rows outside [0, K) (time shift) and columns >= N are zero-filled
              tma_load_2d(smem_b + stage * B_STAGE_BYTES + j * 8192, &tmap_w, &full_bar[stage], n_blk * BN + 64 * j,
                          kb * BK + p.b_row_shift);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_2d(smem_a + stage * A_STAGE_BYTES, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmap_w, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // rows mode: a_major = b_major = MN (instruction-descriptor bits 15 / 16)
      const uint32_t idesc = make_idesc_f16(BM, BN) | (p.mn_major ? ((1u << 15) | (1u << 16)) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + as * MAX_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
          if (p.mn_major) {
            // MN-major, 128B swizzle: 64 columns x 8 k-rows per 1 KB atom; LBO = next 64-column box (8 KB),
            // SBO = next 8 k-rows (1 KB); one MMA (K = 16) spans two atoms
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * 2048, 8192, 1024, 2);
              const uint64_t db = make_smem_desc(b_addr + k * 2048, 8192, 1024, 2);
              umma_f16(tmem_d, da, db, idesc, (kb | k) != 0);
            }
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024, 2);
              const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 1024, 2);
              umma_f16(tmem_d, da, db, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
  } else {
    // ===================== epilogue (EPI_WARPS warps) =====================
    constexpr int NSEL = EPI_WARPS / 4;           // warps per TMEM lane quarter
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int hsel = (warp - 2) >> 2;             // which of the NSEL warps of that quarter: chunks hsel, hsel+NSEL, ..
    float* stg = epi_stage + (warp - 2) * EPI_WARP_FLOATS;
    float* bias_s = bias_stage + (warp - 2) * MAX_BN;
    const bool vec_ok = ((p.ld_out & 3) == 0) && ((p.N & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_blk = tile % p.num_m_blocks;
      const int n_blk = tile / p.num_m_blocks;
      const int m0 = m_blk * BM + q * 32;
      const int n0 = n_blk * BN;
      // this tile's bias slice -> warp-private smem (overlaps the wait for the accumulator)
      for (int i = lane; i < BN; i += 32)
        bias_s[i] = (p.bias != nullptr && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
      __syncwarp();
      const float oscale = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.0f;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * MAX_BN;

      // output row offset (elements) of the row this lane owns while staging rows are written out
      auto out_row_off = [&](int m) -> long long {
        long long r = m;
        if (p.remap_inner > 0) {
          const int o = m / p.remap_inner;
          const int i = m - o * p.remap_inner;
          r = (long long)i * p.remap_outer + o;
        }
        return r * p.ld_out;
      };

      if (D > 0 && p.epi == 3) {
        if constexpr (D > 0) {
          constexpr int PITCH = D + 4;
          const int ngroups = BN / D;
          for (int g = hsel; g < ngroups; g += NSEL) {
            const int nb = n0 + g * D;
            if (nb >= p.N) break;   // warp-uniform
            uint32_t v[D];
            tmem_ld_group<D>(taddr + g * D, v);
            tmem_wait_ld();
            float ss = 0.f;
            float f[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
              f[j] = fmaf(__uint_as_float(v[j]), oscale, bias_s[g * D + j]);
              ss = fmaf(f[j], f[j], ss);
            }
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            if (p.inv_norm != nullptr && m0 + lane < p.M)
              p.inv_norm[out_row_off(m0 + lane) / p.ld_out * (p.N / D) + (nb / D)] = inv;
#pragma unroll
            for (int j = 0; j < D; j += 4) {
              float4 o4 = make_float4(f[j] * inv, f[j + 1] * inv, f[j + 2] * inv, f[j + 3] * inv);
              *reinterpret_cast<float4*>(stg + lane * PITCH + j) = o4;
            }
            __syncwarp();
            constexpr int V4_PER_ROW = D / 4;
            for (int idx = lane; idx < 32 * V4_PER_ROW; idx += 32) {
              const int r = idx / V4_PER_ROW;
              const int c4 = idx - r * V4_PER_ROW;
              const int m = m0 + r;
              if (m < p.M) {
                const float4 o4 = *reinterpret_cast<const float4*>(stg + r * PITCH + c4 * 4);
                *reinterpret_cast<float4*>(p.out + out_row_off(m) + nb + c4 * 4) = o4;
              }
            }
            __syncwarp();
          }
        }
      } else {
        constexpr int PITCH = 36;
        // this lane writes column group c4 = lane & 7 of rows 4*i + lane/8: output row offsets once per tile
        long long roff[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + 4 * i + (lane >> 3);
          roff[i] = m < p.M ? out_row_off(m) : -1;
        }
        for (int c0 = 32 * hsel; c0 < BN; c0 += 32 * NSEL) {
          const int nb = n0 + c0;
          if (nb >= p.N) break;   // warp-uniform
          uint32_t v[32];
          if (BN - c0 >= 32) {
            tmem_ld32(taddr + c0, v);
          } else {            // BN multiple of 16: 16-column tail
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int j = 16; j < 32; ++j) v[j] = 0;
          }
          tmem_wait_ld();
          auto stage_chunk = [&](auto act) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + j);   // smem broadcast
              float4 o4;
              o4.x = act(fmaf(__uint_as_float(v[j]), oscale, b4.x));
              o4.y = act(fmaf(__uint_as_float(v[j + 1]), oscale, b4.y));
              o4.z = act(fmaf(__uint_as_float(v[j + 2]), oscale, b4.z));
              o4.w = act(fmaf(__uint_as_float(v[j + 3]), oscale, b4.w));
              *reinterpret_cast<float4*>(stg + lane * PITCH + j) = o4;
            }
          };
          if (p.epi == 1) stage_chunk([](float x) { return 1.0f / (1.0f + __expf(-x)); });
          else if (p.epi == 2) stage_chunk([](float x) { return fmaxf(x, 0.0f); });
          else stage_chunk([](float x) { return x; });
          __syncwarp();
          const int ncols = min(32, min(BN - c0, p.N - nb));
          if (vec_ok) {
            const int c4 = lane & 7;
            if (c4 * 4 < ncols) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (roff[i] >= 0) {
                  const float4 o4 = *reinterpret_cast<const float4*>(stg + (4 * i + (lane >> 3)) * PITCH + c4 * 4);
                  *reinterpret_cast<float4*>(p.out + roff[i] + nb + c4 * 4) = o4;
                }
              }
            }
          } else {
            for (int r = 0; r < 32; ++r) {
              const int m = m0 + r;
              if (m < p.M && lane < ncols) p.out[out_row_off(m) + nb + lane] = stg[r * PITCH + lane];
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 2-D fp16 K-major tensor map: dims {K, rows}, box {64, box_rows}, 128B swizzle.
int make_tmap_f16(CUtensorMap* map, const void* base, int rows, int K, long long ld_elems, int box_rows) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return ONSSEN_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? ONSSEN_OK : ONSSEN_ERR_DRIVER;
}

// rows mode: source [rows = contraction][cols] fp16 row-major; dims {cols, rows}, box {64 columns, 64 rows}
int make_tmap_f16_rows(CUtensorMap* map, const void* base, int rows, int cols, long long ld_elems) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return ONSSEN_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? ONSSEN_OK : ONSSEN_ERR_DRIVER;
}

int pick_block_n(int N, int epi, int group) {
  if (epi == 3) {
    // largest multiple of lcm(group,16) that is <= 256
    int l = group;
    while (l % 16 != 0) l += group;
    if (l > 256) return -1;
    return (256 / l) * l;
  }
  if (N >= 256) return 256;
  return ((N + 15) / 16) * 16;
}

// Tile width for a problem of m_blocks x N: 256 columns unless that leaves SMs idle in the last (or only) wave -- the
// W_hh weight gradients (2432 x 608 outputs, K = 12800) are 57 tiles of 128x256 on 148 SMs.  Cost model: waves x (BN + 64)
// (the 128-row A tile is loaded per k-block whatever BN is).
int pick_block_n_fill(int m_blocks, int N) {
  const int sms = onssen::num_sms();
  const int widest = pick_block_n(N, 0, 0);
  static const bool enabled = []() { const char* e = getenv("ONSSEN_GEMM_FILL"); return !(e && e[0] == '0'); }();
  if (!enabled) return widest;      // ONSSEN_GEMM_FILL=0: always the widest tile (A/B switch)
  int best = widest;
  long long best_cost = -1;
  for (int bn : {widest, 192, 128, 96, 64}) {
    if (bn > widest) continue;
    const long long tiles = (long long)m_blocks * ((N + bn - 1) / bn);
    const long long cost = ((tiles + sms - 1) / sms) * (bn + 64);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

template <int D, int EW>
int launch_ew(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tc05_kernel<D, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)gemm_smem_bytes(EW)) != cudaSuccess)
      return ONSSEN_ERR_CUDA;
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < onssen::num_sms() ? tiles : onssen::num_sms();
  gemm_tc05_kernel<D, EW><<<grid, gemm_threads(EW), gemm_smem_bytes(EW), stream>>>(ta, tw, p);
  return cudaGetLastError() == cudaSuccess ? ONSSEN_OK : ONSSEN_ERR_CUDA;
}

template <int D>
int launch(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, cudaStream_t stream) {
  // short K: the tile drains slower than it is computed -> 8 epilogue warps; long K: 4 stages of operands matter more
  return p.K <= 2048 ? launch_ew<D, 8>(ta, tw, p, stream) : launch_ew<D, 4>(ta, tw, p, stream);
}

}  // namespace

namespace onssen {

// A: [M][lda] fp16, W: [N][ldw] fp16 (both K-major, K multiple of 8, lda/ldw multiples of 8),
// out: fp32. See GemmParams for epi/remap semantics.
int gemm_f16(const void* A, const void* W, const float* bias, float* out, int M, int N, int K, long long lda,
             long long ldw, long long ld_out, int epi, int group, int remap_inner, int remap_outer,
             cudaStream_t stream, const float* out_scale, float* inv_norm) {
  if (M <= 0 || N <= 0 || K <= 0) return ONSSEN_ERR_ARG;
  if ((lda & 7) || (ldw & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return ONSSEN_ERR_ARG;
  if (epi == 3 && (bias == nullptr || N % group != 0)) return ONSSEN_ERR_ARG;
  const int bn = pick_block_n(N, epi, group);
  if (bn <= 0) return ONSSEN_ERR_UNSUPPORTED;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.out = out; p.ld_out = ld_out; p.epi = epi; p.group = group;
  p.remap_inner = remap_inner; p.remap_outer = remap_outer; p.block_n = bn;
  p.out_scale = out_scale; p.inv_norm = inv_norm;
  p.mn_major = 0; p.b_row_shift = 0;
  p.num_m_blocks = (M + BM - 1) / BM;
  p.num_n_blocks = (N + bn - 1) / bn;
  CUtensorMap ta, tw;
  int rc = make_tmap_f16(&ta, A, M, K, lda, BM);
  if (rc != ONSSEN_OK) return rc;
  rc = make_tmap_f16(&tw, W, N, K, ldw, bn);
  if (rc != ONSSEN_OK) return rc;
  if (epi == 3) {
    switch (group) {
      case 4: return launch<4>(ta, tw, p, stream);
      case 8: return launch<8>(ta, tw, p, stream);
      case 12: return launch<12>(ta, tw, p, stream);
      case 16: return launch<16>(ta, tw, p, stream);
      case 20: return launch<20>(ta, tw, p, stream);
      case 24: return launch<24>(ta, tw, p, stream);
      case 32: return launch<32>(ta, tw, p, stream);
      case 40: return launch<40>(ta, tw, p, stream);
      default: return ONSSEN_ERR_UNSUPPORTED;
    }
  }
  return launch<0>(ta, tw, p, stream);
}

// out[m][n] = out_scale * sum_k X[k][m] * Y[k + y_row_shift][n]   (rows of Y outside [0, Kc) count as zero)
// X: [Kc][ldx] fp16, Y: [Kc][ldy] fp16, both row-major and read in place (MN-major UMMA operands): the weight
// gradients dW = dZ^T A contract over the T*B rows of two time-major buffers, no transposed copies.
int gemm_f16_rows(const void* X, const void* Y, float* out, int M, int N, int Kc, long long ldx, long long ldy,
                  long long ld_out, int y_row_shift, const float* out_scale, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || Kc <= 0) return ONSSEN_ERR_ARG;
  if ((ldx & 7) || (ldy & 7) || (reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Y) & 15))
    return ONSSEN_ERR_ARG;
  GemmParams p;
  p.M = M; p.N = N; p.K = Kc; p.bias = nullptr; p.out = out; p.ld_out = ld_out; p.epi = 0; p.group = 0;
  p.remap_inner = 0; p.remap_outer = 0; p.block_n = pick_block_n_fill((M + BM - 1) / BM, N);
  p.out_scale = out_scale; p.inv_norm = nullptr;
  p.mn_major = 1; p.b_row_shift = y_row_shift;
  p.num_m_blocks = (M + BM - 1) / BM;
  p.num_n_blocks = (N + p.block_n - 1) / p.block_n;
  CUtensorMap ta, tw;
  int rc = make_tmap_f16_rows(&ta, X, Kc, M, ldx);
  if (rc != ONSSEN_OK) return rc;
  rc = make_tmap_f16_rows(&tw, Y, Kc, N, ldy);
  if (rc != ONSSEN_OK) return rc;
  return launch<0>(ta, tw, p, stream);
}

bool gemm_l2norm_group_supported(int group) {
  return group == 4 || group == 8 || group == 12 || group == 16 || group == 20 || group == 24 || group == 32 ||
         group == 40;
}

}  // namespace onssen
