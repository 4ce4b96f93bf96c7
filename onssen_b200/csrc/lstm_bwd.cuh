// Shared between the two BPTT implementations (lstm_bwd.cu: per-step + persistent mma.sync kernels;
// lstm_bwd_tc.cu: persistent tcgen05 kernel, K split over a 4-CTA cluster).
#pragma once
#include "common.cuh"

namespace onssen {

struct BwdParams {
  float* actg;          // [T*B][2*4Hp] activated gates (i,f,g,o) -> overwritten with dG (fp32)
  __half* dg16;         // [T*B][2*4Hp] scaled fp16 dG
  const float* c;       // [T*B][2*Hp]
  const float* dy;      // [T*B][2*Hp] gradient w.r.t. the layer output (after dropout)
  const uint32_t* wt;   // W_hh^T in mma A-fragment order: [dir][ub][kstep][mtile][lane][4 words]
  const __half* wslab;  // W_hh^T as TMEM slabs [dir][unit block of 128][K quarter][128 units][Hp gate rows] (tcgen05 kernel)
  uint8_t* xbuf;        // tcgen05 kernel: dG exchange tiles [parity][dir][slice][4Hp/8][NBP][8] fp16 with flag bits
  uint32_t* frag;       // dG of the last processed step in B-fragment order: [parity][dir][bb][kstep][ntile][lane][2]
  float* dc;            // [2][B][Hp] cell-gradient carry
  const float* scale2;  // {scale, 1/scale}
  unsigned int* sat;    // optional: number of published values that had to be clamped into the flag range (or NaN)
  int B, T, H, Hp, s;
  float dropout_p;
  unsigned int seed_lo, seed_hi;
  long long* trace;     // debug: clock64 stamps of CTA (0,0,0), steps [BWD_TRACE_S0, +8): [step][slot 0..7][warp 0..7]
};


// elements (fp16) of the two W_hh^T layouts inside the buffer filled by onssen_lstm_pack_whh_t
inline size_t whh_t_frag_elems(int Hp) { return (size_t)2 * Hp * 4 * Hp; }
__host__ __device__ inline int bwd_tc_nub(int Hp) { return (Hp + 127) / 128; }
inline size_t whh_t_slab_elems(int Hp) { return (size_t)2 * bwd_tc_nub(Hp) * 4 * 128 * Hp; }
// bytes of the tcgen05 kernel's exchange area for this shape (0 when the shape is not supported by that kernel)
size_t bwd_tc_xbuf_bytes(int B, int Hp);

// Persistent tcgen05 BPTT (lstm_bwd_tc.cu). Returns ONSSEN_ERR_UNSUPPORTED when the shape does not fit one launch
// (the caller falls back to the mma.sync kernels).
int launch_bwd_tc(const BwdParams& p, cudaStream_t stream);
void bwd_tc_set_trace(long long* buf);
void bwd_tc_set_sm_reserve(int sms);

}  // namespace onssen
