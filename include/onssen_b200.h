/*
 * onssen_b200 -- C ABI of the B200 (sm_100a) STFT-mask separation hot path.
 *
 * The reference (speechLabBcCuny/onssen) is pure Python/PyTorch: it has no FFI of its own.  Every entry
 * point below therefore cites the reference *Python call site* whose arithmetic it replaces; the
 * reference-side binding (ctypes) is shown in INTEGRATION.md and implemented in onssen_b200/_lib.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - every function returns ONSSEN_OK (0) or a negative ONSSEN_ERR_* code; nothing throws, nothing
 *     allocates, no hidden global state (except per-process device attribute caches)
 *   - tensors are dense row-major; "time-major" means row index m = t * B + b
 *   - Hp = 32*ceil(H/32) is the hidden size padded to whole 32-unit row blocks; "packed" LSTM buffers
 *     use the layouts documented at onssen_lstm_pack_* 
 */
#ifndef ONSSEN_B200_H_
#define ONSSEN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONSSEN_OK 0
#define ONSSEN_ERR_ARG (-1)         /* bad argument (null, misaligned, size) */
#define ONSSEN_ERR_UNSUPPORTED (-2) /* shape outside what the kernels are built for */
#define ONSSEN_ERR_CUDA (-3)        /* CUDA runtime launch/config error (see cudaGetLastError) */
#define ONSSEN_ERR_DRIVER (-4)      /* cuTensorMapEncodeTiled unavailable / failed */
#define ONSSEN_ERR_RESIDENCY (-5)   /* persistent recurrent grid does not fit on the device */

/* label dtypes accepted by the loss kernels (reference labels are float64, feature_utils.py:86) */
#define ONSSEN_DT_F32 0
#define ONSSEN_DT_F64 1
#define ONSSEN_DT_U8 2

const char* onssen_version(void);
const char* onssen_error_string(int code);
/* number of SMs of the current device (cached) */
int onssen_num_sms(void);

/* ------------------------------------------------------------------------------------------------
 * STFT featurizer  (replaces onssen/data/feature_utils.py:5-21,49-95 + the crop/tile logic of
 * onssen/data/wsj0_2mix.py:114-152, executed per batch on the device instead of per item on the host)
 *
 * wav_*      [B][nsample] float32 waveforms (mix, s1, s2); s1/s2 may be NULL when no label output needs them
 * nsample_per_utt  optional [B] int32 true lengths (each n_fft/2 < len <= nsample, rows zero padded): lets a
 *            batch hold utterances of different duration (frames, reflect padding and tiling per utterance)
 * crop_start [B] int32 first frame of the crop in the (possibly tiled) frame sequence
 *            (reference: np.random.randint(frames - T), wsj0_2mix.py:125) -- explicit so indexing is exact
 * n_fft in {64..2048, power of two}; frames = 1 + nsample/hop (center=True, reflect padding, periodic Hann);
 * when frames <= T the STFT is tiled (frame f -> f % frames), wsj0_2mix.py:118-123.
 * Outputs (any may be NULL): all float32
 *   feature  [B][T][F]   log10(|mix| + 1e-7)            feature_utils.py:49-51
 *   mag_mix, mag_s1, mag_s2 [B][T][F]                   wsj0_2mix.py:132-134
 *   cos_s1, cos_s2 [B][T][F]  cos(angle(mix)-angle(s))  feature_utils.py:67-80
 *   ph_mix, ph_s1, ph_s2 [B][T][F][2]  (re, im)         feature_utils.py:54-64
 *   feat_max [B]  per-utterance max of `feature` over the crop (input of the VAD threshold)
 * F = n_fft/2 + 1.
 */
int onssen_stft_features(const float* wav_mix, const float* wav_s1, const float* wav_s2, int B, int nsample,
                         int n_fft, int hop, const int32_t* crop_start, int T, float* feature, float* mag_mix,
                         float* mag_s1, float* mag_s2, float* cos_s1, float* cos_s2, float* ph_mix,
                         float* ph_s1, float* ph_s2, float* feat_max, const int32_t* nsample_per_utt,
                         void* stream);

/* Ideal-binary labels + VAD  (replaces get_one_hot, feature_utils.py:83-95).
 * one_hot [B][T][F][2] written as out_dtype (ONSSEN_DT_*): argmax over (mag_s1, mag_s2) (ties -> 0),
 * zeroed where feature < feat_max[b] - db_threshold/20 (strict <). Bit-exact integer result. */
int onssen_one_hot_vad(const float* feature, const float* mag_s1, const float* mag_s2, const float* feat_max,
                       float db_threshold, int B, int T, int F, void* one_hot, int out_dtype, void* stream);

/* Masked iSTFT overlap-add (replaces stft_mix*mask -> librosa.core.istft(..., hop_length, length=nsample),
 * egs/wsj0-2mix/deep_clustering/evaluate.py:42-45, chimera/evaluate.py:40-43).
 * stft_re/stft_im [B][frames][F]; mask [B][S][frames][F] (NULL = all ones); out [B][S][nsample] float32.
 * Periodic Hann synthesis window, window-sum-square normalisation, centre trim n_fft/2. */
size_t onssen_istft_scratch_bytes(int B, int S, int frames, int n_fft);
int onssen_istft_masked(const float* stft_re, const float* stft_im, const float* mask, int B, int S, int frames,
                        int n_fft, int hop, int nsample, float* out, void* scratch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BLSTM stack (replaces torch.nn.LSTM(bidirectional=True, batch_first=True) as called at
 * onssen/nn/deep_clustering.py:34-35, chimera.py:35-36, enhancement.py:43-44, phase_network.py:50,57).
 * Gate order i,f,g,o; h0 = c0 = 0.
 */

/* x [B][T][I] float32 (batch-first, reference layout) -> xh [T*B][Kp] fp16 time-major, zero padded,
 * Kp = 64*ceil(I/64). */
int onssen_pack_input_f16(const float* x, int B, int T, int I, void* xh, int Kp, void* stream);

/* Pack one layer's weights for both directions.
 *   w_ih_{f,r} [4H][I], w_hh_{f,r} [4H][H], b_ih_*, b_hh_* [4H]   (PyTorch parameter layout, fp32)
 *   in_is_blstm: 0 -> input index k maps to column k (layer 0), Kp = 64*ceil(I/64)
 *                1 -> I == 2*Hin; input index j maps to column (j/Hin)*Hinp + j%Hin, Kp = 2*Hinp
 * Outputs:
 *   wih_p  [2*4Hp][Kp] fp16, row n = dir*4Hp + rb*128 + 4*ul + gate  <->  source row gate*H + rb*32 + ul
 *   whh_p  [2][Hp/32][128][Hp] fp16 (per (dir,row-block) a contiguous row-major slab of the 128 permuted gate
 *          rows; each row becomes one TMEM lane of the resident A operand), zero padded
 *   bias_p [2*4Hp] fp32 = b_ih + b_hh, same permutation
 */
int onssen_lstm_pack_layer(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                           const float* w_ih_r, const float* w_hh_r, const float* b_ih_r, const float* b_hh_r,
                           int H, int I, int in_is_blstm, int Hin, void* wih_p, void* whh_p, float* bias_p,
                           void* stream);

/* Linear weight w [N][K] fp32 -> out [N][Kp] fp16 (zero padded). in_is_blstm: K == 2*Hin and input j maps
 * to column (j/Hin)*Hinp + j%Hin (Kp = 2*Hinp) -- the padded channel layout of the BLSTM outputs; else
 * Kp >= K, multiple of 8 (use 64*ceil(K/64)). (fc_dc / fc_mi / fc_pre / fc_post of the reference models) */
int onssen_pack_linear_f16(const float* w, int N, int K, int in_is_blstm, int Hin, void* out, int Kp,
                           void* stream);

/* fp16 tensor-core GEMM with fused epilogue: out = epi(A[M][K] * W[N][K]^T + bias).
 * epi: 0 none, 1 sigmoid, 2 relu, 3 L2-normalise consecutive groups of `group` columns
 * (F.normalize eps 1e-12, deep_clustering.py:41). remap_inner>0: time-major row m=t*B+b (inner=B,outer=T)
 * is written to batch-first output row b*T+t. lda/ldw in elements (multiples of 8). */
int onssen_gemm_l2norm_supported(int group); /* 1 if epi=3 is fused for this group size */
int onssen_gemm_f16(const void* A, const void* W, const float* bias, float* out, int M, int N, int K,
                    long long lda, long long ldw, long long ld_out, int epi, int group, int remap_inner,
                    int remap_outer, void* stream);

/* Workspace (bytes) for the recurrent kernel's h exchange buffers + flags, for batch B, hidden H. */
size_t onssen_blstm_rec_workspace_bytes(int B, int H);

/* Persistent recurrent kernel for one layer, both directions.
 *   gates [T*B][2*4Hp] fp32 pre-activations W_ih x + b (output of onssen_gemm_f16 with wih_p/bias_p)
 *   whh_p packed recurrent weights (onssen_lstm_pack_layer)
 *   y_h   [T*B][2*Hp] fp16 layer output (column dir*Hp + u), NULL to skip
 *   y_f   [T*B][2*Hp] fp32 layer output, NULL to skip
 *   dropout_p > 0: y_h/y_f are scaled by Bernoulli(1-p)/(1-p) (Philox, `seed`,`offset`); the recurrence
 *                  itself uses the undropped h (PyTorch inter-layer dropout semantics)
 *   use_tensor_cores: 1 = tcgen05 path (product), 0 = SIMT fp32-accumulate path (debug/validation)
 * Returns ONSSEN_ERR_UNSUPPORTED if ceil(B / slices) > 32 for the slice count that fits the device;
 * callers split the batch. */
int onssen_blstm_rec_fwd(const float* gates, const void* whh_p, int B, int T, int H, void* y_h, float* y_f,
                         float dropout_p, unsigned long long seed, unsigned long long offset, void* workspace,
                         size_t workspace_bytes, int use_tensor_cores, void* stream);
/* Inference on a zero-padded batch of utterances with different lengths (the evaluation loop of
 * egs/wsj0-2mix/.../evaluate.py runs batch 1; this runs B at once): frames_per_utt [B] int32 on the device. State and output
 * of column b are held at zero for t >= frames_per_utt[b], so every utterance gets exactly its batch-1 result (the
 * reverse direction starts at its own last frame). No dropout, tensor-core path. */
int onssen_blstm_rec_fwd_var(const float* gates, const void* whh_p, int B, int T, int H, void* y_h, float* y_f,
                             const int32_t* frames_per_utt, void* workspace, size_t workspace_bytes, void* stream);

/* Debug/profiling hook: when set to a device buffer of 64 int64, the next recurrent launches record clock64
 * stamps of CTA 0 for steps 100..103 (16 slots per step; see REC_TRACE in csrc/lstm_rec.cu). NULL disables. */
void onssen_blstm_rec_set_trace(void* device_buf_64_int64);
/* Tuning knob: SM cycles every CTA waits between publishing h_t and its first gather round (default tuned on B200). */
void onssen_blstm_rec_set_poll_delay(int cycles);


/* ------------------------------------------------------------------------------------------------
 * BatchNorm1d over (B,T) per channel (replaces permute + nn.BatchNorm1d + permute,
 * deep_clustering.py:36-38, enhancement.py:45-47).
 * y [M][2Hp] fp32 (padded channel layout: channel j of 2H lives at (j/H)*Hp + j%H)
 * training != 0: batch statistics (biased var for normalisation; running stats updated with momentum,
 *                unbiased var), else running statistics.
 * out_h [M][2Hp] fp16 = fp16(bn(y)) (pad columns zero) -- the A operand of the head GEMM.
 * scratch: onssen_bn_scratch_bytes(M, H) bytes (per-chunk fp64 partial sums + folded scale/shift).
 */
int onssen_bn_num_chunks(int M);
size_t onssen_bn_scratch_bytes(int M, int H);
int onssen_bn_forward_f16(const float* y, int M, int H, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float eps, float momentum, int training,
                          void* out_h, float* save_mean, float* save_invstd, void* scratch, void* stream);
/* plain fp32 -> fp16 cast of an [M][2Hp] activation (models without BN, chimera.py) */
int onssen_cast_f16(const float* y, long long n, void* out_h, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Losses
 */
/* Deep-clustering affinity loss (replaces loss_dc, onssen/loss/loss_dc.py:6-44 incl. its un-squared norms
 * and (B,B) result): emb [B][N][D] fp32, label [B][N][S] (label_dtype), mag [B][N] fp32.
 * loss_bb [B][B]: element [i][j] = sum_n mag[i][n] * l_j.  l [B] and mag_sum [B] are also written
 * (l and mag_sum are required as intermediates).
 * scratch: float[B * (onssen_loss_dc_num_chunks(N) + 1) * (D*D + D*S + S*S + 4)] */
int onssen_loss_dc_num_chunks(int N);
int onssen_loss_dc_fwd(const float* emb, const void* label, int label_dtype, const float* mag, int B, int N,
                       int D, int S, float* loss_bb, float* l, float* mag_sum, float* scratch, void* stream);

/* Backward of onssen_loss_dc_fwd w.r.t. the embedding (autograd of loss_dc.py:24-44; labels and mag_mix carry no
 * gradient, mag_mix is detached at :25). summed_record = the per-utterance summed Gram record the forward left
 * at scratch + B*nchunk*(D*D+D*S+S*S+4) floats; g_bb [B][B] = upstream gradient of the (B,B) result;
 * d_emb [B][N][D]. Supported for S == 2, D % 4 == 0 (the forward's fast layout). */
int onssen_loss_dc_bwd(const float* emb, const void* label, int label_dtype, const float* mag,
                       const float* summed_record, const float* g_bb, int B, int N, int D, int S, float* d_emb,
                       void* stream);

/* PIT L1 mask-inference loss for two speakers (replaces the mask part of loss_chimera_msa/psa,
 * onssen/loss/loss_chimera.py:25-29,53-57): per utterance
 *   min( L1(mA*mix - t1) + L1(mB*mix - t2),  L1(mB*mix - t1) + L1(mA*mix - t2) )
 * t_s = mag_s (cos_s == NULL, MSA) or min(mix, relu(mag_s*cos_s)) (PSA). mask strides allow the reference's
 * strided views masks[...,0]/[...,1] (mask_stride = elements between consecutive (t,f) entries).
 * out [B]; perm [B] int32 (0: A->s1, 1: swapped) may be NULL. */
int onssen_loss_pit_l1_fwd(const float* mask_a, const float* mask_b, long long mask_stride, const float* mag_mix,
                           const float* mag_s1, const float* mag_s2, const float* cos_s1, const float* cos_s2,
                           int B, int N, float* out, int32_t* perm, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training path: backward of the deep-clustering stack (autograd of onssen/nn/deep_clustering.py:34-42 as driven
 * by loss_avg.backward() at onssen/utils/train.py:82).  fp32 gradients feed the fp16 tensor-core GEMMs through a
 * power-of-two scale (scale2 = {2^k, 2^-k} on the device) so that nothing underflows.
 */
/* onssen_gemm_f16 with two extras: out_scale (device scalar multiplied into the accumulators, NULL = 1) and, for
 * epi 3, inv_norm [output rows][N/group] = 1/max(||z||, 1e-12) of every normalised group (NULL to skip). */
/* Weight-gradient GEMM (the dW = dZ^T A products autograd forms behind onssen/utils/train.py:82):
 * out[m][n] = out_scale * sum_k X[k][m] * Y[k + y_row_shift][n], X [Kc][ldx], Y [Kc][ldy] fp16 row-major, contracted
 * over their ROWS and read in place (MN-major tcgen05 operands, no transposed copies); rows of Y outside [0, Kc) count
 * as zero (y_row_shift = -B / +B pairs dG_t with h_{t-1} of the forward / reverse direction); out fp32 [M][ld_out]. */
int onssen_gemm_f16_rows(const void* X, const void* Y, float* out, int M, int N, int Kc, long long ldx, long long ldy,
                         long long ld_out, int y_row_shift, const float* out_scale, void* stream);
int onssen_gemm_f16_ex(const void* A, const void* W, const float* bias, float* out, int M, int N, int K,
                       long long lda, long long ldw, long long ld_out, int epi, int group, int remap_inner,
                       int remap_outer, const float* out_scale, float* inv_norm, void* stream);
/* scale2[0] = 2^k with amax(x) * 2^k ~ target, scale2[1] = 2^-k. scratch_u32: 4 bytes. */
int onssen_amax_scale(const float* x, long long n, float target, void* scratch_u32, float* scale2, void* stream);
int onssen_scale_from_amax_bits(const void* amax_bits_u32, float target, float* scale2, void* stream);
/* src fp32 [R][C] (pitch ld) * scale2[0] -> out_n fp16 [R][Cp] and/or out_t fp16 [C][Rp] (zero padded). */
int onssen_cast_transpose_f16(const float* src, int R, int C, long long ld, const float* scale2, void* out_n,
                              int Cp, void* out_t, int Rp, void* stream);
/* fp16 src [R][*] (pitch ld), columns [col0, col0+ncol) -> out_t fp16 [ncol][Rp], out_t[c][r] = src[r-shift][col0+c]
 * (zero outside): transposed and time-shifted copy for the W_hh weight gradient (h_{t-1} against dG_t). */
int onssen_transpose_shift_f16(const void* src, int R, long long ld, int col0, int ncol, int shift, void* out_t,
                               int Rp, void* stream);
size_t onssen_colsum_scratch_bytes(int R, int C);
/* out[c] = mult * sum_r x[r][c] (bias gradients), deterministic two-stage reduction. */
int onssen_colsum(const float* x, int R, int C, long long ld, float mult, float* out, void* scratch, void* stream);
/* F.normalize backward (deep_clustering.py:41): d_emb, emb batch-first (B,T,F,D), inv_norm (B,T,F) ->
 * dz time-major [T*B][F*D] fp32; amax_bits_u32 receives the bit pattern of max|dz|. */
int onssen_normalize_bwd(const float* d_emb, const float* emb, const float* inv_norm, int B, int T, int F, int D,
                         float* dz, void* amax_bits_u32, void* stream);
size_t onssen_bn_backward_scratch_bytes(int M, int H);
/* BatchNorm1d backward (batch statistics) on the padded layout: d_out, y, d_y [M][2Hp] fp32; save_mean /
 * save_invstd from onssen_bn_forward_f16; d_gamma, d_beta [2H]. */
int onssen_bn_backward(const float* d_out, const float* y, int M, int H, const float* gamma, const float* save_mean,
                       const float* save_invstd, float* d_y, float* d_gamma, float* d_beta, void* scratch,
                       void* stream);
/* Inverse of onssen_pack_linear_f16 / onssen_lstm_pack_layer for fp32 gradient buffers:
 * gp [N][Kp] -> g [N][K];  gp [2*4Hp][Kp] rows of `dir` -> g [4H][K] in PyTorch gate order. */
int onssen_unpack_linear_grad(const float* gp, int N, int K, int in_is_blstm, int Hin, int Kp, float* g, void* stream);
int onssen_unpack_lstm_grad(const float* gp, int H, int K, int in_is_blstm, int Hin, int Kp, int dir, float* g,
                            void* stream);
/* W_hh (both directions, fp32 [4H][H]) -> fp16, 2*Hp*4Hp elements: W_hh^T pre-laid-out in mma.m16n8k16 A-fragment
 * order [dir][unit block][k-step][m-tile][lane][4 words] (permuted gate rows, zero padded) so that every operand
 * load of the BPTT step kernel is one coalesced 16-byte access. */
int onssen_lstm_pack_whh_t(const float* w_hh_f, const float* w_hh_r, int H, void* out, void* stream);
/* fp16 elements of `out` above: the fragment layout is followed by the same matrix as the tcgen05 BPTT kernel's
 * tensor-memory slabs [dir][unit block of 128][K quarter][128 units][Hp gate rows]. */
size_t onssen_lstm_pack_whh_t_elems(int H);
/* Forward recurrence that also saves the BPTT state: the activated gates overwrite gates_inout in place, c_out
 * [T*B][2Hp] fp32, h_raw [T*B][2Hp] fp16 = h before dropout (NULL when dropout_p == 0: use y_h). */
int onssen_blstm_rec_fwd_train(float* gates_inout, const void* whh_p, int B, int T, int H, void* y_h, float* y_f,
                               float* c_out, void* h_raw, float dropout_p, unsigned long long seed,
                               unsigned long long offset, void* workspace, size_t workspace_bytes, void* stream);
/* BPTT over one layer, both directions (one launch per step). act_gates: saved activations, overwritten with the
 * fp32 pre-activation gradients dG; dg16 receives dG * scale2[0] in fp16; dy = gradient w.r.t. the layer output
 * (the dropout mask is re-derived from seed/offset); scratch: onssen_blstm_rec_bwd_scratch_bytes(B, H) (cell
 * gradient carry + the per-step dG exchange buffer in B-fragment order), zeroed by the call. */
size_t onssen_blstm_rec_bwd_scratch_bytes(int B, int H);
/* 2 (default): persistent tcgen05 kernel -- W_hh^T resident in tensor memory, K split over 4-CTA clusters, partial dh
 * reduce-scattered through distributed shared memory (falls back to 1 when the shape does not fit one launch);
 * 1: persistent mma.sync kernel (W_hh^T resident in smem). Both exchange dG with a flag bit inside the fp16 data and
 * require the scale passed to onssen_blstm_rec_bwd to keep |dG*scale| < 2 (scale2 from a target <= 2^-4 * headroom);
 * 0: one launch per time step (validation path). */
void onssen_blstm_rec_bwd_set_persistent(int mode);
/* SMs the tcgen05 BPTT kernel leaves free (default 0). Its CTAs must all be co-resident (cooperative launch, one CTA
 * per SM); a data-parallel trainer overlaps the gradient all-reduce of the layer above with this kernel and sets the
 * reserve to what the collective's kernel occupies, so that neither launch has to wait for the other to drain. */
void onssen_blstm_rec_bwd_set_sm_reserve(int sms);
/* debug: clock64 stamps of CTA 0 of the persistent BPTT kernel, steps 100..107, [step][slot 0..7][warp 0..7]
   (512 int64); NULL = off */
void onssen_blstm_rec_bwd_set_trace(void* device_buf_512_int64);
/* sat_count_u32 (optional, device uint32, accumulated): how many published dG values did not fit the flag range of
 * the persistent kernel (|dG*scale| >= 2, or NaN) and were clamped -- an exploding recurrent gradient must not be
 * silent (the trainer reads it at its logging interval). */
int onssen_blstm_rec_bwd(float* act_gates, void* dg16, const float* c, const float* dy, const void* whh_t,
                         void* scratch, const float* scale2, void* sat_count_u32, int B, int T, int H,
                         float dropout_p, unsigned long long seed, unsigned long long offset, void* stream);

/* Backward of onssen_loss_pit_l1_fwd w.r.t. the two masks (g [B] = upstream gradient, perm from the forward);
 * d_mask_a / d_mask_b dense [B][N]. */
int onssen_loss_pit_l1_bwd(const float* mask_a, const float* mask_b, long long mask_stride, const float* mag_mix,
                           const float* mag_s1, const float* mag_s2, const float* cos_s1, const float* cos_s2,
                           const int32_t* perm, const float* g, int B, int N, float* d_mask_a, float* d_mask_b,
                           void* stream);
/* sigmoid backward of the mask head (chimera.py:42): d_out/out batch-first [B][T][C] -> dz time-major [T*B][C]. */
int onssen_sigmoid_bwd(const float* d_out, const float* out, int B, int T, int C, float* dz, void* amax_bits_u32,
                       void* stream);
int onssen_add_inplace(float* a, const float* b, long long n, void* stream);
/* relu backward with the same layout change as onssen_sigmoid_bwd (enhancement.py:49,51). */
int onssen_relu_bwd(const float* d_out, const float* out, int B, int T, int C, float* dz, void* amax_bits_u32,
                    void* stream);
/* backward of est = relu-output `pre` * sigmoid-output `mask` (enhancement.py:49-50), all time-major [M][F] (d_est with
 * row pitch ld): dz_pre = d_est*mask*(pre>0), dz_mi = d_est*pre*mask*(1-mask); amax bit patterns of both in
 * amax_bits_2xu32[0..1]. */
int onssen_enhance_mid_bwd(const float* d_est, long long ld, const float* pre, const float* mask, long long M, int F,
                           float* dz_pre, float* dz_mi, void* amax_bits_2xu32, void* stream);
/* nn.MSELoss backward: d_a = 2 (a - b) / n * g[0]  (g = device scalar upstream gradient). */
int onssen_loss_mse_bwd(const float* a, const float* b, long long n, const float* g, float* d_a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Enhancement / phase-network variants of the path
 */
/* out[m][k] = fp16(a[m][k] * b[m][k]) for k < F, zero padded to Kp: mask * relu(fc_pre(mag_noisy)), the fp16
 * operand of fc_post (onssen/nn/enhancement.py:49-50). a, b [M][F] fp32. */
int onssen_mul_pack_f16(const float* a, const float* b, long long M, int F, void* out, int Kp, void* stream);

/* Second-BLSTM input of phase_net (onssen/nn/phase_network.py:46-49,56): time-major fp16 rows
 * [ x_mag*mask (F) | x_phase viewed (2F) | 0 pad ]; x_mag [B][T][F], mask element (b,t,f) at
 * mask[((b*T+t)*F+f)*mask_stride], x_phase [B][T][F][2]. Kp = 64*ceil(3F/64). */
int onssen_pack_phase_input_f16(const float* x_mag, const float* mask, long long mask_stride,
                                const float* x_phase, int B, int T, int F, void* out, int Kp, void* stream);

/* out = F.normalize(x + residual, dim=-1) over (re,im) pairs (phase_network.py:63-66). */
int onssen_add_l2norm_pairs(const float* x, const float* residual, long long npairs, float* out, void* stream);

/* loss_mask_psa (onssen/loss/loss_mask.py:25-40): per utterance sum |mask*noisy - min(noisy, relu(clean*cos))| */
int onssen_loss_l1_psa_fwd(const float* mask, const float* mag_noisy, const float* mag_clean,
                           const float* cos_diff, int B, int N, float* out, void* stream);

/* loss_mask_msa (onssen/loss/loss_mask.py:6-22): nn.MSELoss()(a, b) -> out[0]. scratch: 256 doubles. */
int onssen_loss_mse_fwd(const float* a, const float* b, long long n, float* out, void* scratch, void* stream);

/* phase term of loss_phase (onssen/loss/loss_phase.py:26-35): per utterance
 * -sum mag_mix * (cos_sim(pX, s1) + cos_sim(pY, s2)), (X,Y) = (A,B) when perm[b]==0 else (B,A);
 * phases [B][N][2], cosine_similarity eps 1e-8. */
int onssen_loss_phase_cos_fwd(const float* phase_a, const float* phase_b, const float* phase_s1,
                              const float* phase_s2, const float* mag_mix, const int32_t* perm, int B, int N,
                              float* out, void* stream);

/* ---- phase network training (REPAIRED restatement; no reference oracle) ----------------
 * loss_phase.py:26-35 backward: gradient of -sum mag*cos_sim w.r.t. the two phase estimates under `perm`;
 * g = upstream gradient per utterance [B]. */
int onssen_loss_phase_cos_bwd(const float* phase_a, const float* phase_b, const float* phase_s1, const float* phase_s2,
                              const float* mag_mix, const int32_t* perm, const float* g, int B, int N,
                              float* d_phase_a, float* d_phase_b, void* stream);
/* phase_network.py:55-56,63-64 backward: y = normalize(x + residual) over (re,im); d_y/x/residual [B][T][F][2];
 * dz = d loss / d x, time-major [T*B][2F]; amax_bits_u32 receives max|dz| as float bits. */
int onssen_l2norm_pairs_bwd(const float* d_y, const float* x, const float* residual, int B, int T, int F, float* dz,
                            void* amax_bits_u32, void* stream);
/* phase_network.py:47-49 backward: d_masks[b][t][f][s_idx] += d_xin[t*B+b][f] * x_mag[b][t][f]
 * (d_xin = gradient of the second BLSTM's packed input, row stride ld). */
int onssen_phase_input_bwd(const float* d_xin, long long ld, const float* x_mag, int B, int T, int F, int S, int s_idx,
                           float* d_masks, void* stream);

/* loss_mask.py:25-40 backward: d_mask = g[b] * sign(mask*noisy - min(noisy, relu(clean*cos))) * noisy. */
int onssen_loss_l1_psa_bwd(const float* mask, const float* mag_noisy, const float* mag_clean, const float* cos_diff,
                           const float* g, int B, int N, float* d_mask, void* stream);

/* ---- optimiser step (onssen/utils/train.py:83-84, onssen/utils/basic.py:6-7) ------------------------------------
 * `tensors`: device array of records {float* param; float* grad; float* exp_avg; float* exp_avg_sq; int64 numel}
 * (40 bytes each); `chunks`: device array of {int64 tensor_index; int64 start} (16 bytes), one per CUDA block,
 * each covering chunk_elems consecutive elements of one tensor.
 * onssen_clip_grad_norm = torch.nn.utils.clip_grad_norm_(params, max_norm): out2 = {total L2 norm, clip coefficient
 * min(1, max_norm/(norm+1e-6))}; gradients are scaled in place only when the coefficient is < 1.  partials_f64:
 * nchunks doubles of scratch.  No host synchronisation. */
int onssen_clip_grad_norm(const void* tensors, const void* chunks, int nchunks, int chunk_elems, float max_norm,
                          void* partials_f64, float* out2, void* stream);
/* torch.optim.Adam.step (amsgrad off); `step` is the 1-based step count used for the bias corrections. */
int onssen_adam_step(const void* tensors, const void* chunks, int nchunks, int chunk_elems, float lr, float beta1,
                     float beta2, float eps, float weight_decay, long long step, void* stream);

/* ---- BatchNorm with cross-rank batch statistics (per-replica statistics differ from the
 * single-device maths of deep_clustering.py:36-38; this is the parity mode).  Forward: onssen_bn_stats writes this
 * rank's column sums and sums of squares as 2*2Hp doubles; the caller all-reduces them (SUM) and calls
 * onssen_bn_forward_f16_stats with M_total = rows over all ranks.  Backward likewise: onssen_bn_backward_stats ->
 * all-reduce -> onssen_bn_backward_apply (d_gamma / d_beta stay this rank's local sums, as data-parallel gradient
 * averaging expects; the SAME scratch buffer must be passed to both backward calls). */
int onssen_bn_stats(const float* y, int M, int H, void* sums_f64, void* scratch, void* stream);
int onssen_bn_forward_f16_stats(const float* y, int M, int M_total, int H, const void* sums_f64, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float eps, float momentum,
                                void* out_h, float* save_mean, float* save_invstd, void* scratch, void* stream);
int onssen_bn_backward_stats(const float* d_out, const float* y, int M, int H, const float* save_mean,
                             const float* save_invstd, void* sums_f64, void* scratch, void* stream);
int onssen_bn_backward_apply(const float* d_out, const float* y, int M, int M_total, int H, const float* gamma,
                             const float* save_mean, const float* save_invstd, const void* sums_total_f64, float* d_y,
                             float* d_gamma, float* d_beta, void* scratch, void* stream);

/* ---- deep-clustering inference: VAD + K-means on the embeddings -> masks
 * (egs/wsj0-2mix/deep_clustering/evaluate.py:36-41).  emb [N][D] (N = frames*F of ONE utterance, D <= 64),
 * feature [N] log-magnitudes (NULL = every bin active); active = feature >= max(feature) - db_threshold/20;
 * K <= 3 clusters, `iters` Lloyd iterations after a deterministic farthest-point seeding (sklearn's k-means++ RNG is
 * not reproduced; partitions are compared up to label permutation).  masks [K][N] (K = 2: mask[0] = label,
 * mask[1] = 1 - label on active bins, 0 elsewhere) and/or labels [N] int32 (-1 inactive).  scratch:
 * onssen_kmeans_scratch_bytes() bytes. */
size_t onssen_kmeans_scratch_bytes(void);
int onssen_kmeans_masks(const float* emb, const float* feature, long long N, int D, int K, float db_threshold,
                        int iters, float* masks, int32_t* labels, void* scratch, void* stream);

/* ---- wav decode on the device (onssen/data/feature_utils.py:15-19) -----------------------------
 * pcm_i16 [R][pitch*channels] interleaved int16, frames[r] valid frames of row r -> out [R][pitch] float32 mono
 * (sample / 32768, mean over channels: what librosa.load(sr=None) returns), zero beyond frames[r]. */
int onssen_pcm16_to_f32(const void* pcm_i16, int R, int pitch, int channels, const int32_t* frames, float* out,
                        void* stream);
/* Polyphase resampling y = upfirdn(h, x, up, down)[n_pre_remove : n_pre_remove + ceil(n_in*up/down)] per row, i.e.
 * scipy.signal.resample_poly with the filter h (designed on the host, data/wavio.py:resample_filter); stands in for
 * librosa.core.resample (feature_utils.py:19).  x [R][pitch_in], y [R][pitch_out] (zero beyond the row's length). */
int onssen_resample_poly(const float* x, int R, int pitch_in, const int32_t* n_in, int up, int down, const float* h,
                         int hlen, int n_pre_remove, float* y, int pitch_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ONSSEN_B200_H_ */
