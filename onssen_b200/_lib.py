"""ctypes binding of libonssen_b200.so (the C ABI declared in include/onssen_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used only as the owner of device memory / streams; every function here takes torch CUDA
tensors, passes raw device pointers + the current CUDA stream, and checks the integer status.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libonssen_b200.so")

DT_F32, DT_F64, DT_U8 = 0, 1, 2
_DT = {torch.float32: DT_F32, torch.float64: DT_F64, torch.uint8: DT_U8}

c_int, c_ll, c_vp, c_sz, c_f, c_ull = (ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_size_t,
                                       ctypes.c_float, ctypes.c_ulonglong)

# name -> (restype, argtypes); mirrors include/onssen_b200.h one to one
_SIGNATURES = {
    "onssen_version": (ctypes.c_char_p, []),
    "onssen_error_string": (ctypes.c_char_p, [c_int]),
    "onssen_num_sms": (c_int, []),
    "onssen_stft_features": (c_int, [c_vp] * 3 + [c_int] * 4 + [c_vp, c_int] + [c_vp] * 10 + [c_vp, c_vp]),
    "onssen_one_hot_vad": (c_int, [c_vp] * 4 + [c_f, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "onssen_istft_scratch_bytes": (c_sz, [c_int] * 4),
    "onssen_istft_masked": (c_int, [c_vp] * 3 + [c_int] * 6 + [c_vp, c_vp, c_vp]),
    "onssen_pack_input_f16": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "onssen_lstm_pack_layer": (c_int, [c_vp] * 8 + [c_int] * 4 + [c_vp] * 4),
    "onssen_pack_linear_f16": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "onssen_gemm_l2norm_supported": (c_int, [c_int]),
    "onssen_gemm_f16": (c_int, [c_vp] * 4 + [c_int] * 3 + [c_ll] * 3 + [c_int] * 4 + [c_vp]),
    "onssen_gemm_f16_rows": (c_int, [c_vp] * 3 + [c_int] * 3 + [c_ll] * 3 + [c_int] + [c_vp] * 2),
    "onssen_blstm_rec_workspace_bytes": (c_sz, [c_int, c_int]),
    "onssen_blstm_rec_set_trace": (None, [c_vp]),
    "onssen_blstm_rec_set_poll_delay": (None, [c_int]),
    "onssen_pcm16_to_f32": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "onssen_resample_poly": (c_int, [c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_int, c_int, c_vp, c_int, c_vp]),
    "onssen_blstm_rec_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_f, c_ull, c_ull, c_vp, c_sz,
                                     c_int, c_vp]),
    "onssen_blstm_rec_fwd_var": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "onssen_bn_num_chunks": (c_int, [c_int]),
    "onssen_bn_scratch_bytes": (c_sz, [c_int, c_int]),
    "onssen_bn_forward_f16": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_int, c_vp, c_vp, c_vp,
                                      c_vp, c_vp]),
    "onssen_cast_f16": (c_int, [c_vp, c_ll, c_vp, c_vp]),
    "onssen_loss_dc_num_chunks": (c_int, [c_int]),
    "onssen_loss_dc_fwd": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                   c_vp]),
    "onssen_loss_dc_bwd": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "onssen_loss_pit_l1_fwd": (c_int, [c_vp, c_vp, c_ll] + [c_vp] * 5 + [c_int, c_int, c_vp, c_vp, c_vp]),
    "onssen_gemm_f16_ex": (c_int, [c_vp] * 4 + [c_int] * 3 + [c_ll] * 3 + [c_int] * 4 + [c_vp, c_vp, c_vp]),
    "onssen_amax_scale": (c_int, [c_vp, c_ll, c_f, c_vp, c_vp, c_vp]),
    "onssen_scale_from_amax_bits": (c_int, [c_vp, c_f, c_vp, c_vp]),
    "onssen_cast_transpose_f16": (c_int, [c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_int, c_vp, c_int, c_vp]),
    "onssen_transpose_shift_f16": (c_int, [c_vp, c_int, c_ll, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "onssen_colsum_scratch_bytes": (c_sz, [c_int, c_int]),
    "onssen_colsum": (c_int, [c_vp, c_int, c_int, c_ll, c_f, c_vp, c_vp, c_vp]),
    "onssen_normalize_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "onssen_bn_backward_scratch_bytes": (c_sz, [c_int, c_int]),
    "onssen_bn_backward": (c_int, [c_vp, c_vp, c_int, c_int] + [c_vp] * 8),
    "onssen_unpack_linear_grad": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "onssen_unpack_lstm_grad": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "onssen_lstm_pack_whh_t": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp]),
    "onssen_lstm_pack_whh_t_elems": (c_sz, [c_int]),
    "onssen_blstm_rec_fwd_train": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_f, c_ull, c_ull,
                                           c_vp, c_sz, c_vp]),
    "onssen_blstm_rec_bwd_scratch_bytes": (c_sz, [c_int, c_int]),
    "onssen_blstm_rec_bwd_set_persistent": (None, [c_int]),
    "onssen_blstm_rec_bwd_set_trace": (None, [c_vp]),
    "onssen_blstm_rec_bwd_set_sm_reserve": (None, [c_int]),
    "onssen_clip_grad_norm": (c_int, [c_vp, c_vp, c_int, c_int, c_f, c_vp, c_vp, c_vp]),
    "onssen_adam_step": (c_int, [c_vp, c_vp, c_int, c_int, c_f, c_f, c_f, c_f, c_f, c_ll, c_vp]),
    "onssen_kmeans_scratch_bytes": (c_sz, []),
    "onssen_kmeans_masks": (c_int, [c_vp, c_vp, c_ll, c_int, c_int, c_f, c_int, c_vp, c_vp, c_vp, c_vp]),
    "onssen_bn_stats": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "onssen_bn_forward_f16_stats": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_vp, c_vp,
                                            c_vp, c_vp, c_vp]),
    "onssen_bn_backward_stats": (c_int, [c_vp, c_vp, c_int, c_int] + [c_vp] * 5),
    "onssen_bn_backward_apply": (c_int, [c_vp, c_vp, c_int, c_int, c_int] + [c_vp] * 9),
    "onssen_loss_l1_psa_bwd": (c_int, [c_vp] * 5 + [c_int] * 2 + [c_vp] * 2),
    "onssen_loss_phase_cos_bwd": (c_int, [c_vp] * 7 + [c_int] * 2 + [c_vp] * 3),
    "onssen_l2norm_pairs_bwd": (c_int, [c_vp] * 3 + [c_int] * 3 + [c_vp] * 3),
    "onssen_phase_input_bwd": (c_int, [c_vp, c_ll, c_vp] + [c_int] * 5 + [c_vp] * 2),
    "onssen_blstm_rec_bwd": (c_int, [c_vp] * 8 + [c_int, c_int, c_int, c_f, c_ull, c_ull, c_vp]),
    "onssen_loss_pit_l1_bwd": (c_int, [c_vp, c_vp, c_ll] + [c_vp] * 7 + [c_int, c_int, c_vp, c_vp, c_vp]),
    "onssen_sigmoid_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "onssen_add_inplace": (c_int, [c_vp, c_vp, c_ll, c_vp]),
    "onssen_relu_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "onssen_enhance_mid_bwd": (c_int, [c_vp, c_ll, c_vp, c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_vp]),
    "onssen_loss_mse_bwd": (c_int, [c_vp, c_vp, c_ll, c_vp, c_vp, c_vp]),
    "onssen_mul_pack_f16": (c_int, [c_vp, c_vp, c_ll, c_int, c_vp, c_int, c_vp]),
    "onssen_pack_phase_input_f16": (c_int, [c_vp, c_vp, c_ll, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp]),
    "onssen_add_l2norm_pairs": (c_int, [c_vp, c_vp, c_ll, c_vp, c_vp]),
    "onssen_loss_l1_psa_fwd": (c_int, [c_vp] * 4 + [c_int, c_int, c_vp, c_vp]),
    "onssen_loss_mse_fwd": (c_int, [c_vp, c_vp, c_ll, c_vp, c_vp, c_vp]),
    "onssen_loss_phase_cos_fwd": (c_int, [c_vp] * 6 + [c_int, c_int, c_vp, c_vp]),
}

_lib = None

# bookkeeping for bench.py: number of kernels launched through this binding, and optional CUDA-event pairs
# around the recurrent kernel launches (set REC_EVENTS to a list to collect them)
LAUNCHES = [0]
REC_EVENTS = None
_KERNELS_PER_CALL = {"onssen_stft_features": 2, "onssen_one_hot_vad": 1, "onssen_istft_masked": 2,
                     "onssen_pack_input_f16": 1, "onssen_lstm_pack_layer": 3, "onssen_pack_linear_f16": 1,
                     "onssen_gemm_f16": 1, "onssen_blstm_rec_fwd": 1, "onssen_bn_forward_f16": 3,
                     "onssen_cast_f16": 1, "onssen_loss_dc_fwd": 4, "onssen_loss_dc_bwd": 1, "onssen_loss_pit_l1_fwd": 1,
                     "onssen_mul_pack_f16": 1, "onssen_pack_phase_input_f16": 1, "onssen_add_l2norm_pairs": 1,
                     "onssen_loss_l1_psa_fwd": 1, "onssen_loss_mse_fwd": 2, "onssen_loss_phase_cos_fwd": 1}


class OnssenB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OnssenB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(onssen_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    if os.environ.get("ONSSEN_BPTT_MODE"):   # 2 (default) tcgen05 cluster kernel, 1 mma.sync persistent, 0 per step
        lib.onssen_blstm_rec_bwd_set_persistent(int(os.environ["ONSSEN_BPTT_MODE"]))
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def _check(rc, what):
    LAUNCHES[0] += _KERNELS_PER_CALL.get(what, 0)
    if rc != 0:
        msg = load().onssen_error_string(rc).decode()
        detail = ""
        if rc == -3 and torch.cuda.is_available():
            try:
                torch.cuda.synchronize()
            except Exception as e:  # surface the sticky CUDA error text
                detail = f" ({e})"
        raise OnssenB200Error(f"{what} failed: {msg} [{rc}]{detail}")


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise OnssenB200Error(f"{name} must be a CUDA tensor (onssen_b200 has no CPU path)")
    if not t.is_contiguous():
        raise OnssenB200Error(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise OnssenB200Error(f"{name} must be {dtype}, got {t.dtype}")
    return t


def num_sms():
    return load().onssen_num_sms()


def hp_of(H):
    return (H + 31) // 32 * 32


# ------------------------------------------------------------------------------------------------ featurizer
def stft_features(wav_mix, wav_s1, wav_s2, n_fft, hop, crop_start, T, want, lengths=None):
    """want: iterable of output names among feature, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2, ph_mix, ph_s1,
    ph_s2, feat_max. Returns dict name -> tensor."""
    lib = load()
    _req(wav_mix, torch.float32, "wav_mix")
    B, ns = wav_mix.shape
    F = n_fft // 2 + 1
    dev = wav_mix.device
    out = {}
    for k in want:
        if k == "feat_max":
            out[k] = torch.empty(B, device=dev, dtype=torch.float32)
        elif k.startswith("ph_"):
            out[k] = torch.empty(B, T, F, 2, device=dev, dtype=torch.float32)
        else:
            out[k] = torch.empty(B, T, F, device=dev, dtype=torch.float32)
    g = out.get
    crop_start = _req(crop_start.to(device=dev, dtype=torch.int32), torch.int32, "crop_start")
    if wav_s1 is not None:
        _req(wav_s1, torch.float32, "wav_s1"); _req(wav_s2, torch.float32, "wav_s2")
    rc = lib.onssen_stft_features(_p(wav_mix), _p(wav_s1), _p(wav_s2), B, ns, n_fft, hop, _p(crop_start), T,
                                  _p(g("feature")), _p(g("mag_mix")), _p(g("mag_s1")), _p(g("mag_s2")),
                                  _p(g("cos_s1")), _p(g("cos_s2")), _p(g("ph_mix")), _p(g("ph_s1")), _p(g("ph_s2")),
                                  _p(g("feat_max")),
                                  _p(None if lengths is None else _req(lengths.to(device=dev, dtype=torch.int32),
                                                                       torch.int32, "lengths")), _stream())
    _check(rc, "onssen_stft_features")
    return out


def one_hot_vad(feature, mag_s1, mag_s2, feat_max, db_threshold, dtype=torch.float32):
    lib = load()
    B, T, F = feature.shape
    out = torch.empty(B, T, F, 2, device=feature.device, dtype=dtype)
    rc = lib.onssen_one_hot_vad(_p(_req(feature, torch.float32)), _p(_req(mag_s1, torch.float32)),
                                _p(_req(mag_s2, torch.float32)), _p(_req(feat_max, torch.float32)),
                                float(db_threshold), B, T, F, _p(out), _DT[dtype], _stream())
    _check(rc, "onssen_one_hot_vad")
    return out


def istft_masked(stft_re, stft_im, mask, n_fft, hop, nsample):
    lib = load()
    B, frames, F = stft_re.shape
    assert F == n_fft // 2 + 1
    S = 1 if mask is None else mask.shape[1]
    out = torch.empty(B, S, nsample, device=stft_re.device, dtype=torch.float32)
    scratch = torch.empty(lib.onssen_istft_scratch_bytes(B, S, frames, n_fft) // 4, device=stft_re.device,
                          dtype=torch.float32)
    rc = lib.onssen_istft_masked(_p(_req(stft_re, torch.float32)), _p(_req(stft_im, torch.float32)),
                                 _p(None if mask is None else _req(mask, torch.float32)), B, S, frames, n_fft, hop,
                                 nsample, _p(out), _p(scratch), _stream())
    _check(rc, "onssen_istft_masked")
    return out


# ------------------------------------------------------------------------------------------------ BLSTM pieces
def pack_input_f16(x):
    lib = load()
    B, T, I = x.shape
    Kp = (I + 63) // 64 * 64
    xh = torch.empty(T * B, Kp, device=x.device, dtype=torch.float16)
    _check(lib.onssen_pack_input_f16(_p(_req(x, torch.float32, "x")), B, T, I, _p(xh), Kp, _stream()),
           "onssen_pack_input_f16")
    return xh


def lstm_pack_layer(wf, wr, H, I, in_is_blstm, Hin):
    """wf / wr: (w_ih, w_hh, b_ih, b_hh) fp32 CUDA tensors for the forward / reverse direction."""
    lib = load()
    Hp = hp_of(H)
    Kp = 2 * hp_of(Hin) if in_is_blstm else (I + 63) // 64 * 64
    dev = wf[0].device
    wih_p = torch.empty(2 * 4 * Hp, Kp, device=dev, dtype=torch.float16)
    whh_p = torch.empty(2 * 4 * Hp * Hp, device=dev, dtype=torch.float16)
    bias_p = torch.empty(2 * 4 * Hp, device=dev, dtype=torch.float32)
    args = [_p(_req(t.detach(), torch.float32, "lstm weight")) for t in list(wf) + list(wr)]
    rc = lib.onssen_lstm_pack_layer(*args, H, I, int(in_is_blstm), Hin, _p(wih_p), _p(whh_p), _p(bias_p), _stream())
    _check(rc, "onssen_lstm_pack_layer")
    return wih_p, whh_p, bias_p


def pack_linear_f16(w, in_is_blstm, Hin=0):
    lib = load()
    N, K = w.shape
    Kp = 2 * hp_of(Hin) if in_is_blstm else (K + 63) // 64 * 64
    out = torch.empty(N, Kp, device=w.device, dtype=torch.float16)
    rc = lib.onssen_pack_linear_f16(_p(_req(w.detach(), torch.float32, "weight")), N, K, int(in_is_blstm), Hin,
                                    _p(out), Kp, _stream())
    _check(rc, "onssen_pack_linear_f16")
    return out


def gemm_f16(A, W, bias, out, M, N, K, ld_out, epi=0, group=0, remap_inner=0, remap_outer=0):
    lib = load()
    _req(A, torch.float16, "A"); _req(W, torch.float16, "W"); _req(out, torch.float32, "out")
    rc = lib.onssen_gemm_f16(_p(A), _p(W), _p(bias), _p(out), M, N, K, A.stride(0), W.stride(0), ld_out, epi, group,
                             remap_inner, remap_outer, _stream())
    _check(rc, "onssen_gemm_f16")
    return out


def gemm_l2norm_supported(group):
    return bool(load().onssen_gemm_l2norm_supported(group))


def blstm_rec_workspace(B, H, device):
    n = load().onssen_blstm_rec_workspace_bytes(B, H)
    return torch.empty(n, device=device, dtype=torch.uint8)


def blstm_rec_fwd(gates, whh_p, B, T, H, y_h=None, y_f=None, dropout_p=0.0, seed=0, offset=0, workspace=None,
                  use_tensor_cores=True, col_len=None):
    """col_len: optional int32 CUDA tensor [B], frames per utterance of a zero-padded batch (inference only)."""
    lib = load()
    if workspace is None:
        workspace = blstm_rec_workspace(B, H, gates.device)
    if col_len is not None:
        if dropout_p != 0.0 or not use_tensor_cores:
            raise OnssenB200Error("per-utterance lengths are an inference feature of the tensor-core path (no dropout)")
        rc = lib.onssen_blstm_rec_fwd_var(_p(_req(gates, torch.float32, "gates")), _p(whh_p), B, T, H, _p(y_h), _p(y_f),
                                          _p(_req(col_len, torch.int32, "col_len")), _p(workspace), workspace.numel(),
                                          _stream())
        _check(rc, "onssen_blstm_rec_fwd")
        return
    ev = None
    if REC_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    rc = lib.onssen_blstm_rec_fwd(_p(_req(gates, torch.float32, "gates")), _p(whh_p), B, T, H, _p(y_h), _p(y_f),
                                  float(dropout_p), int(seed), int(offset), _p(workspace), workspace.numel(),
                                  int(use_tensor_cores), _stream())
    if ev is not None:
        ev[1].record()
        REC_EVENTS.append(ev)
    _check(rc, "onssen_blstm_rec_fwd")


def bn_forward_f16(y, M, H, gamma, beta, running_mean, running_var, eps, momentum, training, save_stats=False,
                   sync=None):
    """sync = (all_reduce_sum(tensor) -> None, world_size) switches train-mode statistics to the whole data-parallel
    batch (equal shards assumed: M_total = M * world_size)."""
    lib = load()
    Hp = hp_of(H)
    out_h = torch.empty(M, 2 * Hp, device=y.device, dtype=torch.float16)
    scratch = torch.empty(lib.onssen_bn_scratch_bytes(M, H), device=y.device, dtype=torch.uint8)
    sm = si = None
    if save_stats:
        sm = torch.empty(2 * H, device=y.device, dtype=torch.float32)
        si = torch.empty(2 * H, device=y.device, dtype=torch.float32)
    if sync is not None and training:
        reduce_sum, world = sync
        sums = torch.empty(2 * 2 * Hp, device=y.device, dtype=torch.float64)
        _check(lib.onssen_bn_stats(_p(_req(y, torch.float32, "y")), M, H, _p(sums), _p(scratch), _stream()),
               "onssen_bn_stats")
        reduce_sum(sums)
        rc = lib.onssen_bn_forward_f16_stats(_p(y), M, M * world, H, _p(sums), _p(gamma), _p(beta), _p(running_mean),
                                             _p(running_var), float(eps), float(momentum), _p(out_h), _p(sm), _p(si),
                                             _p(scratch), _stream())
        _check(rc, "onssen_bn_forward_f16_stats")
        return out_h, sm, si
    rc = lib.onssen_bn_forward_f16(_p(_req(y, torch.float32, "y")), M, H, _p(gamma), _p(beta), _p(running_mean),
                                   _p(running_var), float(eps), float(momentum), int(training), _p(out_h), _p(sm),
                                   _p(si), _p(scratch), _stream())
    _check(rc, "onssen_bn_forward_f16")
    return out_h, sm, si


def cast_f16(y):
    lib = load()
    out = torch.empty(y.shape, device=y.device, dtype=torch.float16)
    _check(lib.onssen_cast_f16(_p(_req(y, torch.float32, "y")), y.numel(), _p(out), _stream()), "onssen_cast_f16")
    return out


# ------------------------------------------------------------------------------------------------ losses
def loss_dc_fwd(emb, label, mag, return_record=False):
    """emb (B,N,D) fp32, label (B,N,S) f32/f64/u8, mag (B,N) fp32 -> (loss_bb (B,B), l (B,), mag_sum (B,))"""
    lib = load()
    B, N, D = emb.shape
    S = label.shape[-1]
    if label.dtype not in _DT:
        label = label.float()
    dev = emb.device
    loss_bb = torch.empty(B, B, device=dev, dtype=torch.float32)
    l = torch.empty(B, device=dev, dtype=torch.float32)
    msum = torch.empty(B, device=dev, dtype=torch.float32)
    nchunk = lib.onssen_loss_dc_num_chunks(N)
    scratch = torch.empty(B * (nchunk + 1) * (D * D + D * S + S * S + 4), device=dev, dtype=torch.float32)
    rc = lib.onssen_loss_dc_fwd(_p(_req(emb, torch.float32, "embedding")), _p(_req(label, None, "label")),
                                _DT[label.dtype], _p(_req(mag, torch.float32, "mag_mix")), B, N, D, S, _p(loss_bb),
                                _p(l), _p(msum), _p(scratch), _stream())
    _check(rc, "onssen_loss_dc_fwd")
    if return_record:
        rmax = D * D + D * S + S * S + 4
        return loss_bb, l, msum, scratch[B * nchunk * rmax:]
    return loss_bb, l, msum


def loss_dc_bwd(emb, label, mag, summed_record, g_bb):
    lib = load()
    B, N, D = emb.shape
    S = label.shape[-1]
    d_emb = torch.empty_like(emb)
    rc = lib.onssen_loss_dc_bwd(_p(_req(emb, torch.float32)), _p(_req(label)), _DT[label.dtype],
                                _p(_req(mag, torch.float32)), _p(summed_record),
                                _p(_req(g_bb.float().contiguous(), torch.float32)), B, N, D, S, _p(d_emb), _stream())
    _check(rc, "onssen_loss_dc_bwd")
    return d_emb


def loss_pit_l1_fwd(mask_a, mask_b, mask_stride, mag_mix, mag_s1, mag_s2, cos_s1=None, cos_s2=None):
    lib = load()
    B = mag_mix.shape[0]
    N = mag_mix.numel() // B
    out = torch.empty(B, device=mag_mix.device, dtype=torch.float32)
    perm = torch.empty(B, device=mag_mix.device, dtype=torch.int32)
    rc = lib.onssen_loss_pit_l1_fwd(_p(mask_a), _p(mask_b), mask_stride, _p(_req(mag_mix, torch.float32)),
                                    _p(_req(mag_s1, torch.float32)), _p(_req(mag_s2, torch.float32)),
                                    _p(None if cos_s1 is None else _req(cos_s1, torch.float32)),
                                    _p(None if cos_s2 is None else _req(cos_s2, torch.float32)), B, N, _p(out),
                                    _p(perm), _stream())
    _check(rc, "onssen_loss_pit_l1_fwd")
    return out, perm


# ------------------------------------------------------------------------------------------------ enhance / phase-net
def mul_pack_f16(a, b):
    lib = load()
    M, F = a.shape
    Kp = (F + 63) // 64 * 64
    out = torch.empty(M, Kp, device=a.device, dtype=torch.float16)
    _check(lib.onssen_mul_pack_f16(_p(_req(a, torch.float32)), _p(_req(b, torch.float32)), M, F, _p(out), Kp,
                                   _stream()), "onssen_mul_pack_f16")
    return out


def pack_phase_input_f16(x_mag, mask, mask_stride, x_phase):
    lib = load()
    B, T, F = x_mag.shape
    Kp = (3 * F + 63) // 64 * 64
    out = torch.empty(T * B, Kp, device=x_mag.device, dtype=torch.float16)
    rc = lib.onssen_pack_phase_input_f16(_p(_req(x_mag, torch.float32)), _p(mask), mask_stride,
                                         _p(_req(x_phase, torch.float32)), B, T, F, _p(out), Kp, _stream())
    _check(rc, "onssen_pack_phase_input_f16")
    return out


def add_l2norm_pairs(x, residual):
    lib = load()
    out = torch.empty_like(x)
    _check(lib.onssen_add_l2norm_pairs(_p(_req(x, torch.float32)), _p(_req(residual, torch.float32)), x.numel() // 2,
                                       _p(out), _stream()), "onssen_add_l2norm_pairs")
    return out


def loss_l1_psa_fwd(mask, noisy, clean, cosd):
    lib = load()
    B = noisy.shape[0]
    N = noisy.numel() // B
    out = torch.empty(B, device=noisy.device, dtype=torch.float32)
    rc = lib.onssen_loss_l1_psa_fwd(_p(_req(mask, torch.float32)), _p(_req(noisy, torch.float32)),
                                    _p(_req(clean, torch.float32)), _p(_req(cosd, torch.float32)), B, N, _p(out),
                                    _stream())
    _check(rc, "onssen_loss_l1_psa_fwd")
    return out


def loss_mse_fwd(a, b):
    lib = load()
    out = torch.empty(1, device=a.device, dtype=torch.float32)
    scratch = torch.empty(256, device=a.device, dtype=torch.float64)
    _check(lib.onssen_loss_mse_fwd(_p(_req(a, torch.float32)), _p(_req(b, torch.float32)), a.numel(), _p(out),
                                   _p(scratch), _stream()), "onssen_loss_mse_fwd")
    return out[0]


def loss_phase_cos_fwd(pa, pb, s1, s2, mag, perm):
    lib = load()
    B = mag.shape[0]
    N = mag.numel() // B
    out = torch.empty(B, device=mag.device, dtype=torch.float32)
    rc = lib.onssen_loss_phase_cos_fwd(*[_p(_req(t, torch.float32)) for t in (pa, pb, s1, s2, mag)],
                                       _p(_req(perm, torch.int32)), B, N, _p(out), _stream())
    _check(rc, "onssen_loss_phase_cos_fwd")
    return out


# ------------------------------------------------------------------------------------------------ training path
def pad64(n):
    return (n + 63) // 64 * 64


def gemm_f16_ex(A, W, bias, out, M, N, K, ld_out, epi=0, group=0, remap_inner=0, remap_outer=0, out_scale=None,
                inv_norm=None):
    lib = load()
    rc = lib.onssen_gemm_f16_ex(_p(_req(A, torch.float16)), _p(_req(W, torch.float16)), _p(bias), _p(out), M, N, K,
                                A.stride(0), W.stride(0), ld_out, epi, group, remap_inner, remap_outer,
                                _p(out_scale), _p(inv_norm), _stream())
    _check(rc, "onssen_gemm_f16")
    return out


def gemm_f16_rows(X, Y, out, M, N, Kc, y_row_shift=0, out_scale=None):
    """out[m][n] = out_scale * sum_k X[k][m] * Y[k + y_row_shift][n] (X, Y fp16 row-major views, out fp32 [M][N])."""
    lib = load()
    for t in (X, Y):     # row-major views (column slices of a wider buffer are fine)
        if not t.is_cuda or t.dtype != torch.float16 or t.dim() != 2 or t.stride(1) != 1:
            raise OnssenB200Error("gemm_f16_rows operands must be 2-D fp16 CUDA tensors with unit column stride")
    rc = lib.onssen_gemm_f16_rows(_p(X), _p(Y), _p(out), M, N, Kc, X.stride(0), Y.stride(0), out.stride(0), int(y_row_shift),
                                  _p(out_scale), _stream())
    _check(rc, "onssen_gemm_f16")
    return out


def amax_scale(x, target=1024.0):
    lib = load()
    scratch = torch.empty(1, device=x.device, dtype=torch.int32)
    scale2 = torch.empty(2, device=x.device, dtype=torch.float32)
    _check(lib.onssen_amax_scale(_p(_req(x, torch.float32)), x.numel(), float(target), _p(scratch), _p(scale2),
                                 _stream()), "onssen_amax_scale")
    return scale2


def cast_transpose_f16(src, scale2, want_n=True, want_t=True):
    """src fp32 [R][C] -> (fp16 [R][pad64(C)] or None, fp16 [C][pad64(R)] or None), both times scale2[0]."""
    lib = load()
    R, C = src.shape
    Cp, Rp = pad64(C), pad64(R)
    out_n = torch.empty(R, Cp, device=src.device, dtype=torch.float16) if want_n else None
    out_t = torch.empty(C, Rp, device=src.device, dtype=torch.float16) if want_t else None
    rc = lib.onssen_cast_transpose_f16(_p(_req(src, torch.float32)), R, C, src.stride(0), _p(scale2), _p(out_n), Cp,
                                       _p(out_t), Rp, _stream())
    _check(rc, "onssen_cast_transpose_f16")
    return out_n, out_t


def transpose_shift_f16(src, col0, ncol, shift=0):
    lib = load()
    R = src.shape[0]
    Rp = pad64(R)
    out = torch.empty(ncol, Rp, device=src.device, dtype=torch.float16)
    rc = lib.onssen_transpose_shift_f16(_p(_req(src, torch.float16)), R, src.stride(0), col0, ncol, shift, _p(out), Rp,
                                        _stream())
    _check(rc, "onssen_transpose_shift_f16")
    return out


def colsum(x, mult=1.0):
    lib = load()
    R, C = x.shape
    out = torch.empty(C, device=x.device, dtype=torch.float32)
    scratch = torch.empty(lib.onssen_colsum_scratch_bytes(R, C), device=x.device, dtype=torch.uint8)
    _check(lib.onssen_colsum(_p(_req(x, torch.float32)), R, C, x.stride(0), float(mult), _p(out), _p(scratch), _stream()),
           "onssen_colsum")
    return out


def normalize_bwd(d_emb, emb, inv_norm):
    lib = load()
    B, T, F, D = emb.shape
    dz = torch.empty(T * B, F * D, device=emb.device, dtype=torch.float32)
    amax = torch.empty(1, device=emb.device, dtype=torch.int32)
    rc = lib.onssen_normalize_bwd(_p(_req(d_emb, torch.float32)), _p(_req(emb, torch.float32)),
                                  _p(_req(inv_norm, torch.float32)), B, T, F, D, _p(dz), _p(amax), _stream())
    _check(rc, "onssen_normalize_bwd")
    scale2 = torch.empty(2, device=emb.device, dtype=torch.float32)
    _check(lib.onssen_scale_from_amax_bits(_p(amax), 1024.0, _p(scale2), _stream()), "onssen_scale_from_amax_bits")
    return dz, scale2


def bn_backward(d_out, y, M, H, gamma, save_mean, save_invstd, sync=None):
    lib = load()
    d_y = torch.empty_like(y)
    dg = torch.empty(2 * H, device=y.device, dtype=torch.float32)
    db = torch.empty(2 * H, device=y.device, dtype=torch.float32)
    scratch = torch.empty(lib.onssen_bn_backward_scratch_bytes(M, H), device=y.device, dtype=torch.uint8)
    if sync is not None:
        reduce_sum, world = sync
        sums = torch.empty(2 * 2 * hp_of(H), device=y.device, dtype=torch.float64)
        _check(lib.onssen_bn_backward_stats(_p(_req(d_out, torch.float32)), _p(_req(y, torch.float32)), M, H,
                                            _p(save_mean), _p(save_invstd), _p(sums), _p(scratch), _stream()),
               "onssen_bn_backward_stats")
        reduce_sum(sums)
        rc = lib.onssen_bn_backward_apply(_p(d_out), _p(y), M, M * world, H, _p(gamma), _p(save_mean), _p(save_invstd),
                                          _p(sums), _p(d_y), _p(dg), _p(db), _p(scratch), _stream())
        _check(rc, "onssen_bn_backward_apply")
        return d_y, dg, db
    rc = lib.onssen_bn_backward(_p(_req(d_out, torch.float32)), _p(_req(y, torch.float32)), M, H, _p(gamma),
                                _p(save_mean), _p(save_invstd), _p(d_y), _p(dg), _p(db), _p(scratch), _stream())
    _check(rc, "onssen_bn_backward")
    return d_y, dg, db


def unpack_linear_grad(gp, N, K, in_is_blstm, Hin=0):
    lib = load()
    g = torch.empty(N, K, device=gp.device, dtype=torch.float32)
    _check(lib.onssen_unpack_linear_grad(_p(_req(gp, torch.float32)), N, K, int(in_is_blstm), Hin, gp.stride(0), _p(g),
                                         _stream()), "onssen_unpack_linear_grad")
    return g


def unpack_lstm_grad(gp, H, K, in_is_blstm, Hin, direction, Kp=None):
    lib = load()
    g = torch.empty(4 * H, K, device=gp.device, dtype=torch.float32)
    kp = gp.stride(0) if Kp is None else Kp
    _check(lib.onssen_unpack_lstm_grad(_p(gp), H, K, int(in_is_blstm), Hin, kp, direction, _p(g), _stream()),
           "onssen_unpack_lstm_grad")
    return g


def lstm_pack_whh_t(w_hh_f, w_hh_r, H):
    lib = load()
    Hp = hp_of(H)
    out = torch.empty(lib.onssen_lstm_pack_whh_t_elems(H), device=w_hh_f.device, dtype=torch.float16)
    _check(lib.onssen_lstm_pack_whh_t(_p(_req(w_hh_f.detach(), torch.float32)), _p(_req(w_hh_r.detach(), torch.float32)),
                                      H, _p(out), _stream()), "onssen_lstm_pack_whh_t")
    return out


def blstm_rec_fwd_train(gates, whh_p, B, T, H, y_h, y_f, c_out, h_raw, dropout_p, seed, offset, workspace):
    lib = load()
    rc = lib.onssen_blstm_rec_fwd_train(_p(_req(gates, torch.float32)), _p(whh_p), B, T, H, _p(y_h), _p(y_f), _p(c_out),
                                        _p(h_raw), float(dropout_p), int(seed), int(offset), _p(workspace),
                                        workspace.numel(), _stream())
    _check(rc, "onssen_blstm_rec_fwd")


_BPTT_SAT = {}     # device -> int32[1]: clamped dG values of the persistent BPTT kernel since the last read


def bptt_saturation_count(device=None, reset=True):
    """Number of BPTT exchange values that were clamped into the flag range since the last call (device -> host read:
    call it at a logging interval, not every step).  Non-zero means the recurrent gradient exceeded 32x the largest
    output gradient (or was NaN) and the step's gradient is not trustworthy."""
    total = 0
    for dev, t in _BPTT_SAT.items():
        if device is None or torch.device(device) == dev:
            total += int(t.item())
            if reset:
                t.zero_()
    return total


def blstm_rec_bwd(act_gates, dg16, c, dy, whh_t, scale2, B, T, H, dropout_p, seed, offset):
    lib = load()
    scratch = torch.empty(lib.onssen_blstm_rec_bwd_scratch_bytes(B, H), device=c.device, dtype=torch.uint8)
    sat = _BPTT_SAT.get(c.device)
    if sat is None:
        sat = _BPTT_SAT[c.device] = torch.zeros(1, device=c.device, dtype=torch.int32)
    rc = lib.onssen_blstm_rec_bwd(_p(act_gates), _p(dg16), _p(c), _p(_req(dy, torch.float32)), _p(whh_t), _p(scratch),
                                  _p(scale2), _p(sat), B, T, H, float(dropout_p), int(seed), int(offset), _stream())
    _check(rc, "onssen_blstm_rec_bwd")


def loss_pit_l1_bwd(mask_a, mask_b, mask_stride, mag_mix, mag_s1, mag_s2, cos_s1, cos_s2, perm, g):
    lib = load()
    B = mag_mix.shape[0]
    N = mag_mix.numel() // B
    da = torch.empty(mag_mix.shape, device=mag_mix.device, dtype=torch.float32)
    db = torch.empty(mag_mix.shape, device=mag_mix.device, dtype=torch.float32)
    rc = lib.onssen_loss_pit_l1_bwd(_p(mask_a), _p(mask_b), mask_stride, _p(mag_mix), _p(mag_s1), _p(mag_s2), _p(cos_s1),
                                    _p(cos_s2), _p(perm), _p(_req(g.float().contiguous(), torch.float32)), B, N, _p(da),
                                    _p(db), _stream())
    _check(rc, "onssen_loss_pit_l1_fwd")
    return da, db


def sigmoid_bwd(d_out, out):
    lib = load()
    B, T = out.shape[0], out.shape[1]
    C = out.numel() // (B * T)
    dz = torch.empty(T * B, C, device=out.device, dtype=torch.float32)
    amax = torch.empty(1, device=out.device, dtype=torch.int32)
    _check(lib.onssen_sigmoid_bwd(_p(_req(d_out, torch.float32)), _p(_req(out, torch.float32)), B, T, C, _p(dz), _p(amax),
                                  _stream()), "onssen_sigmoid_bwd")
    scale2 = torch.empty(2, device=out.device, dtype=torch.float32)
    _check(lib.onssen_scale_from_amax_bits(_p(amax), 1024.0, _p(scale2), _stream()), "onssen_scale_from_amax_bits")
    return dz, scale2


def add_inplace(a, b):
    _check(load().onssen_add_inplace(_p(_req(a, torch.float32)), _p(_req(b, torch.float32)), a.numel(), _stream()),
           "onssen_add_inplace")
    return a


def relu_bwd(d_out, out):
    lib = load()
    B, T = out.shape[0], out.shape[1]
    C = out.numel() // (B * T)
    dz = torch.empty(T * B, C, device=out.device, dtype=torch.float32)
    amax = torch.empty(1, device=out.device, dtype=torch.int32)
    _check(lib.onssen_relu_bwd(_p(_req(d_out, torch.float32)), _p(_req(out, torch.float32)), B, T, C, _p(dz), _p(amax),
                               _stream()), "onssen_sigmoid_bwd")
    scale2 = torch.empty(2, device=out.device, dtype=torch.float32)
    _check(lib.onssen_scale_from_amax_bits(_p(amax), 1024.0, _p(scale2), _stream()), "onssen_scale_from_amax_bits")
    return dz, scale2


def enhance_mid_bwd(d_est, pre, mask):
    lib = load()
    M, F = pre.shape
    dz_pre, dz_mi = torch.empty_like(pre), torch.empty_like(pre)
    amax = torch.empty(2, device=pre.device, dtype=torch.int32)
    rc = lib.onssen_enhance_mid_bwd(_p(_req(d_est, torch.float32)), d_est.stride(0), _p(_req(pre, torch.float32)),
                                    _p(_req(mask, torch.float32)), M, F, _p(dz_pre), _p(dz_mi), _p(amax), _stream())
    _check(rc, "onssen_enhance_mid_bwd")
    scs = []
    for k in range(2):
        sc = torch.empty(2, device=pre.device, dtype=torch.float32)
        _check(lib.onssen_scale_from_amax_bits(_p(amax[k:]), 1024.0, _p(sc), _stream()), "onssen_scale_from_amax_bits")
        scs.append(sc)
    return dz_pre, scs[0], dz_mi, scs[1]


def loss_mse_bwd(a, b, g):
    lib = load()
    d_a = torch.empty_like(a)
    _check(lib.onssen_loss_mse_bwd(_p(_req(a, torch.float32)), _p(_req(b, torch.float32)), a.numel(),
                                   _p(_req(g.float().reshape(1).contiguous(), torch.float32)), _p(d_a), _stream()),
           "onssen_loss_mse_fwd")
    return d_a


def loss_phase_cos_bwd(pa, pb, s1, s2, mag, perm, g):
    lib = load()
    B = mag.shape[0]
    N = mag.numel() // B
    d_pa, d_pb = torch.empty_like(pa), torch.empty_like(pb)
    rc = lib.onssen_loss_phase_cos_bwd(*[_p(_req(t, torch.float32)) for t in (pa, pb, s1, s2, mag)],
                                       _p(_req(perm, torch.int32)), _p(_req(g.float().contiguous(), torch.float32)), B, N,
                                       _p(d_pa), _p(d_pb), _stream())
    _check(rc, "onssen_loss_phase_cos_bwd")
    return d_pa, d_pb


def l2norm_pairs_bwd(d_y, x, residual):
    """-> (dz fp32 [T*B][2F] time-major, scale2)"""
    lib = load()
    B, T, F, _ = x.shape
    dz = torch.empty(T * B, 2 * F, device=x.device, dtype=torch.float32)
    amax = torch.empty(1, device=x.device, dtype=torch.int32)
    _check(lib.onssen_l2norm_pairs_bwd(_p(_req(d_y, torch.float32)), _p(_req(x, torch.float32)),
                                       _p(_req(residual, torch.float32)), B, T, F, _p(dz), _p(amax), _stream()),
           "onssen_l2norm_pairs_bwd")
    scale2 = torch.empty(2, device=x.device, dtype=torch.float32)
    _check(lib.onssen_scale_from_amax_bits(_p(amax), 1024.0, _p(scale2), _stream()), "onssen_scale_from_amax_bits")
    return dz, scale2


def phase_input_bwd(d_xin, x_mag, d_masks, s_idx):
    lib = load()
    B, T, F = x_mag.shape
    S = d_masks.shape[-1]
    _check(lib.onssen_phase_input_bwd(_p(_req(d_xin, torch.float32)), d_xin.stride(0), _p(_req(x_mag, torch.float32)), B, T,
                                      F, S, s_idx, _p(_req(d_masks, torch.float32)), _stream()), "onssen_phase_input_bwd")
    return d_masks


def loss_l1_psa_bwd(mask, noisy, clean, cosd, g):
    lib = load()
    B = noisy.shape[0]
    d_mask = torch.empty_like(mask)
    rc = lib.onssen_loss_l1_psa_bwd(*[_p(_req(t, torch.float32)) for t in (mask, noisy, clean, cosd)],
                                    _p(_req(g.float().contiguous(), torch.float32)), B, noisy.numel() // B, _p(d_mask),
                                    _stream())
    _check(rc, "onssen_loss_l1_psa_bwd")
    return d_mask


# ------------------------------------------------------------------------------------------------ optimiser step
OPT_CHUNK = 65536


class OptTables:
    """Device tables for the multi-tensor optimiser kernels.  The chunk list only depends on the tensor sizes (built
    once); the pointer records are rebuilt per call (gradient tensors are new objects after every backward)."""

    def __init__(self, numels, device):
        import numpy as np
        self.numels = list(numels)
        ch = [(i, s) for i, n in enumerate(self.numels) for s in range(0, n, OPT_CHUNK)]
        self.nchunks = len(ch)
        self.chunks = torch.from_numpy(np.array(ch, dtype=np.int64).reshape(-1, 2)).to(device)
        self.partials = torch.empty(self.nchunks, device=device, dtype=torch.float64)
        self.dev = torch.empty(len(self.numels), 5, device=device, dtype=torch.int64)

    def fill(self, rows):
        """rows: sequence of (param_ptr, grad_ptr, m_ptr, v_ptr) ints; numel comes from the constructor.  The source of
        the copy is a fresh PAGEABLE array: the runtime stages pageable memory before cudaMemcpyAsync returns, so
        nothing has to outlive this call although the loop never synchronises."""
        import numpy as np
        arr = np.empty((len(self.numels), 5), dtype=np.int64)
        arr[:, :4] = np.asarray(rows, dtype=np.int64).reshape(len(self.numels), 4)
        arr[:, 4] = self.numels
        self.dev.copy_(torch.from_numpy(arr), non_blocking=True)
        return self.dev


def clip_grad_norm(tables, rows, max_norm, out2):
    lib = load()
    dev = tables.fill(rows)
    _check(lib.onssen_clip_grad_norm(_p(dev), _p(tables.chunks), tables.nchunks, OPT_CHUNK, float(max_norm),
                                     _p(tables.partials), _p(out2), _stream()), "onssen_clip_grad_norm")
    return out2


def adam_step(tables, rows, lr, beta1, beta2, eps, weight_decay, step):
    lib = load()
    dev = tables.fill(rows)
    _check(lib.onssen_adam_step(_p(dev), _p(tables.chunks), tables.nchunks, OPT_CHUNK, float(lr), float(beta1),
                                float(beta2), float(eps), float(weight_decay), int(step), _stream()), "onssen_adam_step")


def kmeans_masks(emb, feature, K=2, db_threshold=40.0, iters=30, want_labels=False):
    """emb (frames, F, D) or (N, D) of one utterance, feature (frames, F) or None -> masks (K, frames, F)
    [, labels int32 (frames, F)]"""
    lib = load()
    D = emb.shape[-1]
    N = emb.numel() // D
    shape = tuple(emb.shape[:-1])
    masks = torch.empty((K,) + shape, device=emb.device, dtype=torch.float32)
    labels = torch.empty(shape, device=emb.device, dtype=torch.int32) if want_labels else None
    scratch = torch.empty(lib.onssen_kmeans_scratch_bytes(), device=emb.device, dtype=torch.uint8)
    rc = lib.onssen_kmeans_masks(_p(_req(emb, torch.float32)), _p(None if feature is None else _req(feature, torch.float32)),
                                 N, D, K, float(db_threshold), int(iters), _p(masks), _p(labels), _p(scratch), _stream())
    _check(rc, "onssen_kmeans_masks")
    return (masks, labels) if want_labels else masks


def pcm16_to_f32(pcm, channels, frames):
    """pcm int16 (R, pitch*channels) CUDA, frames int32 (R,) CUDA -> float32 (R, pitch) mono (sample/32768)"""
    lib = load()
    R = pcm.shape[0]
    pitch = pcm.shape[1] // channels
    out = torch.empty(R, pitch, device=pcm.device, dtype=torch.float32)
    _check(lib.onssen_pcm16_to_f32(_p(_req(pcm, torch.int16, "pcm")), R, pitch, int(channels),
                                   _p(_req(frames, torch.int32, "frames")), _p(out), _stream()), "onssen_pcm16_to_f32")
    return out


def resample_poly(x, n_in, max_n_in, up, down, h, n_pre_remove):
    """x float32 (R, pitch_in) CUDA, n_in int32 (R,) CUDA; h float32 filter on the device -> (y (R, pitch_out), pitch_out)"""
    lib = load()
    R, pitch_in = x.shape
    pitch_out = (int(max_n_in) * up + down - 1) // down
    y = torch.empty(R, pitch_out, device=x.device, dtype=torch.float32)
    _check(lib.onssen_resample_poly(_p(_req(x, torch.float32)), R, pitch_in, _p(_req(n_in, torch.int32)), int(up),
                                    int(down), _p(_req(h, torch.float32)), h.numel(), int(n_pre_remove), _p(y), pitch_out,
                                    _stream()), "onssen_resample_poly")
    return y, pitch_out
