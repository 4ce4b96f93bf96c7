"""Host-side driver of the BLSTM stack: packing cache + per-layer (projection GEMM -> persistent recurrence).

Replaces the `self.rnn(x)` call of the reference models (deep_clustering.py:34-35, chimera.py:35-36,
enhancement.py:43-44, phase_network.py:50,57).  Parameters stay in a `torch.nn.LSTM` container so that
state_dict keys (`rnn.weight_ih_l0`, ... ) and default initialisation are identical to the reference; the
container's own forward is never called.
"""
import torch

from .. import _lib

# Largest batch one recurrent launch supports is device dependent (32 columns per CTA x slices that fit).


_seed_counter = [0]


def _next_seed():
    # host-side counter mixed with torch's seed: no device sync (dropout masks are statistical-parity only)
    _seed_counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seed_counter[0] * 0xD1B54A32D192ED03) % (1 << 63)


def _versions(params):
    return tuple((p.data_ptr(), p._version) for p in params)


class PackCache:
    """fp16 packed copies of weights, rebuilt when any source parameter changes (optimizer steps bump
    `_version`)."""

    def __init__(self):
        self.key = None
        self.val = None

    def get(self, params, build):
        key = _versions(params)
        if key != self.key:
            self.val = build()
            self.key = key
        return self.val


def lstm_layer_params(rnn, layer):
    sufs = ("", "_reverse")
    return [tuple(getattr(rnn, f"{n}_l{layer}{s}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
            for s in sufs]


def pack_lstm(rnn):
    H, L, I = rnn.hidden_size, rnn.num_layers, rnn.input_size
    packed = []
    for l in range(L):
        wf, wr = lstm_layer_params(rnn, l)
        packed.append(_lib.lstm_pack_layer(wf, wr, H, I if l == 0 else 2 * H, l > 0, H if l > 0 else 0))
    return packed


def bn_sync(model):
    """(all_reduce_sum, world_size) when `model.sync_bn` is set and a process group with more than one rank is up
    (utils.ddp.enable_sync_batchnorm), else None: train-mode BatchNorm statistics then cover the whole
    data-parallel batch, which is the single-device arithmetic of deep_clustering.py:36-38."""
    import torch.distributed as dist
    if not getattr(model, "sync_bn", False) or not (dist.is_available() and dist.is_initialized()):
        return None
    world = dist.get_world_size()
    if world == 1:
        return None
    return (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)), world


def require_no_grad(module_name, *tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            f"{module_name}: the backward kernels are not part of this build yet; call under torch.no_grad() "
            "(there is deliberately no autograd/PyTorch fallback)")


def blstm_forward(rnn, cache, x, training, want_f32, want_f16, use_tensor_cores=True, lengths=None):
    """x (B,T,I) fp32 CUDA. Returns (y_h fp16 [T*B][2Hp] or None, y_f fp32 [T*B][2Hp] or None).
    lengths: optional int32 CUDA tensor (B,) of frames per utterance (zero-padded inference batch)."""
    B, T, _ = x.shape
    return blstm_forward_packed(rnn, cache, _lib.pack_input_f16(x.contiguous()), B, T, training, want_f32,
                                want_f16, use_tensor_cores, lengths)


def eval_lengths(model, B, device):
    """frames per utterance set by the batched evaluation loop (`model.frame_lengths`, utils/test.py), or None"""
    lens = getattr(model, "frame_lengths", None)
    if lens is None:
        return None
    if model.training:
        raise _lib.OnssenB200Error("frame_lengths (padded batches) are supported in eval() mode only")
    lens = torch.as_tensor(lens, dtype=torch.int32, device=device)
    assert lens.numel() == B, "frame_lengths must have one entry per utterance"
    return lens


def blstm_forward_packed(rnn, cache, a, B, T, training, want_f32, want_f16, use_tensor_cores=True, lengths=None):
    """Same, from an already packed time-major fp16 input a [T*B][Kp]."""
    assert rnn.bidirectional and rnn.batch_first and rnn.proj_size == 0
    H, L = rnn.hidden_size, rnn.num_layers
    Hp = _lib.hp_of(H)
    M = T * B
    x = a
    params = [p for l in range(L) for d in lstm_layer_params(rnn, l) for p in d]
    packed = cache.get(params, lambda: pack_lstm(rnn))
    gates = torch.empty(M, 8 * Hp, device=x.device, dtype=torch.float32)
    ws = _lib.blstm_rec_workspace(B, H, x.device)
    y_h = y_f = None
    for l in range(L):
        wih_p, whh_p, bias_p = packed[l]
        _lib.gemm_f16(a, wih_p, bias_p, gates, M, 8 * Hp, a.shape[1], 8 * Hp)
        last = l == L - 1
        y_h = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float16) if (not last or want_f16) else None
        y_f = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float32) if (last and want_f32) else None
        p = float(rnn.dropout) if (training and not last) else 0.0
        seed = _next_seed() if p > 0 else 0
        _lib.blstm_rec_fwd(gates, whh_p, B, T, H, y_h, y_f, p, seed, l, ws, use_tensor_cores, col_len=lengths)
        a = y_h
    return y_h, y_f
