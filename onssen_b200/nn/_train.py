"""Training path of the deep-clustering stack: forward that saves the BPTT state and a hand-written backward
(no autograd graph of torch ops; every step below is a C-ABI call into libonssen_b200.so).

Replaces what `loss_avg.backward()` (/root/reference/onssen/utils/train.py:82) gets from cuDNN/cuBLAS autograd
for deep_clustering.py:34-42: F.normalize backward -> head dgrad/wgrad (tcgen05 GEMMs on scaled fp16 copies) ->
BatchNorm backward -> per layer BPTT (one launch per step) + W_ih / W_hh / bias gradients (tcgen05 GEMMs
contracting over all (t,b)) + dgrad to the layer below."""
import torch

from .. import _lib
from ._blstm import _next_seed, lstm_layer_params, pack_lstm


def dc_forward_train(model, x):
    rnn, bn = model.rnn, model.bn
    B, T, F = x.shape
    H, L, D = rnn.hidden_size, rnn.num_layers, model.embedding_dim
    Hp, M = _lib.hp_of(H), T * B
    params = [p for l in range(L) for d in lstm_layer_params(rnn, l) for p in d]
    packed = model._rnn_cache.get(params, lambda: pack_lstm(rnn))
    a = _lib.pack_input_f16(x.contiguous())
    ws = _lib.blstm_rec_workspace(B, H, x.device)
    layers, y_f = [], None
    for l in range(L):
        wih_p, whh_p, bias_p = packed[l]
        gates = torch.empty(M, 8 * Hp, device=x.device, dtype=torch.float32)
        _lib.gemm_f16(a, wih_p, bias_p, gates, M, 8 * Hp, a.shape[1], 8 * Hp)
        last = l == L - 1
        p = float(rnn.dropout) if (model.training and not last) else 0.0
        seed = _next_seed() if p > 0 else 0
        y_h = None if last else torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float16)
        y_f = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float32) if last else None
        c = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float32)
        h_raw = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float16) if (last or p > 0) else None
        _lib.blstm_rec_fwd_train(gates, whh_p, B, T, H, y_h, y_f, c, h_raw, p, seed, l, ws)
        layers.append(dict(a_in=a, gates=gates, c=c, h16=h_raw if h_raw is not None else y_h, p=p, seed=seed))
        a = y_h
    a_h, mean, invstd = _lib.bn_forward_f16(y_f, M, H, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                            bn.running_var, bn.eps, bn.momentum, True, save_stats=True)
    bn.num_batches_tracked += 1
    w_p = model._fc_cache.get([model.fc_dc.weight], lambda: _lib.pack_linear_f16(model.fc_dc.weight, True, H))
    emb = torch.empty(B, T, F, D, device=x.device, dtype=torch.float32)
    inv_norm = torch.empty(B, T, F, device=x.device, dtype=torch.float32)
    _lib.gemm_f16_ex(a_h, w_p, model.fc_dc.bias.detach(), emb, M, F * D, a_h.shape[1], F * D, epi=3, group=D,
                     remap_inner=B, remap_outer=T, inv_norm=inv_norm)
    saved = dict(layers=layers, y_f=y_f, a_h=a_h, mean=mean, invstd=invstd, emb=emb, inv_norm=inv_norm, packed=packed,
                 w_p=w_p, shape=(B, T, F))
    return emb, saved


def dc_backward(model, saved, d_emb, on_grads=None):
    """on_grads(dict) is called with each group of gradients as soon as it is final (head+BN, then every BLSTM
    layer from the top): the hook point for the overlapped data-parallel all-reduce (utils/ddp.py)."""
    rnn, bn = model.rnn, model.bn
    B, T, F = saved["shape"]
    H, L, D = rnn.hidden_size, rnn.num_layers, model.embedding_dim
    Hp, M, N = _lib.hp_of(H), T * B, F * D
    Mp = _lib.pad64(M)
    dev = d_emb.device
    grads = {}
    f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)

    # ---- F.normalize + fc_dc
    dz32, sc = _lib.normalize_bwd(d_emb, saved["emb"], saved["inv_norm"])
    dz_n, dz_t = _lib.cast_transpose_f16(dz32, sc)
    grads["fc_dc.bias"] = _lib.colsum(dz32)
    ahT = _lib.transpose_shift_f16(saved["a_h"], 0, 2 * Hp)
    dWp = f32(N, 2 * Hp)
    _lib.gemm_f16_ex(dz_t, ahT, None, dWp, N, 2 * Hp, Mp, 2 * Hp, out_scale=sc[1:])
    grads["fc_dc.weight"] = _lib.unpack_linear_grad(dWp, N, 2 * H, True, H)
    wT = _lib.transpose_shift_f16(saved["w_p"], 0, 2 * Hp)
    dA = f32(M, 2 * Hp)
    _lib.gemm_f16_ex(dz_n, wT, None, dA, M, 2 * Hp, dz_n.shape[1], 2 * Hp, out_scale=sc[1:])
    del dz32, dz_n, dz_t
    # ---- BatchNorm1d
    dY, grads["bn.weight"], grads["bn.bias"] = _lib.bn_backward(dA, saved["y_f"], M, H, bn.weight.detach(),
                                                                saved["mean"], saved["invstd"])
    if on_grads is not None:
        on_grads(dict(grads))
    # ---- BLSTM, top layer first
    for l in reversed(range(L)):
        done_before = set(grads)
        lay = saved["layers"][l]
        wih_p, _, _ = saved["packed"][l]
        (wf, wr) = lstm_layer_params(rnn, l)
        sc = _lib.amax_scale(dY)
        dg16 = torch.empty(M, 8 * Hp, device=dev, dtype=torch.float16)
        whh_t = _lib.lstm_pack_whh_t(wf[1], wr[1], H)
        _lib.blstm_rec_bwd(lay["gates"], dg16, lay["c"], dY, whh_t, sc, B, T, H, lay["p"], lay["seed"], l)
        dG32 = lay["gates"]                      # now the fp32 pre-activation gradients
        inv = sc[1:]
        gb = _lib.colsum(dG32)                   # [8Hp], permuted
        kp_in = lay["a_in"].shape[1]
        I_l = rnn.input_size if l == 0 else 2 * H
        dgT = _lib.transpose_shift_f16(dg16, 0, 8 * Hp)
        xT = _lib.transpose_shift_f16(lay["a_in"], 0, kp_in)
        dWih_p = f32(8 * Hp, kp_in)
        _lib.gemm_f16_ex(dgT, xT, None, dWih_p, 8 * Hp, kp_in, Mp, kp_in, out_scale=inv)
        for d, suf in enumerate(("", "_reverse")):
            grads[f"rnn.weight_ih_l{l}{suf}"] = _lib.unpack_lstm_grad(dWih_p, H, I_l, l > 0, H if l > 0 else 0, d)
            gbd = _lib.unpack_lstm_grad(gb, H, 1, False, 0, d, Kp=1).view(4 * H)
            grads[f"rnn.bias_ih_l{l}{suf}"] = gbd
            grads[f"rnn.bias_hh_l{l}{suf}"] = gbd
            hT = _lib.transpose_shift_f16(lay["h16"], d * Hp, Hp, shift=B if d == 0 else -B)
            dWhh_p = f32(4 * Hp, Hp)
            _lib.gemm_f16_ex(dgT[d * 4 * Hp:(d + 1) * 4 * Hp], hT, None, dWhh_p, 4 * Hp, Hp, Mp, Hp, out_scale=inv)
            grads[f"rnn.weight_hh_l{l}{suf}"] = _lib.unpack_lstm_grad(dWhh_p, H, H, False, 0, 0)
        if l > 0:
            wihT = _lib.transpose_shift_f16(wih_p, 0, kp_in)
            dX = f32(M, kp_in)
            _lib.gemm_f16_ex(dg16, wihT, None, dX, M, kp_in, 8 * Hp, kp_in, out_scale=inv)
            dY = dX
        lay["gates"] = None
        if on_grads is not None:
            # bias_ih / bias_hh share one tensor: reduce it once
            on_grads({k: v for k, v in grads.items() if k not in done_before and "bias_hh" not in k})
    return grads


class DCFunction(torch.autograd.Function):
    """emb = deep_clustering(x) with a hand-written backward; `params` only carries the autograd edges."""

    @staticmethod
    def forward(ctx, model, x, *params):
        emb, saved = dc_forward_train(model, x)
        ctx.model, ctx.saved = model, saved
        ctx.names = [n for n, _ in model.named_parameters()]
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        sync = getattr(ctx.model, "grad_sync", None)
        grads = dc_backward(ctx.model, ctx.saved, d_emb.contiguous(), None if sync is None else sync.reduce_bucket)
        if sync is not None:
            sync.wait()
        ctx.saved = None
        return (None, None) + tuple(grads[n] for n in ctx.names)
