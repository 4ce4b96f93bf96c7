"""Training path of the BLSTM models: forward passes that save the BPTT state and hand-written backwards
(no autograd graph of torch ops; every step below is a C-ABI call into libonssen_b200.so).

Replaces what `loss_avg.backward()` (/root/reference/onssen/utils/train.py:82) gets from cuDNN/cuBLAS autograd
for deep_clustering.py:34-42 and chimera.py:35-45: F.normalize / sigmoid backward -> head dgrad/wgrad (tcgen05
GEMMs on scaled fp16 copies) -> BatchNorm backward -> per layer BPTT (ONE persistent tcgen05 launch per layer, both
directions, all T steps: csrc/lstm_bwd_tc.cu) + W_ih / W_hh / bias gradients (rows-mode tcgen05 GEMMs contracting over
all (t,b), operands read in place) + dgrad to the layer below."""
import torch

from .. import _lib
from ._blstm import _next_seed, bn_sync, lstm_layer_params, pack_lstm


# ------------------------------------------------------------------------------------------------ shared pieces
def blstm_forward_train(rnn, cache, x, training, last_f32):
    """Returns (layers, packed, y_last) where y_last is fp32 [M][2Hp] (last_f32) or fp16; layers hold the BPTT state."""
    B, T, _ = x.shape
    return blstm_forward_train_packed(rnn, cache, _lib.pack_input_f16(x.contiguous()), B, T, training, last_f32)


def blstm_forward_train_packed(rnn, cache, a, B, T, training, last_f32):
    """Same, from an already packed time-major fp16 input a [T*B][Kp]."""
    x = a
    H, L = rnn.hidden_size, rnn.num_layers
    Hp, M = _lib.hp_of(H), T * B
    params = [p for l in range(L) for d in lstm_layer_params(rnn, l) for p in d]
    packed = cache.get(params, lambda: pack_lstm(rnn))
    ws = _lib.blstm_rec_workspace(B, H, x.device)
    layers, y_h, y_f = [], None, None
    for l in range(L):
        wih_p, whh_p, bias_p = packed[l]
        gates = torch.empty(M, 8 * Hp, device=x.device, dtype=torch.float32)
        _lib.gemm_f16(a, wih_p, bias_p, gates, M, 8 * Hp, a.shape[1], 8 * Hp)
        last = l == L - 1
        p = float(rnn.dropout) if (training and not last) else 0.0
        seed = _next_seed() if p > 0 else 0
        y_h = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float16) if (not last or not last_f32) else None
        y_f = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float32) if (last and last_f32) else None
        c = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float32)
        need_raw = p > 0 or y_h is None
        h_raw = torch.empty(M, 2 * Hp, device=x.device, dtype=torch.float16) if need_raw else None
        _lib.blstm_rec_fwd_train(gates, whh_p, B, T, H, y_h, y_f, c, h_raw, p, seed, l, ws)
        layers.append(dict(a_in=a, gates=gates, c=c, h16=h_raw if h_raw is not None else y_h, p=p, seed=seed))
        a = y_h
    return layers, packed, (y_f if last_f32 else y_h)


def linear_backward(dz32, sc, a_in_h, w_p, N, H, grads, name, K=None, want_dA=True):
    """Linear on packed operands: dz32 [M][N] fp32 (time-major), sc = scale2 of dz32, a_in_h fp16 [M][Kp] the layer's
    input, w_p fp16 [N][Kp].  K is None: the input is a BLSTM output in the padded layout (K = 2H, Kp = 2Hp); else a
    plain K-feature input padded to Kp.  Fills grads[name.weight/bias]; returns dA [M][Kp] (or None)."""
    M = dz32.shape[0]
    Kp, Mp = a_in_h.shape[1], _lib.pad64(M)
    # dW = dz^T a contracts over the M rows of two row-major buffers: read in place (MN-major tcgen05 operands)
    dz_n, _ = _lib.cast_transpose_f16(dz32, sc, want_n=True, want_t=False)
    grads[name + ".bias"] = _lib.colsum(dz32)
    dWp = torch.empty(N, Kp, device=dz32.device, dtype=torch.float32)
    _lib.gemm_f16_rows(dz_n[:, :N], a_in_h, dWp, N, Kp, M, out_scale=sc[1:])
    grads[name + ".weight"] = (_lib.unpack_linear_grad(dWp, N, 2 * H, True, H) if K is None
                               else _lib.unpack_linear_grad(dWp, N, K, False, 0))
    if not want_dA:
        return None
    wT = _lib.transpose_shift_f16(w_p, 0, Kp)
    dA = torch.empty(M, Kp, device=dz32.device, dtype=torch.float32)
    _lib.gemm_f16_ex(dz_n, wT, None, dA, M, Kp, dz_n.shape[1], Kp, out_scale=sc[1:])
    return dA


def blstm_backward(rnn, layers, packed, dY, B, T, grads, prefix, on_grads=None, want_dx0=False):
    """BPTT through the stack.  Returns grads, or (grads, dX0 fp32 [M][Kp of layer 0]) with want_dx0."""
    H, L = rnn.hidden_size, rnn.num_layers
    Hp, M = _lib.hp_of(H), T * B
    Mp = _lib.pad64(M)
    dev = dY.device
    for l in reversed(range(L)):
        done_before = set(grads)
        lay = layers[l]
        wih_p = packed[l][0]
        (wf, wr) = lstm_layer_params(rnn, l)
        # scale so that amax(dY)*scale = 2^-4: |dG*scale| stays below 2 (bit 14 of the fp16 exchange words is the
        # step-parity flag of the persistent BPTT kernel) with 32x headroom for the cell-gradient accumulation
        sc = _lib.amax_scale(dY, target=0.0625)
        dg16 = torch.empty(M, 8 * Hp, device=dev, dtype=torch.float16)
        whh_t = _lib.lstm_pack_whh_t(wf[1], wr[1], H)
        _lib.blstm_rec_bwd(lay["gates"], dg16, lay["c"], dY, whh_t, sc, B, T, H, lay["p"], lay["seed"], l)
        if l == 0 and on_grads is not None:
            # the last cooperative launch of the backward is enqueued: gradient buckets held back so far (GradSync
            # without overlap) can be all-reduced beside the weight-gradient GEMMs that follow
            flush = getattr(getattr(on_grads, "__self__", None), "flush", None)
            if flush is not None:
                flush()
        dG32 = lay["gates"]                      # now the fp32 pre-activation gradients
        inv = sc[1:]
        gb = _lib.colsum(dG32)                   # [8Hp], permuted
        kp_in = lay["a_in"].shape[1]
        I_l = rnn.input_size if l == 0 else 2 * H
        # weight gradients contract over the M = T*B rows of dG and of the layer's input / hidden state: both are read
        # in place as MN-major tcgen05 operands (round 1 made transposed fp16 copies: 6 % of the training step)
        dWih_p = torch.empty(8 * Hp, kp_in, device=dev, dtype=torch.float32)
        _lib.gemm_f16_rows(dg16, lay["a_in"], dWih_p, 8 * Hp, kp_in, M, out_scale=inv)
        for d, suf in enumerate(("", "_reverse")):
            grads[f"{prefix}weight_ih_l{l}{suf}"] = _lib.unpack_lstm_grad(dWih_p, H, I_l, l > 0, H if l > 0 else 0, d)
            gbd = _lib.unpack_lstm_grad(gb, H, 1, False, 0, d, Kp=1).view(4 * H)
            grads[f"{prefix}bias_ih_l{l}{suf}"] = gbd
            grads[f"{prefix}bias_hh_l{l}{suf}"] = gbd
            # dW_hh pairs dG_t with h_{t-1} (forward) / h_{t+1} (reverse): a shift of B rows, zero outside the sequence
            dWhh_p = torch.empty(4 * Hp, Hp, device=dev, dtype=torch.float32)
            _lib.gemm_f16_rows(dg16[:, d * 4 * Hp:(d + 1) * 4 * Hp], lay["h16"][:, d * Hp:(d + 1) * Hp], dWhh_p,
                               4 * Hp, Hp, M, y_row_shift=-B if d == 0 else B, out_scale=inv)
            grads[f"{prefix}weight_hh_l{l}{suf}"] = _lib.unpack_lstm_grad(dWhh_p, H, H, False, 0, 0)
        if l > 0 or want_dx0:
            wihT = _lib.transpose_shift_f16(wih_p, 0, kp_in)
            dX = torch.empty(M, kp_in, device=dev, dtype=torch.float32)
            _lib.gemm_f16_ex(dg16, wihT, None, dX, M, kp_in, 8 * Hp, kp_in, out_scale=inv)
            dY = dX
        lay["gates"] = None
        if on_grads is not None:
            # bias_ih / bias_hh share one tensor: reduce it once
            on_grads({k: v for k, v in grads.items() if k not in done_before and "bias_hh" not in k})
    return (grads, dY) if want_dx0 else grads


class _ModelFunction(torch.autograd.Function):
    """outputs = model(x) with a hand-written backward; `params` only carries the autograd edges."""

    @staticmethod
    def forward(ctx, model, fwd, bwd, x, *params):
        # x is a tensor, or a tuple of tensors for models with several inputs (none of them needs a gradient)
        outs, saved = fwd(model, x)
        # Entries of `saved` that ARE outputs go through save_for_backward: held in the plain dict they would close the
        # cycle ctx -> output -> grad_fn -> ctx, and a forward that is never backpropagated would keep its hundreds of
        # MB of gates / cell states until the garbage collector runs.
        outs_t = outs if isinstance(outs, tuple) else (outs,)
        keep, ctx.out_keys = [], {}
        for k, v in list(saved.items()):
            if any(v is t for t in outs_t):
                ctx.out_keys[k] = len(keep)
                keep.append(v)
                saved[k] = None
        ctx.save_for_backward(*keep)
        ctx.model, ctx.saved, ctx.bwd = model, saved, bwd
        ctx.names = [n for n, _ in model.named_parameters()]
        return outs

    @staticmethod
    def backward(ctx, *d_outs):
        if ctx.saved is None:
            raise RuntimeError("onssen_b200: the saved forward state of this graph was already consumed by a backward "
                               "(BPTT overwrites it in place); run the forward again")
        for k, i in ctx.out_keys.items():
            ctx.saved[k] = ctx.saved_tensors[i]
        sync = getattr(ctx.model, "grad_sync", None)
        grads = ctx.bwd(ctx.model, ctx.saved, [None if d is None else d.contiguous() for d in d_outs],
                        None if sync is None else sync.reduce_bucket)
        if sync is not None:
            sync.wait()
        ctx.saved = None
        return (None, None, None, None) + tuple(grads[n] for n in ctx.names)


def run_model(model, fwd, bwd, x):
    return _ModelFunction.apply(model, fwd, bwd, x, *[p for _, p in model.named_parameters()])


# ------------------------------------------------------------------------------------------------ deep clustering
def dc_forward_train(model, x):
    rnn, bn = model.rnn, model.bn
    B, T, F = x.shape
    H, D = rnn.hidden_size, model.embedding_dim
    M = T * B
    layers, packed, y_f = blstm_forward_train(rnn, model._rnn_cache, x, model.training, last_f32=True)
    a_h, mean, invstd = _lib.bn_forward_f16(y_f, M, H, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                            bn.running_var, bn.eps, bn.momentum, True, save_stats=True,
                                            sync=bn_sync(model))
    bn.num_batches_tracked += 1
    w_p = model._fc_cache.get([model.fc_dc.weight], lambda: _lib.pack_linear_f16(model.fc_dc.weight, True, H))
    emb = torch.empty(B, T, F, D, device=x.device, dtype=torch.float32)
    inv_norm = torch.empty(B, T, F, device=x.device, dtype=torch.float32)
    _lib.gemm_f16_ex(a_h, w_p, model.fc_dc.bias.detach(), emb, M, F * D, a_h.shape[1], F * D, epi=3, group=D,
                     remap_inner=B, remap_outer=T, inv_norm=inv_norm)
    saved = dict(layers=layers, y_f=y_f, a_h=a_h, mean=mean, invstd=invstd, emb=emb, inv_norm=inv_norm, packed=packed,
                 w_p=w_p, shape=(B, T, F))
    return emb, saved


def dc_backward(model, saved, d_outs, on_grads=None):
    """on_grads(dict) is called with each group of gradients as soon as it is final (head+BN, then every BLSTM
    layer from the top): the hook point for the overlapped data-parallel all-reduce (utils/ddp.py)."""
    d_emb, = d_outs
    rnn, bn = model.rnn, model.bn
    B, T, F = saved["shape"]
    H, D = rnn.hidden_size, model.embedding_dim
    M, N = T * B, F * D
    grads = {}
    dz32, sc = _lib.normalize_bwd(d_emb, saved["emb"], saved["inv_norm"])
    dA = linear_backward(dz32, sc, saved["a_h"], saved["w_p"], N, H, grads, "fc_dc")
    del dz32
    dY, grads["bn.weight"], grads["bn.bias"] = _lib.bn_backward(dA, saved["y_f"], M, H, bn.weight.detach(),
                                                                saved["mean"], saved["invstd"], sync=bn_sync(model))
    if on_grads is not None:
        on_grads(dict(grads))
    return blstm_backward(rnn, saved["layers"], saved["packed"], dY, B, T, grads, "rnn.", on_grads)


# ------------------------------------------------------------------------------------------------ chimera(++)
def chimera_forward_train(model, x):
    rnn = model.rnn
    B, T, F = x.shape
    H, D, S = rnn.hidden_size, model.embedding_dim, model.num_speaker
    M = T * B
    layers, packed, y_h = blstm_forward_train(rnn, model._rnn_cache, x, model.training, last_f32=False)
    wdc = model._dc_cache.get([model.fc_dc.weight], lambda: _lib.pack_linear_f16(model.fc_dc.weight, True, H))
    wmi = model._mi_cache.get([model.fc_mi.weight], lambda: _lib.pack_linear_f16(model.fc_mi.weight, True, H))
    emb = torch.empty(B, T, F, D, device=x.device, dtype=torch.float32)
    inv_norm = torch.empty(B, T, F, device=x.device, dtype=torch.float32)
    _lib.gemm_f16_ex(y_h, wdc, model.fc_dc.bias.detach(), emb, M, F * D, y_h.shape[1], F * D, epi=3, group=D,
                     remap_inner=B, remap_outer=T, inv_norm=inv_norm)
    masks = torch.empty(B, T, F, S, device=x.device, dtype=torch.float32)
    _lib.gemm_f16(y_h, wmi, model.fc_mi.bias.detach(), masks, M, F * S, y_h.shape[1], F * S, epi=1, remap_inner=B,
                  remap_outer=T)
    saved = dict(layers=layers, packed=packed, y_h=y_h, emb=emb, inv_norm=inv_norm, masks=masks, wdc=wdc, wmi=wmi,
                 shape=(B, T, F))
    return (emb, masks), saved


def chimera_backward(model, saved, d_outs, on_grads=None):
    d_emb, d_masks = d_outs
    rnn = model.rnn
    B, T, F = saved["shape"]
    H, D, S = rnn.hidden_size, model.embedding_dim, model.num_speaker
    M = T * B
    dev = saved["emb"].device
    grads = {}
    dY = None
    if d_emb is not None:
        dz32, sc = _lib.normalize_bwd(d_emb, saved["emb"], saved["inv_norm"])
        dY = linear_backward(dz32, sc, saved["y_h"], saved["wdc"], F * D, H, grads, "fc_dc")
        del dz32
    else:
        grads["fc_dc.weight"] = torch.zeros_like(model.fc_dc.weight)
        grads["fc_dc.bias"] = torch.zeros_like(model.fc_dc.bias)
    if d_masks is not None:
        dzm, scm = _lib.sigmoid_bwd(d_masks, saved["masks"])
        dYm = linear_backward(dzm, scm, saved["y_h"], saved["wmi"], F * S, H, grads, "fc_mi")
        dY = dYm if dY is None else _lib.add_inplace(dY, dYm)
    else:
        grads["fc_mi.weight"] = torch.zeros_like(model.fc_mi.weight)
        grads["fc_mi.bias"] = torch.zeros_like(model.fc_mi.bias)
    if dY is None:
        dY = torch.zeros(M, 2 * _lib.hp_of(H), device=dev, dtype=torch.float32)
    if on_grads is not None:
        on_grads(dict(grads))
    return blstm_backward(rnn, saved["layers"], saved["packed"], dY, B, T, grads, "rnn.", on_grads)


# ------------------------------------------------------------------------------------------------ enhance
def enhance_forward_train(model, x_and_noisy):
    x, mag_noisy = x_and_noisy
    rnn, bn = model.rnn, model.bn
    B, T, F = x.shape
    H = rnn.hidden_size
    M = T * B
    dev = x.device
    layers, packed, y_f = blstm_forward_train(rnn, model._rnn_cache, x, model.training, last_f32=True)
    a_h, mean, invstd = _lib.bn_forward_f16(y_f, M, H, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                            bn.running_var, bn.eps, bn.momentum, True, save_stats=True,
                                            sync=bn_sync(model))
    bn.num_batches_tracked += 1
    w_mi = model._mi.get([model.fc_mi.weight], lambda: _lib.pack_linear_f16(model.fc_mi.weight, True, H))
    w_pre = model._pre.get([model.fc_pre.weight], lambda: _lib.pack_linear_f16(model.fc_pre.weight, False))
    w_post = model._post.get([model.fc_post.weight], lambda: _lib.pack_linear_f16(model.fc_post.weight, False))
    mask = torch.empty(M, F, device=dev, dtype=torch.float32)
    _lib.gemm_f16(a_h, w_mi, model.fc_mi.bias.detach(), mask, M, F, a_h.shape[1], F, epi=1)
    noisy_h = _lib.pack_input_f16(mag_noisy.float().contiguous())
    pre = torch.empty(M, F, device=dev, dtype=torch.float32)
    _lib.gemm_f16(noisy_h, w_pre, model.fc_pre.bias.detach(), pre, M, F, noisy_h.shape[1], F, epi=2)
    est_h = _lib.mul_pack_f16(pre, mask)
    clean = torch.empty(B, T, F, device=dev, dtype=torch.float32)
    _lib.gemm_f16(est_h, w_post, model.fc_post.bias.detach(), clean, M, F, est_h.shape[1], F, epi=2, remap_inner=B,
                  remap_outer=T)
    saved = dict(layers=layers, packed=packed, y_f=y_f, a_h=a_h, mean=mean, invstd=invstd, mask=mask, pre=pre,
                 noisy_h=noisy_h, est_h=est_h, clean=clean, w_mi=w_mi, w_pre=w_pre, w_post=w_post, shape=(B, T, F))
    return clean, saved


def enhance_backward(model, saved, d_outs, on_grads=None):
    d_clean, = d_outs
    rnn, bn = model.rnn, model.bn
    B, T, F = saved["shape"]
    H = rnn.hidden_size
    M = T * B
    grads = {}
    dzpost, sc = _lib.relu_bwd(d_clean, saved["clean"])
    d_est = linear_backward(dzpost, sc, saved["est_h"], saved["w_post"], F, H, grads, "fc_post", K=F)
    dz_pre, sc_pre, dz_mi, sc_mi = _lib.enhance_mid_bwd(d_est, saved["pre"], saved["mask"])
    linear_backward(dz_pre, sc_pre, saved["noisy_h"], saved["w_pre"], F, H, grads, "fc_pre", K=F, want_dA=False)
    dA = linear_backward(dz_mi, sc_mi, saved["a_h"], saved["w_mi"], F, H, grads, "fc_mi")
    dY, grads["bn.weight"], grads["bn.bias"] = _lib.bn_backward(dA, saved["y_f"], M, H, bn.weight.detach(),
                                                                saved["mean"], saved["invstd"], sync=bn_sync(model))
    if on_grads is not None:
        on_grads(dict(grads))
    return blstm_backward(rnn, saved["layers"], saved["packed"], dY, B, T, grads, "rnn.", on_grads)


# ------------------------------------------------------------------------------------------------ phase_net (repaired)
def phase_forward_train(model, inp):
    x_mag, x_phase = inp
    ch = model.chimera
    (emb, masks), saved_ch = chimera_forward_train(ch, x_mag)
    B, T, F = x_mag.shape
    H = model.hidden_dim
    M = T * B
    bn = model.bn
    w_ph = model._ph.get([model.fc_phase.weight], lambda: _lib.pack_linear_f16(model.fc_phase.weight, True, H))
    branches, outs = [], []
    for s_idx in range(2):
        mk = masks[:, :, :, s_idx]
        xin = _lib.pack_phase_input_f16(x_mag, mk, mk.stride(-1), x_phase)
        layers, packed, y_f = blstm_forward_train_packed(model.rnn, model._rnn_cache, xin, B, T, model.training, True)
        a_h, mean, invstd = _lib.bn_forward_f16(y_f, M, H, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                                bn.running_var, bn.eps, bn.momentum, True, save_stats=True,
                                            sync=bn_sync(model))
        bn.num_batches_tracked += 1
        ph = torch.empty(B, T, F, 2, device=x_mag.device, dtype=torch.float32)
        _lib.gemm_f16(a_h, w_ph, model.fc_phase.bias.detach(), ph, M, 2 * F, a_h.shape[1], 2 * F, remap_inner=B,
                      remap_outer=T)
        outs.append(_lib.add_l2norm_pairs(ph, x_phase))
        branches.append(dict(layers=layers, packed=packed, y_f=y_f, a_h=a_h, mean=mean, invstd=invstd, ph=ph))
    saved = dict(ch=saved_ch, branches=branches, x_mag=x_mag, x_phase=x_phase, w_ph=w_ph, shape=(B, T, F))
    return (emb, masks, outs[0], outs[1]), saved


def phase_backward(model, saved, d_outs, on_grads=None):
    d_emb, d_masks, d_pa, d_pb = d_outs
    B, T, F = saved["shape"]
    H = model.hidden_dim
    M = T * B
    bn = model.bn
    dev = saved["x_mag"].device
    S = 2
    d_masks = torch.zeros(B, T, F, S, device=dev, dtype=torch.float32) if d_masks is None else d_masks.clone()
    total = None
    for s_idx, d_p in enumerate((d_pa, d_pb)):
        if d_p is None:
            continue
        br = saved["branches"][s_idx]
        g = {}
        dz, sc = _lib.l2norm_pairs_bwd(d_p, br["ph"], saved["x_phase"])
        dA = linear_backward(dz, sc, br["a_h"], saved["w_ph"], 2 * F, H, g, "fc_phase")
        dY, g["bn.weight"], g["bn.bias"] = _lib.bn_backward(dA, br["y_f"], M, H, bn.weight.detach(), br["mean"],
                                                            br["invstd"], sync=bn_sync(model))
        g, d_xin = blstm_backward(model.rnn, br["layers"], br["packed"], dY, B, T, g, "rnn.", None, want_dx0=True)
        _lib.phase_input_bwd(d_xin, saved["x_mag"], d_masks, s_idx)
        if total is None:
            total = g
        else:                                  # shared weights: the two passes add up
            done = set()
            for k, v in g.items():
                if v.data_ptr() not in done:    # bias_ih / bias_hh share one tensor
                    _lib.add_inplace(total[k], v)
                    done.add(v.data_ptr())
    if total is None:
        total = {n: torch.zeros_like(p) for n, p in model.named_parameters() if not n.startswith("chimera.")}
    g_ch = chimera_backward(model.chimera, saved["ch"], [d_emb, d_masks], None)
    grads = dict(total)
    grads.update({"chimera." + k: v for k, v in g_ch.items()})
    if on_grads is not None:
        seen, uniq = set(), {}
        for k, v in grads.items():
            if v.data_ptr() not in seen:
                uniq[k] = v
                seen.add(v.data_ptr())
        on_grads(uniq)
    return grads
