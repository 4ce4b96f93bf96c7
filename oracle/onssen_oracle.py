"""CPU oracle for the onssen STFT-mask separation hot path -- TEST INFRASTRUCTURE ONLY.

Plain-numpy restatement of the reference algorithm (speechLabBcCuny/onssen @ 179cff9).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module;
the product (`onssen_b200/`) never does and fails loudly without its CUDA library.

Pinning status (see tests/golden/README.md and oracle/make_golden.py):
  * models + losses (lstm stack, batchnorm, heads, loss_dc, chimera/mask losses): PINNED against the live
    reference torch modules imported from /root/reference (fixtures in tests/golden/*.npz).
  * STFT / iSTFT / wav featurizer: the reference delegates to librosa (not installed, version unpinned in the
    reference).  These functions restate librosa<=0.9 semantics (periodic Hann, win_length=n_fft,
    center=True, reflect padding, window-sum-square iSTFT) and are pinned only against torch.stft/istft and
    scipy windows  ->  "parity unpinned" at the librosa boundary.
  * phase_net / loss_phase: the reference raises as written (phase_network.py:28, loss_phase.py:7-13);
    these are REPAIRED restatements with no reference oracle (repairs listed at the functions).

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------------
# featurizer  (onssen/data/feature_utils.py, onssen/data/wsj0_2mix.py)
# ----------------------------------------------------------------------------------------------------
def hann_periodic(n_fft):
    """scipy.signal.get_window('hann', n_fft, fftbins=True) (what librosa.stft uses), float64."""
    n = np.arange(n_fft, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)


def stft(sig, n_fft, hop):
    """feature_utils.py:20  np.transpose(librosa.core.stft(sig, n_fft=window_size, hop_length=hop_size)).
    Returns (frames, n_fft//2+1) complex64, frames = 1 + len(sig)//hop."""
    sig = np.asarray(sig, dtype=F32)
    pad = n_fft // 2
    y = np.pad(sig, pad, mode="reflect")
    frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(frames)[:, None]
    win = hann_periodic(n_fft)
    spec = np.fft.rfft(y[idx].astype(np.float64) * win[None, :], axis=1)
    return spec.astype(np.complex64)


def tile_and_crop(spec, frame_length, start):
    """wsj0_2mix.py:118-128. `start` replaces np.random.randint(frames - frame_length)."""
    if spec.shape[0] <= frame_length:
        times = frame_length // spec.shape[0] + 1
        spec = np.concatenate([spec] * times, axis=0)
    assert 0 <= start < spec.shape[0] - frame_length or (start == 0 and spec.shape[0] > frame_length)
    return spec[start:start + frame_length]


def num_crop_starts(nsample, hop, frame_length):
    """exclusive upper bound of the crop start (argument of np.random.randint at wsj0_2mix.py:125)."""
    frames = 1 + nsample // hop
    if frames <= frame_length:
        frames *= frame_length // frames + 1
    return frames - frame_length


def log_magnitude(spec, epsilon=1e-7):
    """feature_utils.py:49-51"""
    return np.log10(np.abs(spec) + F32(epsilon)).astype(F32)


def phase(spec):
    """feature_utils.py:54-64 -> (frames, F, 2) raw (re, im)"""
    return np.stack([np.real(spec), np.imag(spec)], axis=-1).astype(F32)


def cos_difference(a, b):
    """feature_utils.py:77-80"""
    return np.cos(np.angle(a) - np.angle(b)).astype(F32)


def one_hot(feature_mix, mag_s1, mag_s2, db_threshold):
    """feature_utils.py:83-95 (float64 result as in the reference; argmax ties -> speaker 0; strict <)."""
    specs = np.asarray([mag_s1, mag_s2])
    vals = np.argmax(specs, axis=0)
    Y = np.zeros(mag_s1.shape + (2,))
    Y[vals == 0, 0] = 1
    Y[vals == 1, 1] = 1
    m = np.max(feature_mix) - db_threshold / 20
    Y[feature_mix < m] = 0
    return Y


def featurize(mix, s1, s2, n_fft, hop, frame_length, start, db_threshold, model_name):
    """wsj0_2mix.py:114-152 for one utterance. Returns (input_list, label_list) of numpy arrays."""
    sm = tile_and_crop(stft(mix, n_fft, hop), frame_length, start)
    s1c = tile_and_crop(stft(s1, n_fft, hop), frame_length, start)
    s2c = tile_and_crop(stft(s2, n_fft, hop), frame_length, start)
    feature = log_magnitude(sm)
    mag_mix, mag_s1, mag_s2 = np.abs(sm), np.abs(s1c), np.abs(s2c)
    oh = one_hot(feature, mag_s1, mag_s2, db_threshold)
    if model_name == "dc":
        return [feature], [oh, mag_mix]
    if model_name == "chimera":
        return [feature], [oh, mag_mix, mag_s1, mag_s2]
    if model_name == "chimera++":
        return [feature], [oh, mag_mix, mag_s1, mag_s2, cos_difference(sm, s1c), cos_difference(sm, s2c)]
    if model_name == "phase":
        return [feature, phase(sm)], [oh, mag_mix, mag_s1, mag_s2, phase(s1c), phase(s2c)]
    raise ValueError(model_name)


def istft(spec, hop, length):
    """egs/wsj0-2mix/deep_clustering/evaluate.py:45  librosa.core.istft(stft.T, hop_length=hop, length=n).
    spec: (frames, F) complex. librosa<=0.9: Hann(periodic) synthesis window, overlap-add, division by the
    window sum-square where > tiny, trim n_fft//2, fix_length."""
    spec = np.asarray(spec)
    frames, F = spec.shape
    n_fft = 2 * (F - 1)
    win = hann_periodic(n_fft)
    total = n_fft + hop * (frames - 1)
    y = np.zeros(total, dtype=np.float64)
    wss = np.zeros(total, dtype=np.float64)
    xt = np.fft.irfft(spec.astype(np.complex128), n=n_fft, axis=1) * win[None, :]
    for i in range(frames):
        y[i * hop:i * hop + n_fft] += xt[i]
        wss[i * hop:i * hop + n_fft] += win * win
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2:]
    if len(y) >= length:
        return y[:length]
    return np.pad(y, (0, length - len(y)))


def masked_istft(stft_re, stft_im, masks, hop, nsample):
    """evaluate.py:31-46: stft_mix * mask[s] -> istft, for each speaker. masks (S, frames, F)."""
    sm = stft_re.astype(np.float64) + 1j * stft_im.astype(np.float64)
    return np.stack([istft(sm * masks[s], hop, nsample) for s in range(masks.shape[0])])


# ----------------------------------------------------------------------------------------------------
# network pieces
# ----------------------------------------------------------------------------------------------------
def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of torch.nn.LSTM (gate order i,f,g,o; h0=c0=0). x (B,T,I) -> (B,T,H)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    G = x.reshape(B * T, -1) @ w_ih.T + (b_ih + b_hh)
    G = G.reshape(B, T, 4 * H)
    h = np.zeros((B, H), dtype=x.dtype)
    c = np.zeros((B, H), dtype=x.dtype)
    out = np.empty((B, T, H), dtype=x.dtype)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    whh_t = np.ascontiguousarray(w_hh.T)
    for t in steps:
        g = G[:, t] + h @ whh_t
        i = _sigmoid(g[:, 0:H])
        f = _sigmoid(g[:, H:2 * H])
        gg = np.tanh(g[:, 2 * H:3 * H])
        o = _sigmoid(g[:, 3 * H:4 * H])
        c = f * c + i * gg
        h = o * np.tanh(c)
        out[:, t] = h
    return out


def blstm_stack(x, params, prefix, num_layers):
    """nn.LSTM(bidirectional=True, batch_first=True, dropout inactive) -- deep_clustering.py:15-22,34-35.
    params: dict of state_dict-style arrays, keys '<prefix>weight_ih_l{k}[_reverse]' ..."""
    y = x
    for l in range(num_layers):
        outs = []
        for suf, rev in (("", False), ("_reverse", True)):
            outs.append(lstm_direction(y, params[f"{prefix}weight_ih_l{l}{suf}"], params[f"{prefix}weight_hh_l{l}{suf}"],
                                       params[f"{prefix}bias_ih_l{l}{suf}"], params[f"{prefix}bias_hh_l{l}{suf}"], rev))
        y = np.concatenate(outs, axis=2)
    return y


def batchnorm_bt(y, params, prefix, training, eps=1e-5, momentum=0.1):
    """permute + nn.BatchNorm1d(2H) + permute (deep_clustering.py:36-38): per-channel stats over (B,T).
    Returns (out, new_running_mean, new_running_var)."""
    B, T, C = y.shape
    flat = y.reshape(B * T, C).astype(np.float64)
    rm, rv = params[prefix + "running_mean"], params[prefix + "running_var"]
    if training:
        mean = flat.mean(0)
        var = flat.var(0)
        n = flat.shape[0]
        new_rm = (1 - momentum) * rm + momentum * mean
        new_rv = (1 - momentum) * rv + momentum * var * n / max(n - 1, 1)
    else:
        mean, var = rm.astype(np.float64), rv.astype(np.float64)
        new_rm, new_rv = rm, rv
    out = (flat - mean) / np.sqrt(var + eps) * params[prefix + "weight"] + params[prefix + "bias"]
    return out.reshape(B, T, C).astype(y.dtype), new_rm.astype(F32), new_rv.astype(F32)


def linear(x, params, prefix):
    return x @ params[prefix + "weight"].T + params[prefix + "bias"]


def l2_normalize(x, eps=1e-12):
    """F.normalize(p=2, dim=-1): x / max(||x||, eps)"""
    n = np.sqrt(np.sum(x * x, axis=-1, keepdims=True))
    return x / np.maximum(n, eps)


def deep_clustering_forward(params, inp, num_layers, training=False):
    """onssen/nn/deep_clustering.py:29-43 (dropout inactive)."""
    assert len(inp) == 1
    x = inp[0].astype(F32)
    B, T, F = x.shape
    y = blstm_stack(x, params, "rnn.", num_layers)
    y, _, _ = batchnorm_bt(y, params, "bn.", training)
    e = linear(y, params, "fc_dc.")
    e = l2_normalize(e.reshape(B, T * F, -1)).reshape(B, T, F, -1)
    return [e.astype(F32)]


def chimera_forward(params, inp, num_layers, prefix=""):
    """onssen/nn/chimera.py:30-46."""
    assert len(inp) == 1
    x = inp[0].astype(F32)
    B, T, F = x.shape
    y = blstm_stack(x, params, prefix + "rnn.", num_layers)
    e = linear(y, params, prefix + "fc_dc.")
    e = l2_normalize(e.reshape(B, T * F, -1)).reshape(B, T, F, -1)
    m = _sigmoid(linear(y, params, prefix + "fc_mi.")).reshape(B, T, F, -1)
    return [e.astype(F32), m[..., 0].astype(F32), m[..., 1].astype(F32)]


def enhance_forward(params, inp, num_layers, training=False):
    """onssen/nn/enhancement.py:38-52."""
    assert len(inp) == 2
    x, mag_noisy = inp
    y = blstm_stack(x.astype(F32), params, "rnn.", num_layers)
    y, _, _ = batchnorm_bt(y, params, "bn.", training)
    mask = _sigmoid(linear(y, params, "fc_mi."))
    pre = np.maximum(linear(mag_noisy.astype(F32), params, "fc_pre."), 0)
    clean = np.maximum(linear(pre * mask, params, "fc_post."), 0)
    return [clean.astype(F32)]


def phase_net_forward(params, inp, num_layers, training=False):
    """REPAIRED restatement of onssen/nn/phase_network.py:34-67 (no reference oracle).
    Repair: fc_phase = Linear(2H, 2*F) (the undefined `output_dim` at :28 := input_dim, num_speaker=2) so that
    the reshape at :55,63 yields (B,T,F,2) and can be added to x_phase."""
    x_mag, x_phase = inp
    emb, mask_a, mask_b = chimera_forward(params, [x_mag], num_layers, prefix="chimera.")
    B, T, F = mask_a.shape
    outs = []
    for mk in (mask_a, mask_b):
        xin = np.concatenate([x_mag * mk, x_phase.reshape(B, T, -1)], axis=2).astype(F32)
        y = blstm_stack(xin, params, "rnn.", num_layers)
        y, _, _ = batchnorm_bt(y, params, "bn.", training)
        ph = linear(y, params, "fc_phase.").reshape(B, T, F, -1) + x_phase
        outs.append(l2_normalize(ph).astype(F32))
    return [emb, mask_a, mask_b, outs[0], outs[1]]


# ----------------------------------------------------------------------------------------------------
# losses  (onssen/loss/*.py)
# ----------------------------------------------------------------------------------------------------
def _norm(x):
    """loss_util.py:7-11: sqrt of the sum of squares per batch item (UN-squared Frobenius norm)."""
    b = x.shape[0]
    return np.sqrt(np.sum((x * x).reshape(b, -1), axis=1))


def _norm_1d(x):
    """loss_util.py:13-16"""
    return np.sum(np.abs(x.reshape(x.shape[0], -1)), axis=1)


def loss_dc(output, label):
    """onssen/loss/loss_dc.py:6-44. Returns the (B,B) array loss_embedding (B,) * mag_mix.sum (B,1)."""
    assert len(output) == 1 and len(label) == 2
    emb, = output
    lab, mag = label
    lab = lab.astype(F32)
    B, T, Fq, S = lab.shape
    D = emb.shape[-1]
    emb = emb.reshape(B, -1, D).astype(F32)
    mag = mag.reshape(B, -1).astype(F32)
    lab = lab.reshape(B, -1, S)
    silence = lab.sum(2, keepdims=True)
    emb = silence * emb
    msum = mag.sum(1, keepdims=True)
    w = np.sqrt(mag / msum)[..., None]
    lab = lab * w
    emb = emb * w
    et = np.transpose(emb, (0, 2, 1))
    lt = np.transpose(lab, (0, 2, 1))
    loss = _norm(et @ emb) - 2 * _norm(et @ lab) + _norm(lt @ lab)
    return (loss * msum).astype(F32)          # (B,) * (B,1) -> (B,B)


def loss_chimera_msa(output, label):
    """onssen/loss/loss_chimera.py:6-31"""
    emb, ma, mb = output
    oh, mix, s1, s2 = label
    le = loss_dc([emb], [oh, mix])
    l1 = _norm_1d(ma * mix - s1) + _norm_1d(mb * mix - s2)
    l2 = _norm_1d(mb * mix - s1) + _norm_1d(ma * mix - s2)
    return (le * F32(0.975) + np.minimum(l1, l2) * F32(0.025)).astype(F32)


def loss_chimera_psa(output, label):
    """onssen/loss/loss_chimera.py:33-59"""
    emb, ma, mb = output
    oh, mix, s1, s2, c1, c2 = label
    le = loss_dc([emb], [oh, mix])
    t1 = np.minimum(mix, np.maximum(s1 * c1, 0))
    t2 = np.minimum(mix, np.maximum(s2 * c2, 0))
    l1 = _norm_1d(ma * mix - t1) + _norm_1d(mb * mix - t2)
    l2 = _norm_1d(mb * mix - t1) + _norm_1d(ma * mix - t2)
    return (le * F32(0.975) + np.minimum(l1, l2) * F32(0.025)).astype(F32)


def loss_mask_msa(output, label):
    """onssen/loss/loss_mask.py:6-22: nn.MSELoss()(clean_est, mag_clean) -> scalar"""
    est, = output
    clean, _ = label
    return np.mean((est - clean) ** 2, dtype=np.float64).astype(F32)


def loss_mask_psa(output, label):
    """onssen/loss/loss_mask.py:25-40"""
    mask, = output
    noisy, clean, cosd = label
    return _norm_1d(mask * noisy - np.minimum(noisy, np.maximum(clean * cosd, 0))).astype(F32)


def loss_mask_psa_grad(mask, noisy, clean, cosd, g):
    """d/dmask of sum_b g[b] * loss_mask_psa(...)[b] (onssen/loss/loss_mask.py:25-40): sign(residual) * noisy."""
    res = mask * noisy - np.minimum(noisy, np.maximum(clean * cosd, 0))
    return (np.sign(res) * noisy * g.reshape(-1, 1, 1)).astype(F32)


def loss_phase(output, label):
    """REPAIRED restatement of onssen/loss/loss_phase.py:6-37 (no reference oracle).
    Repairs: assert 5 outputs (the reference asserts 6 then unpacks 5, :7,9); the embedding term is
    loss_dc([embedding], [one_hot_label, mag_mix]) (the reference passes mag_mix on the wrong side, :13)."""
    emb, ma, mb, pa, pb = output
    oh, mix, s1, s2, p1, p2 = label
    B = mix.shape[0]
    le = loss_dc([emb], [oh, mix])
    l1 = _norm_1d(ma * mix - s1) + _norm_1d(mb * mix - s2)
    l2 = _norm_1d(mb * mix - s1) + _norm_1d(ma * mix - s2)
    amin = l1 < l2
    lm = np.where(amin, l1, l2)

    def cs(a, b, eps=1e-8):
        num = np.sum(a * b, axis=3)
        den = np.maximum(np.sqrt(np.sum(a * a, axis=3)), eps) * np.maximum(np.sqrt(np.sum(b * b, axis=3)), eps)
        return num / den

    lp1 = np.sum((-mix * cs(pa, p1) - mix * cs(pb, p2)).reshape(B, -1), axis=1)
    lp2 = np.sum((-mix * cs(pb, p1) - mix * cs(pa, p2)).reshape(B, -1), axis=1)
    lp = np.where(amin, lp1, lp2)
    return (le * F32(0.975) + lm * F32(0.025) + lp * F32(0.025)).astype(F32)


# ----------------------------------------------------------------------------------------------------
# optimiser step  (onssen/utils/train.py:83-84, onssen/utils/basic.py:6-7)
# ----------------------------------------------------------------------------------------------------
def clip_adam_step(params, grads, exp_avg, exp_avg_sq, step, max_norm=5.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.nn.utils.clip_grad_norm_(params, max_norm) then torch.optim.Adam(lr).step() [ext: torch semantics --
    coefficient min(1, max_norm/(norm+1e-6)); bias-corrected moments, denom = sqrt(v)/sqrt(1-b2^t) + eps].
    All arguments are lists of float32 arrays, updated in place; `step` is the 1-based step count.
    Returns the total gradient norm."""
    norm = np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads))
    coef = min(1.0, max_norm / (norm + 1e-6))
    b1, b2 = betas
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        if coef < 1.0:
            g *= F32(coef)
        m[...] = F32(b1) * m + F32(1 - b1) * g
        v[...] = F32(b2) * v + F32(1 - b2) * g * g
        denom = np.sqrt(v) / F32(np.sqrt(bc2)) + F32(eps)
        p -= F32(lr / bc1) * (m / denom)
    return norm


# ----------------------------------------------------------------------------------------------------
# synthetic workload shared by tests and bench
# ----------------------------------------------------------------------------------------------------
def synth_utterance(index, nsample=32000, seed0=1234):
    """Gated Gaussian noise sources so that the -40 dB VAD rejects a realistic share of bins."""
    rng = np.random.RandomState(seed0 + index)
    def src():
        x = rng.standard_normal(nsample).astype(F32)
        gate = (rng.uniform(size=nsample // 800 + 1) > 0.35).astype(F32)
        env = np.repeat(gate, 800)[:nsample]
        k = np.hanning(401).astype(F32)
        env = np.convolve(env, k / k.sum(), mode="same").astype(F32)
        col = np.convolve(x, np.array([1.0, 0.9, 0.5, 0.2], dtype=F32), mode="same")  # mild spectral tilt
        return (col * (env * 0.999 + 0.001) * F32(0.1)).astype(F32)
    s1, s2 = src(), src()
    return (s1 + s2).astype(F32), s1, s2


def init_params_like_torch(rng, input_dim, hidden, num_layers, heads):
    """PyTorch default inits (uniform +-1/sqrt(H) for LSTM, kaiming-uniform-ish for Linear) as numpy."""
    p = {}
    k = 1.0 / np.sqrt(hidden)
    for l in range(num_layers):
        I = input_dim if l == 0 else 2 * hidden
        for suf in ("", "_reverse"):
            p[f"rnn.weight_ih_l{l}{suf}"] = rng.uniform(-k, k, (4 * hidden, I)).astype(F32)
            p[f"rnn.weight_hh_l{l}{suf}"] = rng.uniform(-k, k, (4 * hidden, hidden)).astype(F32)
            p[f"rnn.bias_ih_l{l}{suf}"] = rng.uniform(-k, k, (4 * hidden,)).astype(F32)
            p[f"rnn.bias_hh_l{l}{suf}"] = rng.uniform(-k, k, (4 * hidden,)).astype(F32)
    for name, (n_out, n_in) in heads.items():
        kk = 1.0 / np.sqrt(n_in)
        p[name + ".weight"] = rng.uniform(-kk, kk, (n_out, n_in)).astype(F32)
        p[name + ".bias"] = rng.uniform(-kk, kk, (n_out,)).astype(F32)
    return p
