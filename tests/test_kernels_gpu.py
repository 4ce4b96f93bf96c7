"""Stage-level parity of the sm_100a kernels (through the C ABI) against numpy restatements / the oracle."""
import numpy as np
import pytest
import torch

from oracle import onssen_oracle as O

pytestmark = pytest.mark.gpu


def f16(a):
    return np.asarray(a).astype(np.float16).astype(np.float64)


@pytest.fixture(scope="module")
def lib(cuda_device):
    from onssen_b200 import _lib
    _lib.load()
    return _lib


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,epi", [(300, 520, 192, 0), (128, 256, 64, 0), (1000, 4864, 1216, 0), (257, 258, 128, 1),
                                       (64, 513, 576, 2), (130, 48, 64, 0)])
def test_gemm_plain_epilogues(lib, M, N, K, epi):
    rng = np.random.RandomState(M + N)
    A = rng.standard_normal((M, K)).astype(np.float16)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float16)
    bias = rng.standard_normal(N).astype(np.float32)
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.gemm_f16(torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(bias).cuda(), out, M, N, K,
                 N, epi=epi)
    ref = A.astype(np.float64) @ W.astype(np.float64).T + bias
    if epi == 1:
        ref = 1 / (1 + np.exp(-ref))
    if epi == 2:
        ref = np.maximum(ref, 0)
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, atol=2e-5 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("M,N,Kc,shift,col0", [(4864, 1216, 1300, 0, 0), (2432, 608, 777, -32, 608), (2432, 608, 640, 32, 0),
                                               (520, 192, 130, 0, 64), (128, 64, 64, -3, 0), (304, 96, 1000, 5, 8)])
def test_gemm_rows_mode_vs_fp64(lib, M, N, Kc, shift, col0):
    """weight-gradient form: out = scale * X^T Y with the contraction over the ROWS of two row-major fp16 buffers read in
    place (MN-major UMMA operands), Y shifted by whole rows (zero outside) and taken as a column slice of a wider buffer"""
    rng = np.random.RandomState(M + N + Kc)
    X = rng.standard_normal((Kc, M)).astype(np.float16)
    Yw = rng.standard_normal((Kc, col0 + N + 8)).astype(np.float16)
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Yw).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    scale = torch.tensor([0.25], device="cuda")
    lib.gemm_f16_rows(Xd, Yd[:, col0:col0 + N], out, M, N, Kc, y_row_shift=shift, out_scale=scale)
    Y = Yw[:, col0:col0 + N].astype(np.float64)
    Ys = np.zeros_like(Y)
    if shift >= 0:
        Ys[:Kc - shift] = Y[shift:]
    else:
        Ys[-shift:] = Y[:Kc + shift]
    ref = 0.25 * X.astype(np.float64).T @ Ys
    got = out.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 2e-5 * np.sqrt(Kc) * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("B,T,F,D,K", [(3, 50, 129, 40, 1216), (2, 70, 33, 20, 128), (5, 26, 9, 8, 64)])
def test_gemm_l2norm_remap_epilogue(lib, B, T, F, D, K):
    rng = np.random.RandomState(B * T)
    M, N = T * B, F * D
    A = rng.standard_normal((M, K)).astype(np.float16)       # time-major rows m = t*B + b
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float16)
    bias = (0.1 * rng.standard_normal(N)).astype(np.float32)
    out = torch.full((B, T, F, D), float("nan"), device="cuda")
    lib.gemm_f16(torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda(), torch.from_numpy(bias).cuda(), out, M, N, K,
                 N, epi=3, group=D, remap_inner=B, remap_outer=T)
    lin = (A.astype(np.float64) @ W.astype(np.float64).T + bias).reshape(T, B, F, D).transpose(1, 0, 2, 3)
    ref = lin / np.maximum(np.sqrt((lin * lin).sum(-1, keepdims=True)), 1e-12)
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, atol=3e-6)


# ------------------------------------------------------------------------------------------------ packing
def test_pack_layouts(lib):
    rng = np.random.RandomState(0)
    H, I = 40, 9
    Hp = 64
    mk = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).cuda()
    wf = (mk(4 * H, I), mk(4 * H, H), mk(4 * H), mk(4 * H))
    wr = (mk(4 * H, I), mk(4 * H, H), mk(4 * H), mk(4 * H))
    wih_p, whh_p, bias_p = lib.lstm_pack_layer(wf, wr, H, I, False, 0)
    assert wih_p.shape == (8 * Hp, 64)
    wih = wih_p.cpu().numpy().astype(np.float32)
    whh = whh_p.cpu().numpy().astype(np.float32).reshape(2, Hp // 32, 128, Hp)
    bias = bias_p.cpu().numpy()
    for d, w in enumerate((wf, wr)):
        w_ih, w_hh, b_ih, b_hh = [t.cpu().numpy() for t in w]
        for u in (0, 7, 31, 32, 39, 40, 63):
            for gate in range(4):
                rb, ul = divmod(u, 32)
                n = d * 4 * Hp + rb * 128 + 4 * ul + gate
                if u < H:
                    np.testing.assert_array_equal(wih[n, :I], w_ih[gate * H + u].astype(np.float16).astype(np.float32))
                    assert (wih[n, I:] == 0).all()
                    row = whh[d, rb, 4 * ul + gate, :]
                    np.testing.assert_array_equal(row[:H], w_hh[gate * H + u].astype(np.float16).astype(np.float32))
                    assert (row[H:] == 0).all()
                    assert bias[n] == b_ih[gate * H + u] + b_hh[gate * H + u]
                else:
                    assert (wih[n] == 0).all() and bias[n] == 0
    x = mk(3, 5, I)
    xh = lib.pack_input_f16(x).cpu().numpy()
    xr = x.cpu().numpy()
    for t in range(5):
        for b in range(3):
            np.testing.assert_array_equal(xh[t * 3 + b, :I], xr[b, t].astype(np.float16))
    assert (xh[:, I:] == 0).all()
    # blstm-input linear packing
    w = mk(6, 2 * H)
    wp = lib.pack_linear_f16(w, True, H).cpu().numpy()
    wn = w.cpu().numpy().astype(np.float16)
    np.testing.assert_array_equal(wp[:, :H], wn[:, :H])
    np.testing.assert_array_equal(wp[:, Hp:Hp + H], wn[:, H:])
    assert (wp[:, H:Hp] == 0).all() and (wp[:, Hp + H:] == 0).all()


# ------------------------------------------------------------------------------------------------ BLSTM layer
def blstm_layer_emulated(x, wf, wr, H):
    """fp16-rounded operands (x, W_ih, W_hh, h), exact accumulation: what the tensor-core path computes."""
    B, T, I = x.shape
    out = np.zeros((B, T, 2 * H))
    for d, (w_ih, w_hh, b_ih, b_hh) in enumerate((wf, wr)):
        G = (f16(x).reshape(B * T, I) @ f16(w_ih).T).astype(np.float32).astype(np.float64) + (b_ih + b_hh)
        G = G.reshape(B, T, 4 * H)
        h = np.zeros((B, H)); c = np.zeros((B, H))
        W = f16(w_hh).T
        for s in range(T):
            t = T - 1 - s if d else s
            g = G[:, t] + f16(h) @ W
            i, f, gg, o = (1 / (1 + np.exp(-g[:, :H])), 1 / (1 + np.exp(-g[:, H:2 * H])), np.tanh(g[:, 2 * H:3 * H]),
                           1 / (1 + np.exp(-g[:, 3 * H:])))
            c = f * c + i * gg
            h = o * np.tanh(c)
            out[:, t, d * H:(d + 1) * H] = h
    return out


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("B,T,H,I", [(5, 7, 600, 129), (33, 12, 40, 9), (32, 6, 600, 129), (2, 9, 300, 33),
                                     (64, 5, 600, 129)])
def test_blstm_layer_vs_emulation(lib, tc, B, T, H, I):
    rng = np.random.RandomState(B * T + H)
    k = 1 / np.sqrt(H)
    mkw = lambda *s: rng.uniform(-k, k, s).astype(np.float32)
    wf = (mkw(4 * H, I), mkw(4 * H, H), mkw(4 * H), mkw(4 * H))
    wr = (mkw(4 * H, I), mkw(4 * H, H), mkw(4 * H), mkw(4 * H))
    x = rng.standard_normal((B, T, I)).astype(np.float32)
    cu = lambda a: torch.from_numpy(a).cuda()
    wih_p, whh_p, bias_p = lib.lstm_pack_layer(tuple(map(cu, wf)), tuple(map(cu, wr)), H, I, False, 0)
    Hp = lib.hp_of(H)
    xh = lib.pack_input_f16(cu(x))
    M = T * B
    gates = torch.empty(M, 8 * Hp, device="cuda")
    lib.gemm_f16(xh, wih_p, bias_p, gates, M, 8 * Hp, xh.shape[1], 8 * Hp)
    y_h = torch.full((M, 2 * Hp), float("nan"), device="cuda", dtype=torch.float16)
    y_f = torch.full((M, 2 * Hp), float("nan"), device="cuda")
    lib.blstm_rec_fwd(gates, whh_p, B, T, H, y_h, y_f, use_tensor_cores=tc)
    torch.cuda.synchronize()
    ref = blstm_layer_emulated(x, wf, wr, H)
    got = y_f.cpu().numpy().reshape(T, B, 2, Hp)
    assert np.isfinite(got).all()
    assert (got[..., H:] == 0).all()                      # padded units stay exactly zero
    got = got[..., :H].transpose(1, 0, 2, 3).reshape(B, T, 2 * H)
    np.testing.assert_allclose(got, ref, atol=2e-5)
    goth = y_h.float().cpu().numpy().reshape(T, B, 2, Hp)[..., :H].transpose(1, 0, 2, 3).reshape(B, T, 2 * H)
    np.testing.assert_allclose(goth, ref, atol=1e-3)


# ------------------------------------------------------------------------------------------------ BatchNorm
@pytest.mark.parametrize("training", [True, False])
def test_bn_forward(lib, training):
    rng = np.random.RandomState(3)
    H, M = 40, 777
    Hp = lib.hp_of(H)
    y = np.zeros((M, 2 * Hp), dtype=np.float32)
    vals = (rng.standard_normal((M, 2 * H)) * rng.uniform(0.1, 2, 2 * H) + rng.uniform(-1, 1, 2 * H)).astype(np.float32)
    y[:, :H] = vals[:, :H]; y[:, Hp:Hp + H] = vals[:, H:]
    params = {"bn.weight": rng.uniform(0.5, 1.5, 2 * H).astype(np.float32), "bn.bias": rng.standard_normal(2 * H).astype(np.float32),
              "bn.running_mean": rng.standard_normal(2 * H).astype(np.float32), "bn.running_var": rng.uniform(0.5, 1.5, 2 * H).astype(np.float32)}
    cu = lambda a: torch.from_numpy(a.copy()).cuda()
    rm, rv = cu(params["bn.running_mean"]), cu(params["bn.running_var"])
    out_h, sm, si = lib.bn_forward_f16(cu(y), M, H, cu(params["bn.weight"]), cu(params["bn.bias"]), rm, rv, 1e-5, 0.1,
                                       training, save_stats=True)
    ref, rm_ref, rv_ref = O.batchnorm_bt(vals.reshape(1, M, 2 * H), params, "bn.", training)
    got = out_h.float().cpu().numpy()
    got = np.concatenate([got[:, :H], got[:, Hp:Hp + H]], 1)
    np.testing.assert_allclose(got, ref[0], atol=2e-3 * np.abs(ref).max())   # fp16 output
    np.testing.assert_allclose(rm.cpu().numpy(), rm_ref, atol=1e-6)
    np.testing.assert_allclose(rv.cpu().numpy(), rv_ref, rtol=1e-5)
    assert (out_h.float().cpu().numpy()[:, H:Hp] == 0).all()


# ------------------------------------------------------------------------------------------------ losses
@pytest.mark.parametrize("D,dtype", [(40, torch.float64), (20, torch.float32), (16, torch.uint8), (12, torch.float32)])
def test_loss_dc_vs_oracle(lib, D, dtype):
    rng = np.random.RandomState(D)
    B, T, F = 3, 37, 129
    emb = O.l2_normalize(rng.standard_normal((B, T, F, D)).astype(np.float32))
    mag = np.abs(rng.standard_normal((B, T, F))).astype(np.float32)
    oh = np.zeros((B, T, F, 2))
    first = rng.uniform(size=(B, T, F)) > 0.4
    oh[..., 0] = first; oh[..., 1] = ~first
    oh[rng.uniform(size=(B, T, F)) < 0.3] = 0
    ref = O.loss_dc([emb], [oh, mag])
    lab = torch.from_numpy(oh).to(dtype).cuda()
    got, l, msum = lib.loss_dc_fwd(torch.from_numpy(emb).cuda().view(B, T * F, D), lab.view(B, T * F, 2),
                                   torch.from_numpy(mag).cuda().view(B, T * F))
    assert got.shape == (B, B)
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-5)
    np.testing.assert_allclose(msum.cpu().numpy(), mag.reshape(B, -1).sum(1), rtol=1e-5)


def test_loss_pit_l1_vs_oracle(lib):
    rng = np.random.RandomState(5)
    B, T, F = 4, 31, 129
    masks = rng.uniform(size=(B, T, F, 2)).astype(np.float32)
    mix, s1, s2 = [np.abs(rng.standard_normal((B, T, F))).astype(np.float32) for _ in range(3)]
    c1, c2 = [np.cos(rng.uniform(-3, 3, (B, T, F))).astype(np.float32) for _ in range(2)]
    cu = lambda a: torch.from_numpy(a).cuda()
    mt = cu(masks)
    got, perm = lib.loss_pit_l1_fwd(mt[..., 0], mt[..., 1], 2, cu(mix), cu(s1), cu(s2))
    n1 = lambda x: np.abs(x.reshape(B, -1)).sum(1)
    ma, mb = masks[..., 0], masks[..., 1]
    l1 = n1(ma * mix - s1) + n1(mb * mix - s2); l2 = n1(mb * mix - s1) + n1(ma * mix - s2)
    np.testing.assert_allclose(got.cpu().numpy(), np.minimum(l1, l2), rtol=1e-5)
    np.testing.assert_array_equal(perm.cpu().numpy(), (~(l1 < l2)).astype(np.int32))
    got, _ = lib.loss_pit_l1_fwd(mt[..., 0], mt[..., 1], 2, cu(mix), cu(s1), cu(s2), cu(c1), cu(c2))
    t1 = np.minimum(mix, np.maximum(s1 * c1, 0)); t2 = np.minimum(mix, np.maximum(s2 * c2, 0))
    l1 = n1(ma * mix - t1) + n1(mb * mix - t2); l2 = n1(mb * mix - t1) + n1(ma * mix - t2)
    np.testing.assert_allclose(got.cpu().numpy(), np.minimum(l1, l2), rtol=1e-5)


# ------------------------------------------------------------------------------------------------ featurizer
@pytest.mark.parametrize("n_fft,hop,ns,T", [(256, 64, 32000, 400), (512, 128, 64000, 400), (1024, 256, 64000, 400),
                                            (64, 16, 2000, 50)])
def test_stft_features_vs_oracle(lib, n_fft, hop, ns, T):
    B = 3
    utts = [O.synth_utterance(i, ns) for i in range(B)]
    cu = lambda k: torch.from_numpy(np.stack([u[k] for u in utts])).cuda()
    hi = O.num_crop_starts(ns, hop, T)
    starts = np.array([0, hi - 1, hi // 2], dtype=np.int32)
    want = ["feature", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2", "ph_mix", "ph_s1", "ph_s2", "feat_max"]
    o = lib.stft_features(cu(0), cu(1), cu(2), n_fft, hop, torch.from_numpy(starts), T, want)
    oh = lib.one_hot_vad(o["feature"], o["mag_s1"], o["mag_s2"], o["feat_max"], 40.0, torch.float64)
    # bit-exact labels given the device's own float inputs (integer work)
    f, m1, m2 = [o[k].cpu().numpy() for k in ("feature", "mag_s1", "mag_s2")]
    for b in range(B):
        np.testing.assert_array_equal(oh[b].cpu().numpy(), O.one_hot(f[b], m1[b], m2[b], 40))
        assert o["feat_max"][b].item() == f[b].max()
    nflip = 0
    for b in range(B):
        inp, lab = O.featurize(*utts[b], n_fft, hop, T, int(starts[b]), 40, "chimera++")
        inpp, labp = O.featurize(*utts[b], n_fft, hop, T, int(starts[b]), 40, "phase")
        scale = lab[1].max()
        np.testing.assert_allclose(o["mag_mix"][b].cpu().numpy(), lab[1], atol=3e-6 * scale)
        np.testing.assert_allclose(o["mag_s1"][b].cpu().numpy(), lab[2], atol=3e-6 * scale)
        np.testing.assert_allclose(o["mag_s2"][b].cpu().numpy(), lab[3], atol=3e-6 * scale)
        np.testing.assert_allclose(o["ph_mix"][b].cpu().numpy(), inpp[1], atol=3e-6 * scale)
        np.testing.assert_allclose(o["ph_s1"][b].cpu().numpy(), labp[4], atol=3e-6 * scale)
        # log-magnitude: compare where the magnitude is well above the 1e-7 floor's rounding noise
        big = lab[1] > 1e-3 * scale
        np.testing.assert_allclose(f[b][big], inp[0][big], atol=2e-5)
        # cos of the phase difference is ill-conditioned for tiny bins: weight by magnitude (it multiplies mag_s in PSA)
        w = lab[2] / scale
        assert np.abs((o["cos_s1"][b].cpu().numpy() - lab[4]) * w).max() < 1e-5
        nflip += int((oh[b].cpu().numpy() != lab[0]).any(-1).sum())
    # end-to-end labels from waveforms: only near-threshold / near-tie bins may differ (float rounding)
    assert nflip <= 1e-4 * B * T * (n_fft // 2 + 1) + 2, nflip


@pytest.mark.parametrize("n_fft,hop,ns", [(256, 64, 32000), (512, 128, 20000), (64, 16, 999)])
def test_istft_masked_vs_oracle(lib, n_fft, hop, ns):
    rng = np.random.RandomState(ns)
    B, S = 2, 2
    specs = [O.stft(O.synth_utterance(i, ns)[0], n_fft, hop) for i in range(B)]
    frames, F = specs[0].shape
    masks = rng.uniform(size=(B, S, frames, F)).astype(np.float32)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    re = cu(np.stack([s.real for s in specs])); im = cu(np.stack([s.imag for s in specs]))
    got = lib.istft_masked(re, im, cu(masks), n_fft, hop, ns).cpu().numpy()
    for b in range(B):
        ref = O.masked_istft(specs[b].real, specs[b].imag, masks[b], hop, ns)
        np.testing.assert_allclose(got[b], ref, atol=2e-6 * max(1e-3, np.abs(ref).max()) + 1e-7)
    # round trip with an all-ones mask reproduces the waveform (size-independent property)
    got = lib.istft_masked(re, im, None, n_fft, hop, ns).cpu().numpy()
    for b in range(B):
        w = O.synth_utterance(b, ns)[0]
        assert np.abs(got[b, 0] - w).max() < 1e-5 * np.abs(w).max() + 1e-7


def test_kmeans_masks_match_sklearn_partition(cuda_device):
    """VAD + 2-means on unit-norm embeddings (deep_clustering/evaluate.py:36-41) against the reference's own tool
    (sklearn KMeans, random_state=0) up to the label permutation."""
    from sklearn.cluster import KMeans
    from onssen_b200 import _lib
    rng = np.random.RandomState(2)
    frames, F, D = 60, 33, 20
    dirs = rng.standard_normal((2, D)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    which = rng.uniform(size=(frames, F)) < 0.4
    emb = dirs[which.astype(int)] + 0.25 * rng.standard_normal((frames, F, D))
    emb = (emb / np.linalg.norm(emb, axis=-1, keepdims=True)).astype(np.float32)
    feature = rng.uniform(-4, 0, size=(frames, F)).astype(np.float32)
    active = feature >= feature.max() - 2.0
    masks, labels = _lib.kmeans_masks(torch.from_numpy(emb).to(cuda_device), torch.from_numpy(feature).to(cuda_device), 2,
                                      40.0, want_labels=True)
    masks, labels = masks.cpu().numpy(), labels.cpu().numpy()
    assert ((labels >= 0) == active).all()                      # VAD is exact (same fp32 compare)
    want = KMeans(n_clusters=2, random_state=0, n_init=10).fit_predict(emb[active])
    got = labels[active]
    agree = max((got == want).mean(), (got == 1 - want).mean())
    assert agree > 0.999, agree
    assert (masks[0][active] == got).all() and (masks[1][active] == 1 - got).all() and (masks[:, ~active] == 0).all()


def test_recurrence_batch_limits_and_single_utterance(lib):
    """edge sizes of the persistent launch: B=1 works, the largest single-launch batch works, and larger batches run
    as column chunks (torch.nn.LSTM has no batch limit); utterances are independent, so any column's output must not
    depend on the batch it was processed with (16- / 32-column tiles, different slice plans)."""
    H, T, I = 600, 5, 129
    Hp = lib.hp_of(H)
    rng = np.random.RandomState(0)
    k = 1 / np.sqrt(H)
    mkw = lambda *s: torch.from_numpy(rng.uniform(-k, k, s).astype(np.float32)).cuda()
    wf = (mkw(4 * H, I), mkw(4 * H, H), mkw(4 * H), mkw(4 * H))
    wr = (mkw(4 * H, I), mkw(4 * H, H), mkw(4 * H), mkw(4 * H))
    _, whh_p, _ = lib.lstm_pack_layer(wf, wr, H, I, False, 0)
    bmax = (lib.load().onssen_num_sms() // (2 * (Hp // 32))) * 32     # slices that fit x 32 columns (96 on 148 SMs)
    assert bmax >= 64
    Bbig = 2 * bmax + 24
    torch.manual_seed(0)
    gates_all = torch.randn(T, Bbig, 8 * Hp, device="cuda")

    def run(cols):
        B = len(cols)
        gates = gates_all[:, cols].contiguous().view(T * B, 8 * Hp)
        y_f = torch.full((T * B, 2 * Hp), float("nan"), device="cuda")
        lib.blstm_rec_fwd(gates, whh_p, B, T, H, None, y_f)
        torch.cuda.synchronize()
        assert torch.isfinite(y_f).all(), B
        return y_f.view(T, B, 2 * Hp)

    y_big = run(list(range(Bbig)))                       # three chunks
    for cols in [[0], [5], list(range(3, 3 + 32)), list(range(bmax)), list(range(40, 40 + 64)),
                 list(range(Bbig - 7, Bbig))]:
        y = run(cols)
        assert (y - y_big[:, cols]).abs().max().item() < 2e-6, len(cols)
