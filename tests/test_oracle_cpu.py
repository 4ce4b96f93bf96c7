"""Oracle vs the reference's own outputs (golden fixtures) and vs torch.stft -- CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import onssen_oracle as O


@pytest.mark.parametrize("name", ["small", "mid"])
def test_oracle_deep_clustering_matches_reference(name):
    p, g = load_golden(f"dc_{name}.npz")
    B, T, F, H, L, D = g["cfg"]
    emb, = O.deep_clustering_forward(p, [g["feature"]], L, training=False)
    np.testing.assert_allclose(emb, g["emb_eval"], atol=2e-5)
    loss = O.loss_dc([emb], [g["one_hot"], g["mag_mix"]])
    assert loss.shape == (B, B)
    np.testing.assert_allclose(loss, g["loss_eval"], rtol=2e-5)
    emb_t, = O.deep_clustering_forward(p, [g["feature"]], L, training=True)
    np.testing.assert_allclose(emb_t, g["emb_train"], atol=2e-5)
    np.testing.assert_allclose(O.loss_dc([emb_t], [g["one_hot"], g["mag_mix"]]), g["loss_train"], rtol=2e-5)
    # running statistics update (momentum 0.1, unbiased variance)
    y = O.blstm_stack(g["feature"], p, "rnn.", L)
    _, rm, rv = O.batchnorm_bt(y, p, "bn.", True)
    np.testing.assert_allclose(rm, g["bn_rm_after"], atol=1e-6)
    np.testing.assert_allclose(rv, g["bn_rv_after"], rtol=1e-5)


@pytest.mark.parametrize("name", ["small", "mid"])
def test_oracle_chimera_matches_reference(name):
    p, g = load_golden(f"chimera_{name}.npz")
    L = g["cfg"][4]
    e, ma, mb = O.chimera_forward(p, [g["feature"]], L)
    np.testing.assert_allclose(e, g["emb"], atol=2e-5)
    np.testing.assert_allclose(ma, g["mask_a"], atol=2e-6)
    np.testing.assert_allclose(mb, g["mask_b"], atol=2e-6)
    lab4 = [g["one_hot"], g["mag_mix"], g["mag_s1"], g["mag_s2"]]
    np.testing.assert_allclose(O.loss_chimera_msa([e, ma, mb], lab4), g["loss_msa"], rtol=2e-5)
    np.testing.assert_allclose(O.loss_chimera_psa([e, ma, mb], lab4 + [g["cos_s1"], g["cos_s2"]]), g["loss_psa"],
                               rtol=2e-5)


@pytest.mark.parametrize("name", ["small", "mid"])
def test_oracle_enhance_matches_reference(name):
    p, g = load_golden(f"enhance_{name}.npz")
    L = g["cfg"][4]
    c_eval, = O.enhance_forward(p, [g["feature"], g["mag_noisy"]], L, training=False)
    np.testing.assert_allclose(c_eval, g["clean_eval"], atol=2e-5)
    c_tr, = O.enhance_forward(p, [g["feature"], g["mag_noisy"]], L, training=True)
    np.testing.assert_allclose(c_tr, g["clean_train"], atol=2e-5)
    np.testing.assert_allclose(O.loss_mask_msa([c_eval], [g["mag_clean"], g["cos_diff"]]), g["loss_msa"], rtol=1e-5)
    sig = 1 / (1 + np.exp(-c_eval))
    np.testing.assert_allclose(O.loss_mask_psa([sig], [g["mag_noisy"], g["mag_clean"], g["cos_diff"]]),
                               g["loss_psa"], rtol=1e-5)


@pytest.mark.parametrize("n_fft,hop,ns", [(256, 64, 32000), (512, 128, 64000), (1024, 256, 64000), (64, 16, 1000)])
def test_oracle_stft_matches_torch(n_fft, hop, ns):
    mix, _, _ = O.synth_utterance(3, ns)
    spec = O.stft(mix, n_fft, hop)
    assert spec.shape == (1 + ns // hop, n_fft // 2 + 1) and spec.dtype == np.complex64
    ref = torch.stft(torch.from_numpy(mix), n_fft, hop, window=torch.hann_window(n_fft, periodic=True), center=True,
                     pad_mode="reflect", return_complex=True).T.numpy()
    scale = np.abs(ref).max()
    assert np.abs(spec - ref).max() <= 2e-6 * scale
    # istft round trip (window-sum-square normalisation) and agreement with torch.istft
    y = O.istft(spec, hop, ns)
    assert np.abs(y - mix).max() <= 1e-5 * np.abs(mix).max() + 1e-7
    yt = torch.istft(torch.from_numpy(spec.T.copy()), n_fft, hop, window=torch.hann_window(n_fft, periodic=True),
                     center=True, length=ns).numpy()
    assert np.abs(y - yt).max() <= 1e-5 * np.abs(mix).max() + 1e-7


@pytest.mark.parametrize("n_fft,hop,ns", [(256, 64, 8000), (512, 128, 9000), (1024, 256, 20000)])
def test_oracle_stft_matches_scipy(n_fft, hop, ns):
    """second, independent restatement of the librosa boundary (rows a1 / a19 are unpinned: librosa is not installable):
    scipy.signal.stft with the periodic Hann window, 'even' (= reflect) boundary extension and its 1/sum(window) scaling
    undone must give the oracle's spectrum; scipy.signal.istft the oracle's waveform."""
    from scipy import signal
    x = O.synth_utterance(3, ns)[0]
    spec = O.stft(x, n_fft, hop)
    win = signal.get_window("hann", n_fft, fftbins=True)
    _, _, z = signal.stft(x.astype(np.float64), window=win, nperseg=n_fft, noverlap=n_fft - hop, boundary="even",
                          padded=False, return_onesided=True)
    z = z.T * win.sum()
    assert z.shape[0] >= spec.shape[0]
    assert np.abs(spec - z[:spec.shape[0]]).max() <= 2e-6 * np.abs(z).max()
    y = O.istft(spec, hop, ns)
    _, y2 = signal.istft((spec / win.sum()).T, window=win, nperseg=n_fft, noverlap=n_fft - hop, boundary=True,
                         input_onesided=True)
    n = min(len(y), len(y2)) - n_fft                     # scipy stops at the last whole hop: compare the common part
    assert n > 0 and np.abs(y[:n] - y2[:n]).max() <= 1e-6 * np.abs(x).max() + 1e-7


def test_oracle_frame_indexing_and_labels():
    # 4 s @ 8 kHz -> 501 frames, crop start exclusive bound 101 (np.random.randint(frames - T))
    assert O.num_crop_starts(32000, 64, 400) == 101
    # 16 kHz, n_fft 1024 / hop 256: 251 frames <= 400 -> tiled x2 = 502 -> bound 102
    assert O.num_crop_starts(64000, 256, 400) == 102
    spec = (np.arange(5 * 3).reshape(5, 3) + 0j).astype(np.complex64)
    tiled = O.tile_and_crop(spec, 7, 2)           # 5 <= 7 -> times = 2 -> 10 frames, rows 2..8
    np.testing.assert_array_equal(tiled[:, 0].real, (np.arange(2, 9) % 5) * 3)
    # labels: ties -> speaker 0, strict < for the VAD
    feat = np.array([[0.0, -1.0, -2.0, -2.0000002]], dtype=np.float32)
    m1 = np.array([[1.0, 2.0, 3.0, 1.0]], dtype=np.float32)
    m2 = np.array([[1.0, 3.0, 1.0, 5.0]], dtype=np.float32)
    oh = O.one_hot(feat, m1, m2, 40)
    np.testing.assert_array_equal(oh[0], [[1, 0], [0, 1], [1, 0], [0, 0]])
    assert oh.dtype == np.float64


def test_oracle_featurize_layouts():
    mix, s1, s2 = O.synth_utterance(0, 8000)
    for name, nin, nlab in [("dc", 1, 2), ("chimera", 1, 4), ("chimera++", 1, 6), ("phase", 2, 6)]:
        inp, lab = O.featurize(mix, s1, s2, 256, 64, 100, 5, 40, name)
        assert len(inp) == nin and len(lab) == nlab
        assert inp[0].shape == (100, 129) and lab[0].shape == (100, 129, 2)
    active = lab[0].sum(-1).mean()
    assert 0.2 < active < 0.98           # the synthetic mixtures do exercise the VAD


def test_oracle_loss_phase_repaired_runs():
    rng = np.random.RandomState(0)
    B, T, F, D = 2, 6, 5, 4
    emb = O.l2_normalize(rng.standard_normal((B, T, F, D)).astype(np.float32))
    ma = rng.uniform(size=(B, T, F)).astype(np.float32)
    ph = lambda: O.l2_normalize(rng.standard_normal((B, T, F, 2)).astype(np.float32))
    mags = [np.abs(rng.standard_normal((B, T, F))).astype(np.float32) for _ in range(3)]
    oh = np.zeros((B, T, F, 2)); oh[..., 0] = 1
    out = O.loss_phase([emb, ma, 1 - ma, ph(), ph()], [oh, mags[0], mags[1], mags[2], ph(), ph()])
    assert out.shape == (B, B) and np.isfinite(out).all()


@pytest.mark.parametrize("name", ["small", "mid"])
def test_oracle_phase_net_matches_torch_restatement(name):
    """phase_net / loss_phase: numpy oracle vs the fixtures of the torch restatement with the same repairs
    (oracle/make_golden.py main_phase; the reference itself raises -> parity unpinned for these rows)."""
    p, g = load_golden(f"phase_{name}.npz")
    B, T, F, H, L, D = [int(v) for v in g["cfg"]]
    out = O.phase_net_forward(p, [g["feature"], g["x_phase"]], L, training=True)
    for got, key in zip(out, ["emb", "mask_a", "mask_b", "phase_a", "phase_b"]):
        assert np.abs(got - g[key]).max() < 2e-5, key
    loss = O.loss_phase([g[k] for k in ("emb", "mask_a", "mask_b", "phase_a", "phase_b")],
                        [g[k] for k in ("one_hot", "mag_mix", "mag_s1", "mag_s2", "phase_s1", "phase_s2")])
    assert np.allclose(loss, g["loss"], rtol=2e-5, atol=1e-4)


def test_oracle_clip_adam_matches_torch():
    """the optimiser restatement against the reference's actual implementation (torch clip_grad_norm_ + optim.Adam)"""
    import torch
    rng = np.random.RandomState(5)
    shapes = [(7, 5), (33,), (4, 3, 2), (1,)]
    ps = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    tp = [torch.nn.Parameter(torch.from_numpy(p.copy())) for p in ps]
    opt = torch.optim.Adam(tp, lr=1e-3)
    m = [np.zeros_like(p) for p in ps]
    v = [np.zeros_like(p) for p in ps]
    for step in range(1, 5):
        gs = [(rng.standard_normal(s) * (3.0 if step % 2 else 0.1)).astype(np.float32) for s in shapes]
        for t, g in zip(tp, gs):
            t.grad = torch.from_numpy(g.copy())
        tn = float(torch.nn.utils.clip_grad_norm_(tp, 5))
        opt.step()
        n = O.clip_adam_step(ps, [g.copy() for g in gs], m, v, step)
        assert abs(n - tn) < 1e-4 * tn
        for a, t in zip(ps, tp):
            assert np.abs(a - t.detach().numpy()).max() < 2e-6


# ------------------------------------------------------------------------------------------------ rows a2-a7
@pytest.mark.parametrize("name", ["crop", "tile"])
def test_oracle_featurizer_bit_exact_vs_reference_get_feature(name):
    """feat_*.npz come from the reference's OWN wsj0_2mix_dataset.get_feature (oracle/make_golden.py:main_feat; only
    get_stft is substituted).  Integer work (crop index, tiling, labels, VAD) AND the float helpers must match
    bit for bit, dtype included."""
    _, g = load_golden(f"feat_{name}.npz")
    nsample, n_fft, hop, T, seed = (int(v) for v in g["cfg"])
    np.random.seed(seed)
    start = np.random.randint(O.num_crop_starts(nsample, hop, T))          # wsj0_2mix.py:125
    assert start == int(g["crop_start"])
    inp, lab = O.featurize(g["mix"], g["s1"], g["s2"], n_fft, hop, T, start, 40, "chimera++")
    for k, a in zip(("feature", "one_hot", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2"), inp + lab):
        assert a.dtype == g[k].dtype and a.shape == g[k].shape, k
        np.testing.assert_array_equal(a, g[k], err_msg=k)
    inp, lab = O.featurize(g["mix"], g["s1"], g["s2"], n_fft, hop, T, start, 40, "phase")
    for k, a in zip(("feature", "phase_mix", "one_hot", "mag_mix", "mag_s1", "mag_s2", "phase_s1", "phase_s2"), inp + lab):
        assert a.dtype == g[k].dtype and a.shape == g[k].shape, k
        np.testing.assert_array_equal(a, g[k], err_msg=k)
    for model_name, n_lab in (("dc", 2), ("chimera", 4)):
        inp, lab = O.featurize(g["mix"], g["s1"], g["s2"], n_fft, hop, T, start, 40, model_name)
        assert len(inp) == 1 and len(lab) == n_lab
    assert (g["one_hot"].sum(-1) == 0).any() and (g["one_hot"].sum(-1) == 1).any()      # the VAD actually bites
    if name == "tile":
        frames = 1 + nsample // hop
        assert frames <= T                                                            # the tiling branch ran
        np.testing.assert_array_equal(g["mag_mix"][frames - start], g["mag_mix"][0] if start == 0 else
                                      np.abs(O.stft(g["mix"], n_fft, hop))[0])


def test_staged_reference_is_unmodified():
    """oracle/_ref (when staged) holds the reference's files byte for byte (sha256 manifest written at staging time)."""
    import hashlib, json, os
    from oracle import ref_loader
    root = ref_loader.REF_STAGED
    if not os.path.isdir(root):
        pytest.skip("oracle/_ref not staged")
    man = json.load(open(os.path.join(root, "MANIFEST.json")))["sha256"]
    assert "onssen/nn/deep_clustering.py" in man and "onssen/loss/loss_dc.py" in man
    for rel, h in man.items():
        assert hashlib.sha256(open(os.path.join(root, rel), "rb").read()).hexdigest() == h, rel
        if os.path.isfile(os.path.join("/root/reference", rel)):
            assert open(os.path.join("/root/reference", rel), "rb").read() == open(os.path.join(root, rel), "rb").read()
