"""End-to-end plugin pipeline on the device: wav files -> loader (device featurizer) -> model -> loss,
validation/checkpoint through `trainer`, evaluation through `tester_*` (masked iSTFT)."""
import os

import numpy as np
import pytest
import torch

from oracle import onssen_oracle as O

pytestmark = pytest.mark.gpu


def _make_corpus(tmp_path, n_utt=3):
    from scipy.io import wavfile
    utts = []
    for part in ("tr", "cv", "tt"):
        for sub in ("mix", "s1", "s2"):
            os.makedirs(tmp_path / "wav8k" / "min" / part / sub, exist_ok=True)
    for i in range(n_utt):
        ns = 9000 + 640 * i
        mix, s1, s2 = O.synth_utterance(i, ns)
        q = lambda x: np.clip(np.round(x * 32768), -32768, 32767).astype(np.int16)
        for part in ("tr", "cv", "tt"):
            for sub, x in (("mix", mix), ("s1", s1), ("s2", s2)):
                wavfile.write(str(tmp_path / "wav8k" / "min" / part / sub / f"u{i}.wav"), 8000, q(x))
        utts.append(tuple(q(x).astype(np.float32) / 32768 for x in (mix, s1, s2)))
    return utts


def test_loader_matches_oracle_featurizer(cuda_device, tmp_path):
    import onssen_b200 as ob
    utts = _make_corpus(tmp_path)
    fo = dict(data_path=str(tmp_path), batch_size=3, frame_length=100, sampling_rate=8000, window_size=256, hop_size=64,
              db_threshold=40)
    loader = ob.data.wsj0_2mix_dataloader("chimera++", fo, "cv", cuda_device)
    loader.shuffle = False
    np.random.seed(11)
    inp, lab = next(iter(loader))
    np.random.seed(11)   # the loader draws one np.random.randint per utterance, like wsj0_2mix.py:125
    assert len(inp) == 1 and len(lab) == 6 and inp[0].shape == (3, 100, 129)
    for b, (mix, s1, s2) in enumerate(utts):
        start = np.random.randint(O.num_crop_starts(len(mix), 64, 100))
        ri, rl = O.featurize(mix, s1, s2, 256, 64, 100, start, 40, "chimera++")
        scale = rl[1].max()
        np.testing.assert_allclose(lab[1][b].cpu().numpy(), rl[1], atol=3e-6 * scale)     # exact same frames
        np.testing.assert_allclose(lab[2][b].cpu().numpy(), rl[2], atol=3e-6 * scale)
        assert (lab[0][b].cpu().numpy() != rl[0]).any(-1).mean() < 1e-3


def test_trainer_validate_and_tester_eval(cuda_device, tmp_path):
    import onssen_b200 as ob
    _make_corpus(tmp_path)
    A = ob.utils.AttrDict
    args = A({"model_name": "chimera", "device": str(cuda_device), "num_epoch": 1, "checkpoint_path": str(tmp_path / "ckpt"),
              "feature_options": A(dict(data_path=str(tmp_path), batch_size=2, frame_length=100, sampling_rate=8000,
                                        window_size=256, hop_size=64, db_threshold=40)),
              "optimizer_options": A({"name": "adam", "lr": 1e-3}), "verbose": False})
    torch.manual_seed(0)
    args.model = ob.nn.chimera(129, 64, 2, 20).to(cuda_device)
    args.train_loader = ob.data.wsj0_2mix_dataloader("chimera", args.feature_options, "tr", cuda_device)
    args.valid_loader = ob.data.wsj0_2mix_dataloader("chimera", args.feature_options, "cv", cuda_device)
    args.test_loader = ob.data.wsj0_2mix_dataloader("chimera", args.feature_options, "tt", cuda_device)
    args.optimizer = ob.utils.build_optimizer(args.model.parameters(), args.optimizer_options)
    args.loss_fn = ob.loss.loss_chimera_msa
    tr = ob.utils.trainer(args)
    w0 = args.model.fc_mi.weight.detach().clone()
    l = tr.train(0)                                       # 2 steps: device-side loss accumulation, clip + Adam kernels
    assert np.isfinite(l) and not torch.equal(w0, args.model.fc_mi.weight.detach())
    v = tr.validate(0)
    assert np.isfinite(v) and os.path.exists(tmp_path / "ckpt" / "final.mdl")
    saved = torch.load(tmp_path / "ckpt" / "final.mdl", weights_only=False)
    assert set(saved) == {"model", "epoch", "optimizer", "cv_loss", "early_stop_count"}     # train.py:111-122
    assert "rnn.weight_hh_l1_reverse" in saved["model"] and "fc_mi.bias" in saved["model"]
    sdr = ob.utils.tester_chimera(args).eval()
    assert np.isfinite(sdr)
    # eval labels: stft_r, stft_i, sig_ref and an all-ones mask reproduces the mixture through the device iSTFT
    inp, lab = next(iter(args.test_loader))
    ones = torch.ones(1, 1, *lab[0].shape[1:], device=cuda_device)
    t = ob.utils.tester_chimera(args)
    rec = t.masked_istft(lab[0], lab[1], ones, lab[2].shape[2])
    mix = lab[2][0].sum(0)
    assert (rec[0, 0] - mix).abs().max().item() < 2e-4 * mix.abs().max().item() + 1e-4   # s1+s2 == mix up to int16 rounding
    # resume: model, epoch AND the Adam state continue (taken from the optimizer object the reference's checkpoint pickles)
    args2 = A(dict(args))
    args2["resume_from_checkpoint"] = "True"
    args2["optimizer"] = ob.utils.build_optimizer(args.model.parameters(), args.optimizer_options)
    tr2 = ob.utils.trainer(args2)
    p0 = next(iter(args.model.parameters()))
    assert tr2.epoch == 1 and int(tr2.optimizer.state[p0]["step"]) == int(args.optimizer.state[p0]["step"]) > 0
    assert torch.equal(tr2.optimizer.state[p0]["exp_avg"], args.optimizer.state[p0]["exp_avg"])


@pytest.mark.parametrize("model_name", ["dc", "chimera"])
def test_batched_evaluation_equals_batch_one(cuda_device, tmp_path, model_name):
    """`eval_batch_size` > 1: utterances of different lengths zero-padded into one batch, per-utterance frame counts in
    the recurrence (state held at zero beyond an utterance's end, so the reverse direction starts at its own last
    frame).  Every utterance must get the outputs -- and the SI-SDR -- of the reference's batch-1 loop
    (egs/wsj0-2mix/*/evaluate.py)."""
    import onssen_b200 as ob
    _make_corpus(tmp_path, n_utt=3)                      # 9000, 9640, 10280 samples: 141 / 151 / 161 frames
    torch.manual_seed(3)
    fo = dict(data_path=str(tmp_path), batch_size=2, frame_length=100, sampling_rate=8000, window_size=256, hop_size=64,
              db_threshold=40)
    model = (ob.nn.deep_clustering(129, 64, 2, 20) if model_name == "dc" else ob.nn.chimera(129, 64, 2, 20))
    model = model.to(cuda_device).eval()
    one = ob.data.wsj0_2mix_dataloader(model_name, fo, "tt", cuda_device)
    many = ob.data.wsj0_2mix_dataloader(model_name, dict(fo, eval_batch_size=3), "tt", cuda_device)
    singles = list(one)
    (inp, lab), = list(many)
    assert inp[0].shape[0] == 3 and len(lab) == 4 and lab[3].tolist() == [9000, 9640, 10280]
    frames = (1 + lab[3].to(torch.int64) // 64).tolist()
    with torch.no_grad():
        model.frame_lengths = torch.tensor(frames, dtype=torch.int32)
        out = model(inp)
        model.frame_lengths = None
        for b, (inp1, lab1) in enumerate(singles):
            out1 = model(inp1)
            fb = frames[b]
            assert inp1[0].shape[1] == fb
            assert torch.equal(inp[0][b, :fb], inp1[0][0])                       # same features as the batch-1 loader
            for o, o1 in zip(out, out1):
                assert (o[b, :fb] - o1[0]).abs().max().item() < 2e-6             # same arithmetic per utterance
    ckpt = tmp_path / "ckpt"
    os.makedirs(ckpt, exist_ok=True)
    torch.save({"model": model.state_dict()}, ckpt / "final.mdl")
    T = ob.utils.tester_dc if model_name == "dc" else ob.utils.tester_chimera
    args = dict(model_name=model_name, device=str(cuda_device), model=model, checkpoint_path=str(ckpt), feature_options=fo)
    sdr1 = T(dict(args, test_loader=one)).eval()
    sdr3 = T(dict(args, test_loader=many)).eval()
    assert np.isfinite(sdr1) and abs(sdr1 - sdr3) < 1e-3, (sdr1, sdr3)


def test_loader_decodes_and_resamples_on_the_device(cuda_device, tmp_path):
    """16 kHz stereo int16 files with feature_options.sampling_rate = 8000: raw PCM staged by the thread pool, int16 ->
    float mono and the polyphase resampling on the device; waveforms must equal scipy.signal.resample_poly of the host
    decode (feature_utils.py:15-19 behaviour with the documented filter), features the oracle's of those waveforms."""
    from scipy.io import wavfile
    from scipy.signal import resample_poly
    import onssen_b200 as ob
    from onssen_b200.data import wavio
    for sub in ("mix", "s1", "s2"):
        os.makedirs(tmp_path / "wav8k" / "min" / "tr" / sub, exist_ok=True)
    rng = np.random.RandomState(5)
    host = []
    for i in range(3):
        n = 20000 + 1234 * i
        sig = {}
        for sub in ("mix", "s1", "s2"):
            x = (rng.standard_normal((n, 2)) * 3000).astype(np.int16)
            t = np.arange(n) / 16000.0
            x[:, 0] += (8000 * np.sin(2 * np.pi * (300 + 100 * i) * t)).astype(np.int16)
            wavfile.write(str(tmp_path / "wav8k" / "min" / "tr" / sub / f"u{i}.wav"), 16000, x)
            mono = (x.astype(np.float32) / 32768.0).mean(axis=1)
            sig[sub] = resample_poly(mono, 1, 2).astype(np.float32)
        host.append(sig)
    names = [[tuple(str(tmp_path / "wav8k" / "min" / "tr" / sub / f"u{i}.wav") for sub in ("mix", "s1", "s2"))
              for i in range(3)]]
    staged = next(iter(wavio.PcmStager(names, workers=4, depth=1)))
    assert staged["pcm"].dtype == torch.int16 and staged["pcm"].is_pinned() and staged["channels"] == 2
    (mix, s1, s2), lengths = wavio.device_waveforms(staged, 8000, cuda_device)
    assert lengths.tolist() == [len(h["mix"]) for h in host]
    for b, h in enumerate(host):
        for dev, key in ((mix, "mix"), (s1, "s1"), (s2, "s2")):
            got = dev[b, :lengths[b]].cpu().numpy()
            assert np.abs(got - h[key]).max() < 2e-6 * np.abs(h[key]).max() + 1e-7
            assert float(dev[b, lengths[b]:].abs().max() if dev.shape[1] > lengths[b] else 0.0) == 0.0
    # and through the loader factory (same files): features of the resampled waveforms
    fo = dict(data_path=str(tmp_path), batch_size=3, frame_length=60, sampling_rate=8000, window_size=256, hop_size=64,
              db_threshold=40, num_workers=3, prefetch_batches=1)
    loader = ob.data.wsj0_2mix_dataloader("dc", fo, "tr", cuda_device)
    loader.shuffle = False
    np.random.seed(3)
    inp, lab = next(iter(loader))
    np.random.seed(3)
    for b, h in enumerate(host):
        start = np.random.randint(O.num_crop_starts(len(h["mix"]), 64, 60))
        ri, rl = O.featurize(h["mix"], h["s1"], h["s2"], 256, 64, 60, start, 40, "dc")
        np.testing.assert_allclose(lab[1][b].cpu().numpy(), rl[1], atol=1e-5 * rl[1].max())


def test_daps_loader_segments_match_oracle(cuda_device, tmp_path):
    """daps_enhance loader: a whole recording featurized on the device, consecutive frame_length segments, batched;
    every item must equal the oracle's features of the same recording at the same frames."""
    import random
    from scipy.io import wavfile
    import onssen_b200 as ob
    os.makedirs(tmp_path / "clean", exist_ok=True)
    os.makedirs(tmp_path / "noisy", exist_ok=True)
    q = lambda x: np.clip(np.round(x * 32768), -32768, 32767).astype(np.int16)
    recs, lines = {}, []
    for i in range(2):
        ns = 16000 + 1280 * i
        mix, s1, _ = O.synth_utterance(40 + i, ns)
        fn = str(tmp_path / "noisy" / f"f{i}_script{i}_ipad.wav")
        wavfile.write(fn, 16000, q(mix))
        wavfile.write(str(tmp_path / "clean" / f"f{i}_script{i}_clean.wav"), 16000, q(s1))
        recs[fn] = (q(mix).astype(np.float32) / 32768, q(s1).astype(np.float32) / 32768)
        lines.append(fn)
    (tmp_path / "train").write_text("\n".join(lines) + "\n")
    T, n_fft, hop = 50, 512, 128
    fo = dict(data_path=str(tmp_path), batch_size=2, frame_length=T, sampling_rate=16000, window_size=n_fft, hop_size=hop)
    random.seed(3)
    ld = ob.data.daps_enhance_dataloader(2, fo, "train", cuda_device)
    order = []
    orig = ld.cursor.featurize
    ld.cursor.featurize = lambda p: (order.append(p), orig(p))[1]
    batches = list(ld)
    assert len(batches) == 2 and batches[0][0][0].shape == (2, T, n_fft // 2 + 1)
    items = [(b[0][0][k], b[0][1][k], b[1][0][k], b[1][1][k]) for b in batches for k in range(2)]
    # 1 + 16000/128 = 126 frames -> 2 segments per recording, then the next file
    want = []
    for fn in order:
        noisy, clean = recs[fn]
        sn, sc = O.stft(noisy, n_fft, hop), O.stft(clean, n_fft, hop)
        for k in range(sn.shape[0] // T):
            sl = slice(k * T, (k + 1) * T)
            want.append((O.log_magnitude(sn)[sl], np.abs(sn)[sl], np.abs(sc)[sl], O.cos_difference(sn, sc)[sl], np.abs(sc)[sl]))
    for got, w in zip(items, want):
        assert np.abs(got[0].cpu().numpy() - w[0]).max() < 2e-3
        assert np.abs(got[1].cpu().numpy() - w[1]).max() < 1e-4
        assert np.abs(got[2].cpu().numpy() - w[2]).max() < 1e-4
        assert (np.abs(got[3].cpu().numpy() - w[3]) * w[4]).max() < 2e-4   # cos is ill-conditioned on tiny bins


def _make_edinburgh(tmp_path, n=4, rate=16000):
    from scipy.io import wavfile
    for sub in ("noisy_trainset_28spk_wav", "clean_trainset_28spk_wav"):
        os.makedirs(tmp_path / sub, exist_ok=True)
    names, sigs = [f"p{i}.wav" for i in range(n)], []
    q = lambda x: np.clip(np.round(x * 32768), -32768, 32767).astype(np.int16)
    for i, nm in enumerate(names):
        mix, clean, _ = O.synth_utterance(50 + i, 9000 + 700 * i)
        wavfile.write(str(tmp_path / "clean_trainset_28spk_wav" / nm), rate, q(clean))
        wavfile.write(str(tmp_path / "noisy_trainset_28spk_wav" / nm), rate, q(mix))
        sigs.append((q(mix).astype(np.float32) / 32768, q(clean).astype(np.float32) / 32768))
    for part in ("train", "validation"):
        (tmp_path / part).write_text("\n".join(names) + "\n")
    return names, sigs


def test_edinburgh_loader_matches_oracle(cuda_device, tmp_path):
    """edinburgh_tts.py:68-112: noisy / clean pairs, "speaker 2" = noisy - clean, fixed crop [:frame_length] after the
    tiling, chimera++ label layout; every batch item against the oracle's featurizer of the same files."""
    import onssen_b200 as ob
    names, sigs = _make_edinburgh(tmp_path)
    fo = dict(data_path=str(tmp_path), batch_size=4, frame_length=50, sampling_rate=16000, window_size=512, hop_size=128,
              db_threshold=40)
    ld = ob.data.edinburgh_tts_dataloader("chimera++", fo, "train", cuda_device)
    order = list(ld.file_list)
    import random
    random.seed(1)
    inp, lab = next(iter(ld))
    random.seed(1)
    idx = list(range(len(order))); random.shuffle(idx)           # the loader's own shuffle of this epoch
    assert len(inp) == 1 and len(lab) == 6 and inp[0].shape == (4, 50, 257)
    for b, j in enumerate(idx):
        mix, clean = sigs[names.index(os.path.basename(order[j]))]
        ri, rl = O.featurize(mix, clean, mix - clean, 512, 128, 50, 0, 40, "chimera++")
        scale = rl[1].max()
        for k in (1, 2, 3):
            np.testing.assert_allclose(lab[k][b].cpu().numpy(), rl[k], atol=4e-6 * scale)
        big = rl[1] > 1e-3 * scale
        np.testing.assert_allclose(inp[0][b].cpu().numpy()[big], ri[0][big], atol=3e-5)
        assert (lab[0][b].cpu().numpy() != rl[0]).any(-1).mean() < 1e-3
    dc = ob.data.edinburgh_tts_dataloader("dc", fo, "train", cuda_device)
    i2, l2 = next(iter(dc))
    assert len(l2) == 1 and l2[0].shape == (4, 50, 257, 2)        # the reference's "dc" branch yields [one_hot] only (:92)


@pytest.mark.parametrize("eg", ["wsj0-2mix/deep_clustering", "wsj0-2mix/chimera/psa", "wsj0-2mix/phase-net",
                                "edinburgh_tts", "daps"])
def test_egs_scripts_run(cuda_device, tmp_path, eg):
    """every egs/*/run.py executed as a user would (`python run.py -c config.json`) on a tiny synthetic corpus: its
    committed config with data path / sizes shrunk, one epoch, a checkpoint and (deep clustering) an SI-SDR at the end."""
    import json
    import subprocess
    import sys
    from scipy.io import wavfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = json.load(open(os.path.join(root, "egs", eg, "config.json")))
    fo = cfg["feature_options"]
    fo.update(data_path=str(tmp_path), batch_size=2, frame_length=40)
    cfg["model_options"].update(hidden_dim=32, num_layers=2)
    cfg.update(num_epoch=1, checkpoint_path=str(tmp_path / "ckpt"), device=str(cuda_device), verbose=False)
    if eg.startswith("wsj0"):
        _make_corpus(tmp_path)
        if fo["sampling_rate"] != 8000:                                     # phase-net config: 16 kHz / 512 / 128
            fo.update(sampling_rate=8000)                                   # the corpus helper writes 8 kHz files
    elif eg == "edinburgh_tts":
        _make_edinburgh(tmp_path, rate=fo["sampling_rate"])
    else:
        os.makedirs(tmp_path / "clean", exist_ok=True)
        q = lambda x: np.clip(np.round(x * 32768), -32768, 32767).astype(np.int16)
        lines = []
        for i in range(2):
            mix, clean, _ = O.synth_utterance(80 + i, fo["hop_size"] * 130)
            wavfile.write(str(tmp_path / f"f{i}_script{i}_iphone.wav"), fo["sampling_rate"], q(mix))
            wavfile.write(str(tmp_path / "clean" / f"f{i}_script{i}_clean.wav"), fo["sampling_rate"], q(clean))
            lines.append(str(tmp_path / f"f{i}_script{i}_iphone.wav"))
        for part in ("train", "validation"):
            (tmp_path / part).write_text("\n".join(lines) + "\n")
        cfg.update(train_num_batch=2, validate_num_batch=1)
    cpath = tmp_path / "config.json"
    cpath.write_text(json.dumps(cfg))
    res = subprocess.run([sys.executable, os.path.join(root, "egs", eg, "run.py"), "-c", str(cpath)], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert os.path.exists(tmp_path / "ckpt" / "final.mdl")
    assert "Model training is finished." in res.stdout or cfg.get("verbose") is False
    if eg.endswith("deep_clustering"):
        assert "SI-SDR" in res.stdout
