"""CPU-side checks: the C-ABI library loads and exports every symbol include/onssen_b200.h declares, the
host mirror keeps the reference's plugin surface, and the product path fails loudly without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "onssen_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(onssen_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from onssen_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/onssen_b200.h but not exported"
    # the ctypes table binds exactly the declared surface
    assert sorted(_lib.exported_symbols()) == declared
    _lib.load()
    assert b"sm_100a" in _lib.load().onssen_version()


def test_plugin_surface_matches_reference_names():
    import onssen_b200 as ob
    m = ob.nn.deep_clustering(129, hidden_dim=600, num_layers=3, embedding_dim=20)
    keys = [k for k in m.state_dict() if "num_batches" not in k]
    for l in range(3):
        for suf in ("", "_reverse"):
            for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                assert f"rnn.{n}_l{l}{suf}" in keys
    for k in ("bn.weight", "bn.bias", "bn.running_mean", "bn.running_var", "fc_dc.weight", "fc_dc.bias"):
        assert k in keys
    assert m.fc_dc.weight.shape == (129 * 20, 1200)
    c = ob.nn.chimera(129, 300, 4, 20, num_speaker=2)
    assert c.fc_mi.weight.shape == (258, 600) and "bn.weight" not in c.state_dict()
    for fn in ("loss_dc", "loss_chimera_msa", "loss_chimera_psa"):
        assert callable(getattr(ob.loss, fn))


def test_product_path_fails_loudly_without_gpu():
    import onssen_b200 as ob
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = ob.nn.deep_clustering(9, 8, 1, 4).eval()
    with torch.no_grad(), pytest.raises(ob._lib.OnssenB200Error):
        m([torch.zeros(1, 4, 9)])
    with pytest.raises(AssertionError):
        m([torch.zeros(1, 4, 9), torch.zeros(1)])       # same arity assert as deep_clustering.py:31


def test_attrdict_and_optimizer_shims():
    import onssen_b200 as ob
    a = ob.utils.AttrDict({"model_options": {"input_dim": 129}, "device": "cpu"})
    assert a.model_options.input_dim == 129 and a["model_options"]["input_dim"] == 129
    a.model = 3
    assert a["model"] == 3 and "model" in a
    p = [torch.nn.Parameter(torch.zeros(2))]
    assert isinstance(ob.utils.build_optimizer(p, a.__class__({"name": "adam", "lr": 1e-3})), torch.optim.Adam)
    assert ob.data.num_crop_starts(32000, 64, 400) == 101


@pytest.mark.parametrize("staged", [True, False])
def test_bench_reference_arm_line_contract(monkeypatch, capsys, staged):
    """`bench.py --impl reference` on a shrunken workload: one JSON line with the keys the driver reads.  With
    oracle/_ref staged the arm runs the reference's own torch modules (kind "reference", same config as the GPU arm);
    without it, the numpy oracle port (kind "port")."""
    import argparse
    import json
    import bench
    from oracle import ref_loader
    if staged and not ref_loader.available():
        pytest.skip("oracle/_ref not staged")
    if not staged:
        monkeypatch.setattr(ref_loader, "available", lambda: False)
    monkeypatch.setitem(bench.CFG, "margs", (129, 8, 1, 4))
    monkeypatch.setitem(bench.CFG, "B", 3)
    monkeypatch.setattr(bench, "T_FRAMES", 20)
    monkeypatch.setitem(bench.CFG, "nsample", 7680)
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference(argparse.Namespace(steps=1, warmup=0, gpus=1))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "utterances/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["same_config"] is staged and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["metric"] == bench.METRIC


def test_egs_configs_match_constructor_and_loader_signatures():
    """every egs/*/config.json parses into the AttrDict shim, its model_options are accepted by the model its run.py
    builds, and its feature_options carry what the loader reads"""
    import glob
    import inspect
    import json
    import onssen_b200 as ob
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "egs")
    model_of = {"deep_clustering": ob.nn.deep_clustering, "chimera": ob.nn.chimera, "edinburgh_tts": ob.nn.chimera,
                "daps": ob.nn.enhance, "phase-net": ob.nn.phase_net}
    seen = 0
    for cfg in glob.glob(os.path.join(root, "**", "config.json"), recursive=True):
        args = ob.utils.AttrDict(json.load(open(cfg)))
        key = next(k for k in model_of if k in cfg.replace(os.sep, "/").split("/egs/")[1])
        params = inspect.signature(model_of[key].__init__).parameters
        assert set(args["model_options"]) <= set(params), (cfg, set(args["model_options"]) - set(params))
        fo = args.feature_options
        for k in ("data_path", "batch_size", "frame_length", "sampling_rate", "window_size", "hop_size"):
            assert k in fo, (cfg, k)
        assert args.model_options.input_dim == fo.window_size // 2 + 1
        assert str(args.device).startswith("cuda") and os.path.exists(os.path.join(os.path.dirname(cfg), "run.py"))
        seen += 1
    assert seen == 6


def test_experiment_config_loading(tmp_path):
    import json
    from onssen_b200.utils.experiment import load_config
    (tmp_path / "config.json").write_text(json.dumps({"device": "cuda:0", "model_options": {"input_dim": 129}}))
    (tmp_path / "other.json").write_text(json.dumps({"device": "cuda:1", "model_options": {"input_dim": 257}}))
    a = load_config(str(tmp_path), argv=[])
    assert a.device == "cuda:0" and a.model_options.input_dim == 129 and a["model_options"]["input_dim"] == 129
    b = load_config(str(tmp_path), argv=["-c", str(tmp_path / "other.json")])
    assert b.device == "cuda:1"
    b.model = object()                       # runtime objects are assigned onto the same mapping (run.py:23-29 usage)
    assert "model" in b


def test_optimiser_pointer_tables_layout():
    """device tables of the multi-tensor optimiser kernels: 40-byte records {p, g, m, v, numel} and 16-byte chunk
    records {tensor, start} covering every element exactly once (built on the CPU here)"""
    from onssen_b200 import _lib
    numels = [5, _lib.OPT_CHUNK, _lib.OPT_CHUNK + 1, 3 * _lib.OPT_CHUNK]
    tb = _lib.OptTables(numels, "cpu")
    rows = [(1000 + i, 2000 + i, 3000 + i, 4000 + i) for i in range(len(numels))]
    dev = tb.fill(rows)
    assert dev.shape == (4, 5) and dev.dtype == torch.int64
    assert dev[:, :4].tolist() == [list(r) for r in rows] and dev[:, 4].tolist() == numels
    ch = tb.chunks.tolist()
    assert tb.nchunks == len(ch) == 1 + 1 + 2 + 3
    covered = {i: 0 for i in range(len(numels))}
    for t, start in ch:
        assert start % _lib.OPT_CHUNK == 0 and start < numels[t]
        covered[t] += min(_lib.OPT_CHUNK, numels[t] - start)
    assert [covered[i] for i in range(len(numels))] == numels


def test_host_index_arithmetic_properties():
    """property tests of the integer host logic: rank shards tile the item range exactly; the crop-start bound of the
    loaders equals the oracle's (wsj0_2mix.py:118-128 tiling rule) for any length / hop / frame_length"""
    from hypothesis import given, settings, strategies as st
    from onssen_b200.data.feature_utils import num_crop_starts
    from onssen_b200.utils.dist import shard_range
    from oracle import onssen_oracle as O

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 5000), st.integers(1, 16))
    def shards(n, world):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1 and all(s >= 0 for s in sizes)

    @settings(max_examples=200, deadline=None)
    @given(st.integers(200, 200000), st.sampled_from([16, 64, 128, 256, 441]), st.integers(1, 600))
    def crops(nsample, hop, T):
        n = num_crop_starts(nsample, hop, T)
        assert n == O.num_crop_starts(nsample, hop, T) and n >= 1
        frames = 1 + nsample // hop
        tiled = frames if frames > T else frames * (T // frames + 1)
        assert n == tiled - T                       # np.random.randint(n) -> start in [0, n-1], crop always fits

    shards()
    crops()


def test_batch_si_sdr_matches_reference_fixture():
    """tests/golden/sdr.npz = the live batch_SDR_torch(est, ref, return_perm=True) (onssen/evaluate/sdr.py:40-87)."""
    import numpy as np
    import torch
    from conftest import load_golden
    from onssen_b200.utils.test import batch_si_sdr
    _, g = load_golden("sdr.npz")
    sdr, perm = batch_si_sdr(torch.from_numpy(g["est"]), torch.from_numpy(g["ref"]), return_perm=True)
    assert sdr.dtype == torch.float32 and sdr.shape == (5,)
    np.testing.assert_allclose(sdr.numpy(), g["sdr"], atol=2e-4)
    np.testing.assert_array_equal(perm.numpy(), g["perm"])
    # three sources: permutation search against brute force in float64
    rng = np.random.RandomState(0)
    ref = torch.from_numpy(rng.standard_normal((2, 3, 500)))
    est = ref[:, [2, 0, 1]] + 0.1 * torch.from_numpy(rng.standard_normal((2, 3, 500)))
    sdr3, p3 = batch_si_sdr(est, ref, return_perm=True)
    assert sdr3.dtype == torch.float64 and (sdr3 > 15).all() and (p3 == p3[0]).all()


def test_wavio_headers_and_resample_filter(tmp_path):
    """data/wavio.py on the host: RIFF walking for the sample formats the loaders accept, the raw-PCM staging of a batch,
    and the polyphase filter bookkeeping -- evaluating y[j] = sum_i x[i] h[(j + pre) down - i up] (the formula the CUDA
    kernel implements) must reproduce scipy.signal.resample_poly."""
    import numpy as np
    from scipy.io import wavfile
    from scipy.signal import resample_poly
    from onssen_b200.data import wavio
    rng = np.random.RandomState(0)
    a16 = (rng.standard_normal(1000) * 5000).astype(np.int16)
    st16 = (rng.standard_normal((700, 2)) * 5000).astype(np.int16)
    a32 = (rng.standard_normal(500) * 1e8).astype(np.int32)
    f32 = rng.standard_normal(300).astype(np.float32)
    for name, rate, arr in (("a.wav", 8000, a16), ("b.wav", 16000, st16), ("c.wav", 44100, a32), ("d.wav", 8000, f32)):
        wavfile.write(str(tmp_path / name), rate, arr)
        r, ch, code, got = wavio.read_pcm(str(tmp_path / name))
        assert r == rate and ch == (arr.shape[1] if arr.ndim > 1 else 1) and got.dtype == arr.dtype
        np.testing.assert_array_equal(got.reshape(arr.shape), arr)
    assert np.allclose(wavio.to_float_mono(st16, "i2"), (st16.astype(np.float32) / 32768).mean(1))
    # staging: (a, a) and (short, a) pairs -> one int16 block, per-utterance length = the shorter file
    wavfile.write(str(tmp_path / "e.wav"), 8000, a16[:600])
    names = [[(str(tmp_path / "a.wav"), str(tmp_path / "a.wav")), (str(tmp_path / "e.wav"), str(tmp_path / "a.wav"))]]
    item = next(iter(wavio.PcmStager(names, workers=2, depth=1)))
    assert item["pcm"].shape == (2, 2, 1000) and item["lengths"].tolist() == [1000, 600] and item["rate"] == 8000
    assert torch.equal(item["pcm"][1, 1, :600], torch.from_numpy(a16[:600])) and int(item["pcm"][0, 1, 600:].abs().max()) == 0
    with pytest.raises(ValueError, match="differ in sampling rate"):
        next(iter(wavio.PcmStager([[(str(tmp_path / "a.wav"), str(tmp_path / "b.wav"))]], workers=1)))
    # polyphase bookkeeping vs scipy for three rate pairs
    x = rng.standard_normal(997).astype(np.float32)
    for rin, rout in ((16000, 8000), (44100, 16000), (8000, 16000)):
        up, down, h, pre = wavio.resample_filter(rin, rout)
        want = resample_poly(x.astype(np.float64), up, down)
        n_out = -(-len(x) * up // down)
        assert len(want) == n_out
        y = np.zeros(n_out)
        for j in range(n_out):
            t = (j + pre) * down
            i_hi = min(len(x) - 1, t // up)
            i_lo = max(0, -(-(t - len(h) + 1) // up))
            idx = np.arange(i_lo, i_hi + 1)
            y[j] = np.dot(x[idx].astype(np.float64), h[t - idx * up].astype(np.float64))
        assert np.abs(y - want).max() < 1e-6 * np.abs(want).max(), (rin, rout)
