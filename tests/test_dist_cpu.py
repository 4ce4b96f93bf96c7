"""world_size-2 gloo tests (CPU) of the batch-sharding host logic used by bench.py --gpus N / the loaders."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from onssen_b200.utils import dist as D
    assert D.env_rank_world() == (rank, world, rank)
    # utterance shards: disjoint, contiguous, cover everything
    n = 37
    lo, hi = D.shard_range(n, rank, world)
    owned = torch.zeros(n)
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    assert torch.equal(owned, torch.ones(n))
    # timing protocol: every rank sees the slowest rank's time
    t = D.max_over_ranks(1.0 + rank)
    assert t == float(world)
    # global loss mean from per-shard vectors == single-process value of mean over the (B,B) outer product
    rng = np.random.RandomState(0)
    l_all = torch.from_numpy(rng.uniform(0.1, 1, n).astype(np.float32))
    m_all = torch.from_numpy(rng.uniform(10, 20, n).astype(np.float32))
    want = float((l_all[None, :] * m_all[:, None]).double().mean())
    got = D.global_loss_mean(l_all[lo:hi], m_all[lo:hi])
    assert abs(got - want) < 1e-9 * abs(want) + 1e-9, (got, want)
    # and it differs from the naive mean of per-shard means (why the 4-scalar reduction exists)
    naive = D.sum_over_ranks(float((l_all[lo:hi][None, :] * m_all[lo:hi][:, None]).double().mean())) / world
    # bucketed gradient all-reduce (gloo here, NCCL on the GPU box): rank average, shared bias tensor reduced once
    from onssen_b200.utils.ddp import GradSync, broadcast_parameters
    sync = GradSync()
    gb = torch.full((5,), float(rank + 1))
    g1 = {"fc.weight": torch.full((3, 4), float(rank + 1)), "fc.bias": torch.arange(4.0) * (rank + 1)}
    g2 = {"rnn.weight_ih_l0": torch.ones(6, 2) * (10 * (rank + 1)), "rnn.bias_ih_l0": gb}
    sync.reduce_bucket(g1)
    sync.reduce_bucket(g2)
    sync.wait()
    avg = sum(range(1, world + 1)) / world
    assert torch.allclose(g1["fc.weight"], torch.full((3, 4), avg)) and torch.allclose(g1["fc.bias"], torch.arange(4.0) * avg)
    assert torch.allclose(g2["rnn.weight_ih_l0"], torch.ones(6, 2) * 10 * avg) and torch.allclose(gb, torch.full((5,), avg))
    assert sync.bytes_reduced == (12 + 4 + 12 + 5) * 4
    # deferred mode (the NCCL default): buckets are held back, flush() launches ONE collective over all of them (the
    # backward calls it once its last cooperative kernel is enqueued), wait() finishes it
    held = GradSync(overlap=False)
    h1 = {"fc.weight": torch.full((3, 4), float(rank + 1))}
    h2 = {"rnn.weight_hh_l0": torch.full((2, 2), 3.0 * (rank + 1)), "rnn.bias_ih_l0": torch.arange(3.0) * (rank + 1)}
    held.reduce_bucket(h1)
    held.reduce_bucket(h2)
    assert not held.pending and len(held.deferred) == 2
    assert float(h1["fc.weight"][0, 0]) == float(rank + 1)            # nothing reduced yet
    held.flush()
    assert len(held.pending) == 1 and not held.deferred
    held.flush()                                                      # idempotent
    assert len(held.pending) == 1
    held.wait()
    assert torch.allclose(h1["fc.weight"], torch.full((3, 4), avg))
    assert torch.allclose(h2["rnn.weight_hh_l0"], torch.full((2, 2), 3.0 * avg))
    assert torch.allclose(h2["rnn.bias_ih_l0"], torch.arange(3.0) * avg)
    assert held.bytes_reduced == (12 + 4 + 3) * 4 and not held.pending
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank))
    v0 = lin.weight._version
    broadcast_parameters(lin, 0)
    assert float(lin.weight.abs().max()) == 0.0
    assert lin.weight._version > v0                   # packed-weight caches keyed on _version see the broadcast
    # exact data-parallel coupling of the (B,B) loss mean (utils.ddp.enable_global_loss_mean): with the first factor
    # replaced by its global mean, the rank-average of mean(loss_bb) is the single-device value and the rank-averaged
    # gradient coefficient of l_j is (sum m / n) / n
    import importlib
    from onssen_b200.utils.ddp import enable_global_loss_mean
    LD = importlib.import_module("onssen_b200.loss.loss_dc")
    enable_global_loss_mean(True)
    assert LD.GLOBAL_MEAN[0]
    nloc = 6
    l_loc = l_all[rank * nloc:(rank + 1) * nloc].clone().requires_grad_(True)
    m_loc = m_all[rank * nloc:(rank + 1) * nloc]
    mbar = LD._global_first_factor(m_loc)
    want_mbar = float(m_all[:world * nloc].double().mean())
    assert abs(float(mbar) - want_mbar) < 1e-5 * want_mbar
    per_rank = ((mbar * l_loc).unsqueeze(0).expand(nloc, nloc)).mean()
    per_rank.backward()
    tot = per_rank.detach().clone().double()
    dist.all_reduce(tot)
    la, ma = l_all[:world * nloc].double(), m_all[:world * nloc].double()
    assert abs(float(tot) / world - float((la[None, :] * ma[:, None]).mean())) < 1e-6 * float(tot)
    assert torch.allclose(l_loc.grad / world, torch.full((nloc,), want_mbar / (world * nloc)), rtol=1e-5)
    LD.GLOBAL_MEAN[0] = False
    # rank-aware trainer: parameters broadcast from rank 0, validation loss averaged over ranks (identical early-stop
    # decisions), only rank 0 writes the checkpoint.  A torch-only stand-in model: the trainer is model-agnostic.
    from onssen_b200.utils import AttrDict, trainer
    model = torch.nn.Linear(3, 1)
    with torch.no_grad():
        model.weight.fill_(1.0 + rank)
        model.bias.fill_(0.0)
    batches = [([torch.full((2, 3), float(rank + 1 + i))], [torch.zeros(2, 1)]) for i in range(3)]
    ckpt = os.path.join(os.environ["ONSSEN_TEST_TMP"], "ckpt")
    args = AttrDict({"model_name": "dc", "device": "cpu", "num_epoch": 1, "checkpoint_path": ckpt, "verbose": False,
                     "model": model, "train_loader": batches, "valid_loader": batches,
                     "loss_fn": lambda out, lab: (out[0] - lab[0]).abs(),
                     "optimizer": torch.optim.SGD(model.parameters(), lr=0.1)})
    model_call = model.forward
    model.forward = lambda inp: [model_call(inp[0])]             # plugin contract: list in, list out
    tr = trainer(args)
    assert torch.equal(model.weight, torch.ones(1, 3)) and model.grad_sync is not None     # rank 0's parameters
    cv = tr.validate(0)
    mine = sum(3.0 * (rank + 1 + i) for i in range(3)) / 3
    both = sum(sum(3.0 * (r + 1 + i) for i in range(3)) / 3 for r in range(world)) / world
    assert abs(cv - both) < 1e-5 and abs(mine - both) > 0.1, (cv, mine, both)
    dist.barrier()
    assert os.path.exists(os.path.join(ckpt, "final.mdl")) and tr.rank == rank and tr.world == world
    ret[rank] = (got, naive, want)
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_reductions_world2(tmp_path):
    os.environ["ONSSEN_TEST_TMP"] = str(tmp_path)
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert len(ret) == world
    got0, naive0, want0 = ret[0]
    assert abs(got0 - ret[1][0]) < 1e-12
    assert abs(naive0 - want0) > 1e-6          # the naive per-shard mean is NOT the global value


def test_loader_file_sharding(tmp_path):
    from scipy.io import wavfile
    from onssen_b200.data import wsj0_2mix_dataloader
    root = tmp_path / "wav8k" / "min" / "tr"
    for sub in ("mix", "s1", "s2"):
        (root / sub).mkdir(parents=True)
    rng = np.random.RandomState(0)
    for i in range(5):
        n = 4000 + 100 * i
        for sub in ("mix", "s1", "s2"):
            wavfile.write(str(root / sub / f"utt{i}.wav"), 8000, (rng.standard_normal(n) * 3000).astype(np.int16))
    fo = dict(data_path=str(tmp_path), batch_size=2, frame_length=40, sampling_rate=8000, window_size=256, hop_size=64,
              db_threshold=40)
    full = wsj0_2mix_dataloader("dc", fo, "tr")
    assert len(full.file_list) == 5 and len(full) == 3
    loaders = [wsj0_2mix_dataloader("dc", fo, "tr", rank=r, world_size=2) for r in range(2)]
    parts = [ld.file_list for ld in loaders]
    # 5 files, 2 ranks x batch 2: wrap-padded to 8 so that both ranks run the SAME number of full batches (a rank with an
    # extra step would issue all-reduces nobody answers); every file is still covered
    assert set(parts[0] + parts[1]) == set(full.file_list) and len(parts[0]) == len(parts[1]) == 4
    assert len(loaders[0]) == len(loaders[1]) == 2
    from onssen_b200.data.wsj0_2mix import shard_files
    for n, world, bs in [(0, 2, 4), (1, 4, 3), (7, 2, 2), (8, 2, 2), (33, 8, 4), (5, 1, 2)]:
        files = [f"f{i}" for i in range(n)]
        sh = [shard_files(files, r, world, bs) for r in range(world)]
        assert len({len(x) for x in sh}) == 1 and (world == 1 or len(sh[0]) % bs == 0)
        assert set(sum(sh, [])) == set(files)
    (mix, s1, s2), lengths = full._load(full.file_list[:2])
    assert mix.shape == (2, 4100) and lengths.tolist() == [4000, 4100]
    assert float(mix[0, 4000:].abs().max()) == 0.0           # zero padded to the batch pitch


def test_edinburgh_loader_file_conventions(tmp_path):
    from scipy.io import wavfile
    from onssen_b200.data import edinburgh_tts_dataloader
    for sub in ("noisy_trainset_28spk_wav", "clean_trainset_28spk_wav"):
        (tmp_path / sub).mkdir()
    rng = np.random.RandomState(1)
    names = [f"p{i}.wav" for i in range(3)]
    for i, n in enumerate(names):
        clean = (rng.standard_normal(3000 + 10 * i) * 2000).astype(np.int16)
        noise = (rng.standard_normal(3000 + 10 * i) * 500).astype(np.int16)
        wavfile.write(str(tmp_path / "clean_trainset_28spk_wav" / n), 16000, clean)
        wavfile.write(str(tmp_path / "noisy_trainset_28spk_wav" / n), 16000, (clean + noise).astype(np.int16))
    (tmp_path / "train").write_text("\n".join(names) + "\n")
    fo = dict(data_path=str(tmp_path), batch_size=2, frame_length=20, sampling_rate=16000, window_size=512, hop_size=128,
              db_threshold=40)
    ld = edinburgh_tts_dataloader("chimera++", fo, "train")
    assert len(ld.file_list) == 3 and len(ld) == 2
    (mix, clean, noise), lengths = ld._load(sorted(ld.file_list)[:2])
    assert mix.shape == (2, 3010) and lengths.tolist() == [3000, 3010]
    assert torch.allclose(mix - clean, noise, atol=1e-6)       # "speaker 2" = mix - clean


def test_daps_segment_cursor_semantics(tmp_path):
    """daps_enhance.py:90-122: consecutive non-overlapping segments of the current recording; a new file is popped at
    index % len(list) when fewer than frame_length frames remain; the list is re-read when exhausted."""
    import random
    from onssen_b200.data.daps_enhance import SegmentCursor, daps_enhance_dataloader
    lengths = {"a_x_noisy.wav": 25, "b_y_noisy.wav": 10, "c_z_noisy.wav": 31}
    (tmp_path / "train").write_text("\n".join(lengths) + "\n")
    calls = []

    def featurize(path):
        calls.append(path)
        n = lengths[path]
        base = torch.arange(n, dtype=torch.float32)[:, None] + 1000 * len(calls)
        return [base, base + 0.5], [base + 0.25, base + 0.75]

    random.seed(0)
    cur = SegmentCursor(str(tmp_path / "train"), 10, featurize)
    segs = [cur.next_item(i) for i in (5, 1, 7, 2, 9, 4, 0)]
    firsts = [float(s[0][0][0, 0]) for s in segs]
    # every item is 10 frames; segments of one file are consecutive and leftovers (< 10 frames) are dropped
    assert all(s[0][0].shape == (10, 1) and s[1][1].shape == (10, 1) for s in segs)
    per_file = {}
    for f in firsts:
        per_file.setdefault(int(f // 1000), []).append(f % 1000)
    assert all(v == [10.0 * i for i in range(len(v))] for v in per_file.values())
    counts = [len(per_file[k]) for k in sorted(per_file)]
    assert counts[:-1] == [lengths[c] // 10 for c in calls][:len(counts) - 1] and sum(counts) == 7
    assert len(calls) >= 4 and set(calls[:3]) == set(lengths)      # list exhausted -> re-read (file 4 = a repeat)
    ld = daps_enhance_dataloader(3, dict(data_path=str(tmp_path), batch_size=2, frame_length=10, sampling_rate=16000,
                                         window_size=512, hop_size=128), "train")
    assert len(ld) == 3 and ld.clean_path("/x/f1_script2_iphone.wav") == str(tmp_path) + "/clean/f1_script2_clean.wav"


def test_read_wav_resamples_on_the_host(tmp_path):
    """files at another rate are brought to feature_options.sampling_rate (feature_utils.py:15-19 behaviour)"""
    from scipy.io import wavfile
    from onssen_b200.data.wsj0_2mix import _read_wav
    t = np.arange(48000) / 48000.0
    x = (0.5 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    wavfile.write(str(tmp_path / "a.wav"), 48000, (x * 32767).astype(np.int16))
    y = _read_wav(str(tmp_path / "a.wav"), 16000)
    assert y.dtype == np.float32 and y.shape == (16000,)
    ref = 0.5 * np.sin(2 * np.pi * 440 * np.arange(16000) / 16000.0)
    assert np.abs(y[200:-200] - ref[200:-200]).max() < 2e-3          # a 440 Hz tone survives 48 -> 16 kHz unchanged
    same = _read_wav(str(tmp_path / "a.wav"), 48000)
    assert same.shape == (48000,) and np.abs(same - (x * 32767).astype(np.int16) / 32768.0).max() < 1e-7
