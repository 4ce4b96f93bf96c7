#!/usr/bin/env python
"""bench.py -- utterances/sec (4 s @ 8 kHz, 2-speaker) for forward + loss of the deep-clustering path.

Headline workload (BASELINE.json configs[1]): deep_clustering 3x600 BLSTM, 129-bin STFT (n_fft 256 / hop 64),
40-D embedding, batch 32 per GPU, T=400 frames, loss_dc.  One step = STFT featurizer (mix, s1, s2 ->
log-magnitude + ideal-binary/VAD labels) -> BLSTM stack -> BatchNorm -> embedding head -> affinity loss on
one batch of synthetic mixtures (oracle.synth_utterance, seeds 1234+i).  Utterances shard across ranks with
no data-path collective (weak scaling: 32 utterances per GPU).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                   # the reference's own torch modules on the host cores

Extra keys of the GPU arm's JSON line (beyond the driver contract):
  train          secondary figure: featurizer + fwd + loss + hand-written backward (+ all-reduce) + clip + Adam
  gpu_reference  the UNMODIFIED reference modules (oracle/_ref) `.to('cuda')` -- cuDNN LSTM / cuBLAS / ATen -- on the
                 same device, same tensors, same warm-up / steps / CUDA events: the same-B200 bar (N=1 only)
  configs        fwd+loss and train-step ms of BASELINE configs[2..4] (cfg3 at B=64; the per-GPU shards of cfg4 / cfg5)
  e2e_disk       training throughput fed from wav files on disk through the data plugin (wall clock, N=1)
  ddp_check      (N>1) parameter-checksum spread across ranks after the train steps, and the N-rank averaged
                 gradient against the 1-rank gradient of the concatenated batch
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "utterances/sec (4s@8kHz, 2-spk) fwd+loss"
T_FRAMES = 400
DB = 40.0

# BASELINE.json configs[1..4].  B is the per-GPU batch: cfg4 = 128 / 8 GPUs, cfg5 = 256 / 8 GPUs.
WORKLOADS = {
    "cfg2": dict(desc="deep_clustering 3x600 BLSTM, 129-bin STFT (n_fft 256, hop 64), 40-D embed",
                 model="deep_clustering", margs=(129, 600, 3, 40), feat="dc", loss="loss_dc",
                 n_fft=256, hop=64, nsample=32000, B=32),
    "cfg3": dict(desc="chimera++ 4x600 BLSTM, mask+embed heads, W_MR / PSA loss, 129 bins",
                 model="chimera", margs=(129, 600, 4, 20), feat="chimera++", loss="loss_chimera_psa",
                 n_fft=256, hop=64, nsample=32000, B=64),
    "cfg4": dict(desc="phase_net (repaired) 257-bin 16 kHz (n_fft 512, hop 128), 3x300 (egs phase-net config), "
                      "batch 128 over 8 GPUs = 16/GPU",
                 model="phase_net", margs=(257, 300, 3, 20), feat="phase", loss="loss_phase",
                 n_fft=512, hop=128, nsample=64000, B=16),
    "cfg5": dict(desc="enhance + restoration layers, Edinburgh-TTS shape: 513-bin 16 kHz (n_fft 1024, hop 256), 3x600, "
                      "batch 256 over 8 GPUs = 32/GPU",
                 model="enhance", margs=(513, 600, 3), feat="edinburgh", loss="loss_mask_msa",
                 n_fft=1024, hop=256, nsample=64000, B=32),
}
CFG = WORKLOADS["cfg2"]


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def rec_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the recurrent kernel from the committed
    `ncu --set full` capture of this same command (profiles/r0N_rec_traffic.json, newest round); None if absent."""
    for name in ("r02_rec_traffic.json", "r01_rec_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["dram_bytes_per_launch"]
        except Exception:
            continue
    return None


class ClockSampler:
    """nvidia-smi clock/throttle sampling DURING the timed regions (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "covers": "all timed regions of this run (value, e2e, train)"}


def synth_batch(first_index, B, wl=CFG):
    from oracle import onssen_oracle as O
    utts = [O.synth_utterance(first_index + i, wl["nsample"]) for i in range(B)]
    hi = O.num_crop_starts(wl["nsample"], wl["hop"], T_FRAMES)
    starts = np.array([np.random.RandomState(1234 + first_index + i).randint(hi) for i in range(B)], dtype=np.int32)
    return [np.stack([u[k] for u in utts]) for k in range(3)], starts


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_featurize(waves, starts, wl=CFG):
    """oracle restatement of the librosa featurizer (librosa is not installable: parity unpinned at that boundary)."""
    from oracle import onssen_oracle as O
    feats, ohs, mags = [], [], []
    for b in range(waves[0].shape[0]):
        inp, lab = O.featurize(waves[0][b], waves[1][b], waves[2][b], wl["n_fft"], wl["hop"], T_FRAMES, int(starts[b]),
                               DB, "dc")
        feats.append(inp[0]); ohs.append(lab[0]); mags.append(lab[1])
    return np.stack(feats), np.stack(ohs), np.stack(mags)


class CpuReference:
    """The reference's own modules (oracle/_ref/onssen: nn.deep_clustering + loss.loss_dc, unmodified) on the host
    cores, same configuration as the GPU arm: B=32, train-mode BatchNorm, dropout 0.3, fp32, fwd + loss under
    no_grad.  Falls back to the numpy oracle port (kind 'port', eval-mode, B=8) only if oracle/_ref is absent."""

    def __init__(self):
        import torch
        from oracle import ref_loader
        self.torch = torch
        self.cores = os.cpu_count()
        self.kind = "reference" if ref_loader.available() else "port"
        if self.kind == "reference":
            torch.set_num_threads(self.cores)
            R = ref_loader.import_reference()
            torch.manual_seed(0)
            self.model = R.nn.deep_clustering(*CFG["margs"]).train()
            self.loss = R.loss.loss_dc
            self.B = CFG["B"]
        else:
            from oracle import onssen_oracle as O
            F, H, L, D = CFG["margs"]
            rng = np.random.RandomState(0)
            p = O.init_params_like_torch(rng, F, H, L, {"fc_dc": (F * D, 2 * H)})
            p["bn.weight"] = np.ones(2 * H, np.float32); p["bn.bias"] = np.zeros(2 * H, np.float32)
            p["bn.running_mean"] = np.zeros(2 * H, np.float32); p["bn.running_var"] = np.ones(2 * H, np.float32)
            self.params, self.B = p, 8
        self.waves, self.starts = synth_batch(0, self.B)

    def step(self):
        feat, oh, mag = cpu_featurize(self.waves, self.starts)
        if self.kind == "reference":
            t = self.torch
            with t.no_grad():
                emb, = self.model([t.from_numpy(feat)])
                return float(t.mean(self.loss([emb], [t.from_numpy(oh), t.from_numpy(mag)])))
        from oracle import onssen_oracle as O
        emb, = O.deep_clustering_forward(self.params, [feat], CFG["margs"][2], training=False)
        return float(O.loss_dc([emb], [oh, mag]).mean())

    def time(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        dt = time.perf_counter() - t0
        return self.B * steps / dt, dt / steps

    def describe(self, steps, per_step):
        what = ("the reference's own torch modules (oracle/_ref: onssen.nn.deep_clustering + onssen.loss.loss_dc, "
                "train-mode BN, dropout 0.3, fp32, no_grad)" if self.kind == "reference"
                else "numpy oracle port (eval-mode BN), oracle/_ref absent")
        return (f"{self.B} utterances x {steps} steps: oracle STFT featurizer (librosa absent) + {what}, "
                f"torch threads = {self.cores}, {per_step:.2f} s/step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    val, per_step = ref.time(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "utterances/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(ref.B, 1), "same_config": ref.kind == "reference",
            "cpu_baseline": {"value": val, "unit": "utterances/s", "cores": ref.cores, "kind": ref.kind,
                             "sample": ref.describe(args.steps, per_step)},
            "e2e": {"value": val, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(batch_per_gpu, n):
    return {"workload": f"{CFG['desc']}, T=400, batch {batch_per_gpu}/GPU, STFT featurizer + forward + loss_dc "
                        "(BASELINE configs[1])",
            "global_batch": batch_per_gpu * n, "frames": T_FRAMES, "parallelism": f"batch-shard x{n}, no data-path collective",
            "bn_mode": "train (batch statistics)", "dropout": 0.3,
            "l2_policy": "per-step working set ~1.3 GB >> 126 MB L2; 4 distinct input batches rotated"}


# ------------------------------------------------------------------------------------------------ GPU arm
def build_workload(ob, wl, dev, dropout=None):
    """-> (model, featurize(ws, st) -> (inp, lab), loss_fn) for one WORKLOADS entry, through the plugin API."""
    kw = {} if dropout is None else {"dropout": dropout}
    model = getattr(ob.nn, wl["model"])(*wl["margs"], **kw).to(dev).train()
    loss_fn = getattr(ob.loss, wl["loss"])
    feat = wl["feat"]

    def featurize(ws, st):
        name = "chimera++" if feat == "edinburgh" else feat
        inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], name, wl["n_fft"], wl["hop"], T_FRAMES, DB, crop_start=st)
        if feat == "edinburgh":          # edinburgh_tts.py:84-97 layout: [feature, mag_noisy] / [mag_clean, cos_diff]
            return [inp[0], lab[1]], [lab[2], lab[4]]
        return inp, lab

    return model, featurize, loss_fn


def _scalar(x):
    return float(x.detach()) if hasattr(x, "detach") else float(x)


def event_ms(torch, fn, n, barrier):
    """-> (total ms of n calls by CUDA events, last result, median ms of the individual calls).  The total is what the
    figures are computed from; the median exposes a run in which one call stalled on the host (a collector pass of the
    interpreter, an allocator growth) -- the garbage of the previous workload is collected before the clock starts."""
    import gc
    gc.collect()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    out = None
    for i in range(n):
        out = fn(i)
        ev[i + 1].record()
    barrier()
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ev[0].elapsed_time(ev[n]), out, per[n // 2]


def time_workload(torch, ob, wl, dev, rank, world, barrier, K, W, grad_sync=None):
    """fwd+loss and train-step ms/step of one workload on device-resident synthetic batches (2 rotated batches)."""
    B = wl["B"]
    model, featurize, loss_fn = build_workload(ob, wl, dev)
    if grad_sync is not None:
        grad_sync(model)
    batches = []
    for i in range(2):
        waves, starts = synth_batch((rank * 2 + i) * B + 5000, B, wl)
        batches.append(([torch.from_numpy(w).to(dev) for w in waves], torch.from_numpy(starts).to(dev)))

    def fwd(i):
        inp, lab = featurize(*batches[i % 2])
        return torch.mean(loss_fn(model(inp), lab))

    opt = ob.utils.build_optimizer(model.parameters(), {"name": "adam", "lr": 1e-3})

    def train(i):
        loss = fwd(i)
        opt.zero_grad()
        loss.backward()
        ob.utils.clip_grad_norm_(model.parameters(), 5)
        opt.step()
        return loss

    with torch.no_grad():
        for i in range(W):
            fwd(i)
        ms_f, lf, med_f = event_ms(torch, fwd, K, barrier)
    for i in range(min(W, 3)):
        train(i)
    ms_t, lt, med_t = event_ms(torch, train, K, barrier)
    t = torch.tensor([ms_f, ms_t], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {"workload": wl["desc"], "batch_per_gpu": B, "steps": K,
           "fwd_loss_ms": t[0].item() / K, "fwd_loss_utt_s": world * B * K / (t[0].item() / 1e3),
           "train_ms": t[1].item() / K, "train_utt_s": world * B * K / (t[1].item() / 1e3),
           "fwd_loss_ms_median_step": med_f, "train_ms_median_step": med_t,
           "loss_fwd": _scalar(lf), "loss_train": _scalar(lt)}
    del model, opt, batches
    torch.cuda.empty_cache()
    return out


def time_gpu_reference(torch, dev, ob, dev_batches, K, W):
    """The UNMODIFIED reference modules on this GPU (cuDNN LSTM / cuBLAS / ATen through torch): fwd+loss under
    no_grad and the reference's own training step (onssen/utils/train.py:75-84, including its per-step .item()).
    Inputs are the tensors this repo's device featurizer produced (the reference's featurizer is host librosa and is
    NOT included on the reference side, which favours the reference)."""
    from oracle import ref_loader
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref not staged"}
    R = ref_loader.import_reference()
    torch.manual_seed(0)
    model = R.nn.deep_clustering(*CFG["margs"]).to(dev).train()
    loss_fn = R.loss.loss_dc
    feats = []
    with torch.no_grad():
        for ws, st in dev_batches:
            feats.append(ob.data.featurize_batch(ws[0], ws[1], ws[2], "dc", CFG["n_fft"], CFG["hop"], T_FRAMES, DB,
                                                 crop_start=st))
    nb = len(feats)
    sync = torch.cuda.synchronize

    def fwd(i):
        inp, lab = feats[i % nb]
        return torch.mean(loss_fn(model(inp), lab))

    opt = torch.optim.Adam(model.parameters(), lr=1e-3)          # onssen/utils/basic.py:6-7

    def train(i):
        inp, lab = feats[i % nb]
        output = model(inp)
        loss_avg = torch.mean(loss_fn(output, lab))
        v = loss_avg.item()                                       # train.py:80
        opt.zero_grad()
        loss_avg.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5)
        opt.step()
        return v

    with torch.no_grad():
        for i in range(W):
            fwd(i)
        ms_f, lf, _ = event_ms(torch, fwd, K, sync)
    for i in range(min(W, 3)):
        train(i)
    ms_t, lt, _ = event_ms(torch, train, K, sync)
    B = CFG["B"]
    out = {"what": "reference onssen.nn.deep_clustering + onssen.loss.loss_dc (oracle/_ref, unmodified) .to('cuda'): "
                   "cuDNN LSTM / cuBLAS / ATen, fp32, train-mode BN, dropout 0.3; model + loss only (no featurizer)",
           "fwd_loss_ms": ms_f / K, "fwd_loss_utt_s": B * K / (ms_f / 1e3),
           "train_ms": ms_t / K, "train_utt_s": B * K / (ms_t / 1e3), "steps": K,
           "allow_tf32": {"cudnn": bool(torch.backends.cudnn.allow_tf32),
                          "matmul": bool(torch.backends.cuda.matmul.allow_tf32)},
           "loss_fwd": _scalar(lf), "loss_train": _scalar(lt)}
    del model, opt, feats
    torch.cuda.empty_cache()
    return out


def time_disk_training(torch, ob, dev, model, opt, K):
    """Training steps fed FROM DISK through the data plugin: a temporary wsj0-2mix-layout corpus of
    synthetic 16-bit wavs (128 utterances x {mix,s1,s2}, each read 25x per epoch = 100 steps), `wsj0_2mix_dataloader` (thread-pool PCM staging two
    batches ahead, decode + featurizer on the device), forward + loss + backward + clip + Adam; WALL CLOCK per step
    over whole epochs, first epoch untimed."""
    import shutil
    import tempfile
    from scipy.io import wavfile
    B, nb = CFG["B"], 4
    root = tempfile.mkdtemp(prefix="onssen_b200_bench_")
    try:
        from oracle import onssen_oracle as O
        for sub in ("mix", "s1", "s2"):
            os.makedirs(os.path.join(root, "wav8k", "min", "tr", sub))
        q = lambda x: np.clip(np.round(x * 32768), -32768, 32767).astype(np.int16)
        for i in range(B * nb):
            for sub, x in zip(("mix", "s1", "s2"), O.synth_utterance(7000 + i, CFG["nsample"])):
                wavfile.write(os.path.join(root, "wav8k", "min", "tr", sub, f"u{i:04d}.wav"), 8000, q(x))
        fo = dict(data_path=root, batch_size=B, frame_length=T_FRAMES, sampling_rate=8000, window_size=CFG["n_fft"],
                  hop_size=CFG["hop"], db_threshold=DB)
        loader = ob.data.wsj0_2mix_dataloader("dc", fo, "tr", dev)
        loader.file_list = loader.file_list * 25         # 100 steps per epoch (wsj0-2mix tr has 625 at this batch size):
                                                         # the start-up of the staging thread is paid once per epoch

        def epoch():
            n = 0
            for inp, lab in loader:
                loss = torch.mean(ob.loss.loss_dc(model(inp), lab))
                opt.zero_grad()
                loss.backward()
                ob.utils.clip_grad_norm_(model.parameters(), 5)
                opt.step()
                n += 1
            return n

        epoch()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 0
        while steps < K:
            steps += epoch()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": B * steps / dt, "unit": "utterances/s", "ms_per_step": dt / steps * 1e3, "steps": steps,
                "files_per_step": 3 * B, "timing": "wall clock (host decode threads + device), one GPU",
                "loader": "wsj0_2mix_dataloader: 8 decode threads, 2 batches staged ahead, int16 -> float + STFT on device"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def ddp_check(torch, ob, dev, rank, world, model_after_training):
    """(i) spread of a parameter checksum across ranks after the timed train steps; (ii) relative error of the
    world-averaged gradient (SyncBN + global loss coupling on) against the 1-rank gradient of the concatenated batch,
    computed on every rank from the same seeded data."""
    import torch.distributed as dist
    from onssen_b200.utils.ddp import GradSync, broadcast_parameters, enable_sync_batchnorm
    cs = torch.stack([p.detach().double().abs().sum() for p in model_after_training.parameters()]).sum().reshape(1)
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo) / hi.abs().clamp_min(1e-30))
    # small-model exactness check (H=64, 2 layers) so that the single-rank full batch is cheap
    wl = dict(WORKLOADS["cfg2"], margs=(129, 64, 2, 20), B=4)
    Bl = wl["B"]
    waves, starts = synth_batch(9000, Bl * world, wl)
    cu = lambda a: torch.from_numpy(a).to(dev)
    torch.manual_seed(5)
    full, featurize, loss_fn = build_workload(ob, wl, dev, dropout=0.0)
    inp, lab = featurize([cu(w) for w in waves], cu(starts))
    torch.mean(loss_fn(full(inp), lab)).backward()
    g_full = torch.cat([p.grad.flatten() for p in full.parameters()]).double()
    torch.manual_seed(5)
    part, featurize, loss_fn = build_workload(ob, wl, dev, dropout=0.0)
    broadcast_parameters(part)
    ob.utils.ddp.enable_global_loss_mean(True)
    part.grad_sync = GradSync()
    enable_sync_batchnorm(part)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    inp, lab = featurize([cu(w[sl]) for w in waves], cu(starts[sl]))
    torch.mean(loss_fn(part(inp), lab)).backward()           # GradSync buckets are waited inside the backward
    g_part = torch.cat([p.grad.flatten() for p in part.parameters()]).double()
    rel = float((g_part - g_full).norm() / g_full.norm())
    return {"param_checksum_spread": spread, "grad_rel_err_vs_single_rank_full_batch": rel,
            "what": f"deep_clustering(129,64,2,20), {Bl} utterances/rank, SyncBN + global (B,B) loss mean"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import onssen_b200 as ob
    from onssen_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T = CFG["B"], T_FRAMES
    torch.manual_seed(0)
    model = ob.nn.deep_clustering(*CFG["margs"]).to(dev).train()
    nbatch = 4
    host_batches = []
    for i in range(nbatch):
        waves, starts = synth_batch((rank * nbatch + i) * B, B)
        host_batches.append(([torch.from_numpy(w).pin_memory() for w in waves], torch.from_numpy(starts).pin_memory()))
    dev_batches = [([w.to(dev) for w in ws], st.to(dev)) for ws, st in host_batches]

    def step(ws, st):
        inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], "dc", CFG["n_fft"], CFG["hop"], T, DB, crop_start=st)
        emb, = model(inp)
        return ob.loss.loss_dc([emb], lab).mean()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = args.warmup, args.steps
    sampler = ClockSampler(local)
    with torch.no_grad():
        # ---------------- device-resident timing (value) ----------------
        for i in range(W):
            step(*dev_batches[i % nbatch])
        barrier()
        if rank == 0:
            sampler.start()
        _lib.LAUNCHES[0] = 0
        _lib.REC_EVENTS = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            loss = step(*dev_batches[i % nbatch])
        e1.record()
        barrier()
        launches = _lib.LAUNCHES[0]
        rec_ms = [a.elapsed_time(b) for a, b in _lib.REC_EVENTS]
        _lib.REC_EVENTS = None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        # ---------------- end-to-end timing (host buffers in, loss out) ----------------
        # Every step's inputs are copied from pinned host memory inside the timed region and its loss is read back
        # (`.item()`); step i+1's copy is enqueued on a copy stream before step i's kernels, so it overlaps them
        # (double buffering, what the loaders' staging does) -- K uploads for K steps, none hoisted out.
        copy_stream = torch.cuda.Stream(device=dev)

        def upload(i):
            ws, st = host_batches[i % nbatch]
            with torch.cuda.stream(copy_stream):
                out = ([w.to(dev, non_blocking=True) for w in ws], st.to(dev, non_blocking=True))
                ev = copy_stream.record_event()
            return out, ev

        loss_host = torch.empty(2, dtype=torch.float32).pin_memory()

        def e2e_steps(n):
            nxt = upload(0)
            lv, prev = None, None
            for i in range(n):
                (ws_d, st_d), ev = nxt
                if i + 1 < n:
                    nxt = upload(i + 1)
                torch.cuda.current_stream().wait_event(ev)
                loss_host[i & 1].copy_(step(ws_d, st_d), non_blocking=True)     # D2H read of the step's result ...
                done = torch.cuda.current_stream().record_event()
                if prev is not None:                     # ... consumed on the host one step later, so that the
                    prev[0].synchronize()                # next step's launches are already queued behind it
                    lv = float(loss_host[prev[1]])
                prev = (done, i & 1)
            prev[0].synchronize()
            return float(loss_host[prev[1]])

        e2e_steps(min(W, 2))
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        lv = e2e_steps(K)
        f1.record()
        barrier()
        ms_e2e = torch.tensor([f0.elapsed_time(f1)], device=dev)
    # ---------------- training step (secondary figure): fwd + loss + backward (+ grad all-reduce) + clip + Adam
    ms_train = None
    KT = max(20, K)

    def attach_sync(m):
        from onssen_b200.utils.ddp import GradSync, broadcast_parameters
        if world > 1:
            broadcast_parameters(m)
            ob.utils.ddp.enable_global_loss_mean(True)
            m.grad_sync = GradSync()
            if args.sync_bn:
                ob.utils.ddp.enable_sync_batchnorm(m)

    if not args.no_train:
        attach_sync(model)
        opt = ob.utils.build_optimizer(model.parameters(), {"name": "adam", "lr": 1e-3})

        def train_step(ws, st):
            inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], "dc", CFG["n_fft"], CFG["hop"], T, DB, crop_start=st)
            loss = torch.mean(ob.loss.loss_dc(model(inp), lab))
            opt.zero_grad()
            loss.backward()
            ob.utils.clip_grad_norm_(model.parameters(), 5)
            opt.step()
            return loss

        for i in range(3):
            train_step(*dev_batches[i % nbatch])
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(KT):
            tl = train_step(*dev_batches[i % nbatch])
        g1.record()
        barrier()
        ms_train = torch.tensor([g0.elapsed_time(g1)], device=dev)
        train_loss = float(tl.item())
    clocks = sampler.stop() if rank == 0 else None
    disk = None
    if world == 1 and not args.no_train and not args.no_disk:
        try:
            disk = time_disk_training(torch, ob, dev, model, opt, KT)
        except Exception as exc:
            disk = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        if ms_train is not None:
            dist.all_reduce(ms_train, op=dist.ReduceOp.MAX)
    check = None
    if world > 1 and not args.no_train:
        check = ddp_check(torch, ob, dev, rank, world, model)
    # ---------------- the other BASELINE configs (extra keys; not the headline) ----------------
    configs = None
    if not args.no_configs:
        del model
        torch.cuda.empty_cache()
        KC = max(5, min(K, 10))
        configs = {name: time_workload(torch, ob, WORKLOADS[name], dev, rank, world, barrier, KC, 3, attach_sync)
                   for name in ("cfg3", "cfg4", "cfg5")}
    gpu_ref = None
    if world == 1 and not args.no_gpu_reference:
        try:
            gpu_ref = time_gpu_reference(torch, dev, ob, dev_batches, K, W)
        except Exception as exc:                        # the bar is a report, never a reason to lose the line
            gpu_ref = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    if rank == 0:
        pk, pk_src = peaks()
        total = ms.item() / 1e3
        value = world * B * K / total
        e2e_val = world * B * K / (ms_e2e.item() / 1e3)
        # roofline of the dominant kernel (persistent BLSTM recurrence), SURVEY.md section 8(d):
        # per layer-step both directions stream W_hh (fp16) once: 2*4H*H*2 bytes, plus the gate
        # pre-activations read and the layer output written once per launch.
        _, H, L, _ = CFG["margs"]
        rec_avg_ms = float(np.mean(rec_ms)) if rec_ms else None
        bytes_stream = T * (2 * 4 * H * H * 2)
        bytes_io = T * B * 8 * H * 4 + T * B * 2 * H * 2
        alg_bytes = bytes_stream + bytes_io
        flops = T * 2 * 2 * B * 4 * H * H
        roof = None
        if rec_avg_ms:
            ach = alg_bytes / (rec_avg_ms * 1e-3) / 1e9
            roof = {"kernel": "blstm_rec_kernel (one launch = one BLSTM layer, both directions, T steps)",
                    "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                    "traffic": rec_traffic(), "peak_source": pk_src + " (burst copy bandwidth, kernel timed per launch)",
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "avg_launch_ms": rec_avg_ms, "us_per_step": rec_avg_ms * 1e3 / T,
                    "share_of_step": float(np.sum(rec_ms)) / ms.item(),
                    "tensor_alt": {"flops_per_launch": flops, "achieved_tflops": flops / (rec_avg_ms * 1e-3) / 1e12,
                                   "peak_tflops": pk.get("bf16_tflops_sustained")},
                    "note": "weights are TMEM-resident, so real DRAM traffic is only the gate/y streams; the "
                            "weight-stream model is the SURVEY 8(d) bound a non-persistent kernel would hit"}
        cpu = CpuReference()
        cpu_val, cpu_step = cpu.time(2, 1)
        line = {"metric": METRIC, "value": value, "unit": "utterances/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms.item() / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate (tcgen05 kind::f16), f32 elsewhere", "data": "synthetic",
                "config": workload_config(B, world),
                "e2e": {"value": e2e_val, "unit": "utterances/s",
                        "h2d_bytes_per_step": 3 * B * CFG["nsample"] * 4 + B * 4,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e.item() / K,
                        "pipeline": "pinned host -> device copy of step i+1 enqueued on a copy stream before step i's "
                                    "kernels (double buffering); every step's loss is copied to pinned host memory "
                                    "and read there while the next step runs; all K copies and K read-backs are "
                                    "inside the timed region"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof,
                "cpu_baseline": {"value": cpu_val, "unit": "utterances/s", "cores": cpu.cores, "kind": cpu.kind,
                                 "sample": cpu.describe(2, cpu_step)},
                "loss_mean_last": lv}
        if ms_train is not None:
            line["train"] = {"value": world * B * KT / (ms_train.item() / 1e3), "unit": "utterances/s",
                             "ms_per_step": ms_train.item() / KT, "steps": KT, "loss_last": train_loss,
                             "sync_bn": bool(args.sync_bn and world > 1),
                             "includes": "featurizer + forward + loss_dc + hand-written backward (BPTT) + "
                                         "bucketed NCCL gradient all-reduce (N>1) + clip_grad_norm_(5) + Adam "
                                         "(multi-tensor kernels of this repo)"}
        if disk is not None:
            line["e2e_disk"] = disk
            if "ms_per_step" in disk and ms_train is not None:
                disk["fraction_of_device_resident_train"] = (ms_train.item() / KT) / disk["ms_per_step"]
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
            if "fwd_loss_ms" in gpu_ref:
                line["gpu_reference"]["speedup_fwd_loss"] = gpu_ref["fwd_loss_ms"] / (ms.item() / K)
                if ms_train is not None:
                    line["gpu_reference"]["speedup_train"] = gpu_ref["train_ms"] / (ms_train.item() / KT)
        if configs is not None:
            line["configs"] = configs
        if check is not None:
            line["ddp_check"] = check
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step figure")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3/cfg4/cfg5 block")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the same-GPU reference (cuDNN) leg")
    ap.add_argument("--no-disk", action="store_true", help="skip the from-disk training figure (e2e_disk)")
    ap.add_argument("--sync-bn", action="store_true",
                    help="training figure with whole-batch BatchNorm statistics across ranks (N>1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
