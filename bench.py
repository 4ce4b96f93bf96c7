#!/usr/bin/env python
"""bench.py -- utterances/sec (4 s @ 8 kHz, 2-speaker) for forward + loss of the deep-clustering path.

Workload (BASELINE.json configs[1]): deep_clustering 3x600 BLSTM, 129-bin STFT (n_fft 256 / hop 64),
40-D embedding, batch 32 per GPU, T=400 frames, loss_dc.  One step = STFT featurizer (mix, s1, s2 ->
log-magnitude + ideal-binary/VAD labels) -> BLSTM stack -> BatchNorm -> embedding head -> affinity loss on
one batch of synthetic mixtures (oracle.synth_utterance, seeds 1234+i).  Utterances shard across ranks with
no data-path collective (weak scaling: 32 utterances per GPU).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                   # the reference algorithm's CPU path (oracle port)
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

CFG = dict(F=129, H=600, L=3, D=40, T=400, B=32, n_fft=256, hop=64, nsample=32000, db=40.0)
METRIC = "utterances/sec (4s@8kHz, 2-spk) fwd+loss"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def rec_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the recurrent kernel from the committed
    `ncu --set full` capture of this same command (profiles/r01_ncu_full_final.csv); None if absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_rec_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clock/throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
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
                "samples": len(sm)}


def synth_batch(first_index, B):
    from oracle import onssen_oracle as O
    utts = [O.synth_utterance(first_index + i, CFG["nsample"]) for i in range(B)]
    starts = np.array([np.random.RandomState(1234 + first_index + i).randint(101) for i in range(B)], dtype=np.int32)
    return [np.stack([u[k] for u in utts]) for k in range(3)], starts


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_fwd_loss(params, waves, starts):
    """The reference algorithm on the host: featurizer + deep_clustering forward + loss_dc (oracle port)."""
    from oracle import onssen_oracle as O
    feats, ohs, mags = [], [], []
    for b in range(waves[0].shape[0]):
        inp, lab = O.featurize(waves[0][b], waves[1][b], waves[2][b], CFG["n_fft"], CFG["hop"], CFG["T"], int(starts[b]),
                               CFG["db"], "dc")
        feats.append(inp[0]); ohs.append(lab[0]); mags.append(lab[1])
    emb, = O.deep_clustering_forward(params, [np.stack(feats)], CFG["L"], training=False)
    return float(O.loss_dc([emb], [np.stack(ohs), np.stack(mags)]).mean())


def cpu_params():
    from oracle import onssen_oracle as O
    rng = np.random.RandomState(0)
    p = O.init_params_like_torch(rng, CFG["F"], CFG["H"], CFG["L"], {"fc_dc": (CFG["F"] * CFG["D"], 2 * CFG["H"])})
    p["bn.weight"] = np.ones(2 * CFG["H"], np.float32); p["bn.bias"] = np.zeros(2 * CFG["H"], np.float32)
    p["bn.running_mean"] = np.zeros(2 * CFG["H"], np.float32); p["bn.running_var"] = np.ones(2 * CFG["H"], np.float32)
    return p


def time_cpu(sample_utts, steps, warmup):
    params = cpu_params()
    waves, starts = synth_batch(0, sample_utts)
    for _ in range(warmup):
        cpu_fwd_loss(params, waves, starts)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_fwd_loss(params, waves, starts)
    dt = time.perf_counter() - t0
    return sample_utts * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    sample = 8
    val, per_step = time_cpu(sample, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "utterances/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(sample, 1),
            "cpu_baseline": {"value": val, "unit": "utterances/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} utterances per step (featurizer + 3x600 BLSTM + head + loss_dc), "
                                       "numpy oracle port of the reference algorithm, BLAS threads = all cores"},
            "e2e": {"value": val, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(batch_per_gpu, n):
    return {"workload": "deep_clustering 3x600 BLSTM, 129-bin STFT (n_fft 256, hop 64), 40-D embed, T=400, "
                        f"batch {batch_per_gpu}/GPU, STFT featurizer + forward + loss_dc (BASELINE configs[1])",
            "global_batch": batch_per_gpu * n, "frames": CFG["T"], "parallelism": f"batch-shard x{n}, no data-path collective",
            "bn_mode": "train (batch statistics)", "dropout": 0.3,
            "l2_policy": "per-step working set ~1.3 GB >> 126 MB L2; 4 distinct input batches rotated"}


# ------------------------------------------------------------------------------------------------ GPU arm
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
    B, T = CFG["B"], CFG["T"]
    torch.manual_seed(0)
    model = ob.nn.deep_clustering(CFG["F"], CFG["H"], CFG["L"], CFG["D"]).to(dev).train()
    nbatch = 4
    host_batches = []
    for i in range(nbatch):
        waves, starts = synth_batch((rank * nbatch + i) * B, B)
        host_batches.append(([torch.from_numpy(w).pin_memory() for w in waves], torch.from_numpy(starts)))
    dev_batches = [([w.to(dev) for w in ws], st.to(dev)) for ws, st in host_batches]

    def step(ws, st):
        inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], "dc", CFG["n_fft"], CFG["hop"], T, CFG["db"],
                                           crop_start=st)
        emb, = model(inp)
        return ob.loss.loss_dc([emb], lab).mean()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = args.warmup, args.steps
    with torch.no_grad():
        # ---------------- device-resident timing (value) ----------------
        for i in range(W):
            step(*dev_batches[i % nbatch])
        sampler = ClockSampler(local)
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
        clocks = sampler.stop() if rank == 0 else None
        launches = _lib.LAUNCHES[0]
        rec_ms = [a.elapsed_time(b) for a, b in _lib.REC_EVENTS]
        _lib.REC_EVENTS = None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        # ---------------- end-to-end timing (host buffers in, loss out) ----------------
        for i in range(min(W, 2)):
            ws, st = host_batches[i % nbatch]
            step([w.to(dev, non_blocking=True) for w in ws], st.to(dev, non_blocking=True)).item()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(K):
            ws, st = host_batches[i % nbatch]             # pinned host waveforms -> H2D inside the timed region
            lv = step([w.to(dev, non_blocking=True) for w in ws], st.to(dev, non_blocking=True)).item()   # D2H read
        f1.record()
        barrier()
        ms_e2e = torch.tensor([f0.elapsed_time(f1)], device=dev)
    # ---------------- training step (secondary figure): fwd + loss + backward (+ grad all-reduce) + clip + Adam
    ms_train = None
    if not args.no_train:
        from onssen_b200.utils.ddp import GradSync, broadcast_parameters
        if world > 1:
            broadcast_parameters(model)
            model.grad_sync = GradSync()
            if args.sync_bn:
                ob.utils.ddp.enable_sync_batchnorm(model)
        opt = ob.utils.build_optimizer(model.parameters(), {"name": "adam", "lr": 1e-3})

        def train_step(ws, st):
            inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], "dc", CFG["n_fft"], CFG["hop"], T, CFG["db"],
                                               crop_start=st)
            loss = torch.mean(ob.loss.loss_dc(model(inp), lab))
            opt.zero_grad()
            loss.backward()
            ob.utils.clip_grad_norm_(model.parameters(), 5)
            opt.step()
            return loss

        KT = max(3, K // 4)
        for i in range(2):
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
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        if ms_train is not None:
            dist.all_reduce(ms_train, op=dist.ReduceOp.MAX)
    if rank == 0:
        pk, pk_src = peaks()
        total = ms.item() / 1e3
        value = world * B * K / total
        e2e_val = world * B * K / (ms_e2e.item() / 1e3)
        # roofline of the dominant kernel (persistent BLSTM recurrence), SURVEY.md section 8(d):
        # per layer-step both directions stream W_hh (fp16) once: 2*4H*H*2 bytes, plus the gate
        # pre-activations read and the layer output written once per launch.
        H, L = CFG["H"], CFG["L"]
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
        cores = os.cpu_count()
        sample = 8
        cpu_val, cpu_step = time_cpu(sample, 2, 1)
        line = {"metric": METRIC, "value": value, "unit": "utterances/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms.item() / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands / f32 accumulate (tcgen05 kind::f16), f32 elsewhere", "data": "synthetic",
                "config": workload_config(B, world),
                "e2e": {"value": e2e_val, "unit": "utterances/s", "h2d_bytes_per_step": 3 * B * CFG["nsample"] * 4 + B * 4,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e.item() / K},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof,
                "cpu_baseline": {"value": cpu_val, "unit": "utterances/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample} utterances x 2 steps, numpy oracle port (featurizer+fwd+loss), "
                                           f"{cpu_step:.2f} s/step"},
                "loss_mean_last": lv}
        if ms_train is not None:
            line["train"] = {"value": world * B * KT / (ms_train.item() / 1e3), "unit": "utterances/s",
                             "ms_per_step": ms_train.item() / KT, "steps": KT, "loss_last": train_loss,
                             "sync_bn": bool(args.sync_bn and world > 1),
                             "includes": "featurizer + forward + loss_dc + hand-written backward (BPTT) + "
                                         "bucketed NCCL gradient all-reduce (N>1) + clip_grad_norm_(5) + Adam "
                                         "(multi-tensor kernels of this repo)"}
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
    ap.add_argument("--sync-bn", action="store_true",
                    help="training figure with whole-batch BatchNorm statistics across ranks (N>1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
