"""Secondary measurements at the other BASELINE.json configs, one GPU each (per-GPU shard of the 8-GPU configs):
forward+loss (utterances/s) and the full training step (featurizer + forward + loss + hand-written backward +
clip_grad_norm_(5) + Adam).  cfg3: chimera++ 4x600 B=64; cfg4: phase_net (repaired) 16 kHz n_fft 512, per-GPU B=16
(global 128 / 8), H=300 and H=600; cfg5: enhance(513, 600, 3) 16 kHz n_fft 1024, per-GPU B=32 (global 256 / 8)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import onssen_b200 as ob
from oracle import onssen_oracle as O

dev = torch.device("cuda:0"); torch.manual_seed(0)
T = 400


def waves(B, nsample):
    utts = [O.synth_utterance(i, nsample) for i in range(B)]
    return [torch.from_numpy(np.stack([u[k] for u in utts])).to(dev) for k in range(3)]


def run(name, model, B, nsample, n_fft, hop, feat_name, to_io, loss_fn, train=True):
    ws = waves(B, nsample)
    from onssen_b200.data.feature_utils import num_crop_starts
    nst = num_crop_starts(nsample, hop, T)
    st = torch.from_numpy(np.array([np.random.RandomState(1234 + i).randint(nst) for i in range(B)], dtype=np.int32)).to(dev)
    def batch():
        inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], feat_name, n_fft, hop, T, 40.0, crop_start=st)
        return to_io(inp, lab)
    def timeit(fn, n):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out
    model.eval()
    def fwd():
        with torch.no_grad():
            inp, lab = batch()
            return loss_fn(model(inp), lab).mean()
    ms, l = timeit(fwd, 10)
    line = f"{name}: fwd+loss {ms:.3f} ms/step = {B / ms * 1e3:.0f} utterances/s (loss {l.item():.4g})"
    if train:
        model.train()
        opt = ob.utils.build_optimizer(model.parameters(), {"name": "adam", "lr": 1e-3})
        def step():
            inp, lab = batch()
            loss = loss_fn(model(inp), lab).mean()
            opt.zero_grad()
            loss.backward()
            ob.utils.clip_grad_norm_(model.parameters(), 5)
            opt.step()
            return loss
        ms, l = timeit(step, 5)
        line += f"; train step {ms:.3f} ms = {B / ms * 1e3:.0f} utterances/s (loss {l.item():.4g})"
    print(line, flush=True)


which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
if "cfg3" in which:
    run("cfg3 chimera++ 4x600 B=64", ob.nn.chimera(129, 600, 4, 20).to(dev), 64, 32000, 256, 64, "chimera++",
        lambda i, l: (i, l), ob.loss.loss_chimera_psa)
if "cfg4" in which:
    for H in (300, 600):
        run(f"cfg4 phase_net(257, H={H}, L=3, D=20) per-GPU B=16", ob.nn.phase_net(257, H, 3, 20).to(dev), 16, 64000, 512,
            128, "phase", lambda i, l: (i, l), ob.loss.loss_phase)
if "cfg5" in which:
    run("cfg5 enhance(513, 600, 3) per-GPU B=32", ob.nn.enhance(513, 600, 3).to(dev), 32, 64000, 1024, 256, "chimera++",
        lambda i, l: ([i[0], l[1]], [l[2], l[4]]), ob.loss.loss_mask_msa)
