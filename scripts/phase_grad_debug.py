"""Stage-by-stage comparison of the phase_net / loss_phase backward against the repaired torch restatement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import onssen_b200 as ob
from oracle import ref_loader
from oracle.phase_repaired import build
from test_reference_gpu import synth_inputs, rel_err
dev = torch.device("cuda:0")
R = ref_loader.import_reference()
PhaseNetRepaired, loss_phase_repaired = build(R)
torch.manual_seed(14)
F, H, L, D, B = 257, 300, 3, 20, int(os.environ.get("B", 16))
T = int(os.environ.get("T", 400))
ours = ob.nn.phase_net(F, H, L, D, dropout=0.0).to(dev)
ref = PhaseNetRepaired(F, H, L, D)
if os.environ.get("BIAS", "1") == "1":
    with torch.no_grad():
        b = ours.chimera.fc_mi.bias.view(F, 2); b[:, 0] += 1.0; b[:, 1] -= 1.0
import test_reference_gpu as TR
inp, lab = synth_inputs("phase", B, 512, 128, 64000, T, dev, first=300)
ref.load_state_dict({k: v.detach().cpu() for k, v in ours.state_dict().items()})
ours.train(); ref.train()
out = ours(inp)
for o in out: o.retain_grad()
lo = ob.loss.loss_phase(out, lab)
torch.mean(lo).backward()
out_r = ref([t.cpu() for t in inp])
for o in out_r: o.retain_grad()
lr = loss_phase_repaired(out_r, [t.cpu() for t in lab])
torch.mean(lr).backward()
names = ["emb", "mask_A", "mask_B", "phase_A", "phase_B"]
for n, a, b in zip(names, out, out_r):
    ga = a.grad if a.grad is not None else torch.zeros_like(a)
    print(f"d loss / d {n:8s}: rel err {rel_err(ga.cpu(), b.grad):.3e}   |ref| {float(b.grad.norm()):.3e} |ours| {float(ga.norm()):.3e}")
rg = dict(ref.named_parameters())
for k, p in ours.named_parameters():
    print(f"{k:40s} rel err {rel_err(p.grad.cpu(), rg[k].grad):.3e}  |ref| {float(rg[k].grad.norm()):.3e} |ours| {float(p.grad.norm()):.3e}")
