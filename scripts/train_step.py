"""Two training steps at cfg2 (for ncu launch lists of the backward)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import onssen_b200 as ob
import bench
torch.manual_seed(0)
dev = torch.device("cuda:0")
C = bench.CFG
model = ob.nn.deep_clustering(*C["margs"]).to(dev).train()
waves, starts = bench.synth_batch(0, C["B"])
ws = [torch.from_numpy(w).to(dev) for w in waves]; st = torch.from_numpy(starts).to(dev)
opt = ob.utils.build_optimizer(model.parameters(), {"name": "adam", "lr": 1e-3})
for it in range(int(os.environ.get("STEPS", 2))):
    inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], "dc", C["n_fft"], C["hop"], bench.T_FRAMES, bench.DB, crop_start=st)
    loss = torch.mean(ob.loss.loss_dc(model(inp), lab))
    opt.zero_grad(); loss.backward()
    ob.utils.clip_grad_norm_(model.parameters(), 5); opt.step()
torch.cuda.synchronize(); print("loss", loss.item())
