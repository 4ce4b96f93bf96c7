"""Secondary measurement: chimera++ 4x600, B=64 (BASELINE configs[2]) forward + loss_chimera_psa, one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import onssen_b200 as ob
import bench
dev = torch.device("cuda:0"); torch.manual_seed(0)
B, T = 64, 400
model = ob.nn.chimera(129, 600, 4, 20).to(dev).train()
waves, starts = bench.synth_batch(0, B)
ws = [torch.from_numpy(w).to(dev) for w in waves]; st = torch.from_numpy(starts).to(dev)
def step():
    inp, lab = ob.data.featurize_batch(ws[0], ws[1], ws[2], "chimera++", 256, 64, T, 40.0, crop_start=st)
    return ob.loss.loss_chimera_psa(model(inp), lab).mean()
with torch.no_grad():
    for _ in range(5): step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): l = step()
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"cfg3 chimera++ 4x600 B=64 fwd+loss: {ms:.3f} ms/step = {B / ms * 1e3:.0f} utterances/s (loss {l.item():.1f})")
