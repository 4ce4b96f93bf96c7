"""Sweep the poll-delay knob of the recurrent kernel at cfg2 shape (us/step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onssen_b200 import _lib
B, T, H = 32, 400, 600
lib = _lib.load(); Hp = _lib.hp_of(H)
torch.manual_seed(0)
k = 1 / np.sqrt(H)
mk = lambda *s: (torch.rand(*s, device="cuda") * 2 - 1) * k
wf = (mk(4 * H, 2 * H), mk(4 * H, H), mk(4 * H), mk(4 * H)); wr = (mk(4 * H, 2 * H), mk(4 * H, H), mk(4 * H), mk(4 * H))
_, whh_p, _ = _lib.lstm_pack_layer(wf, wr, H, 2 * H, True, H)
gates = torch.randn(T * B, 8 * Hp, device="cuda")
y_h = torch.empty(T * B, 2 * Hp, device="cuda", dtype=torch.float16)
ws = _lib.blstm_rec_workspace(B, H, "cuda")
for delay in [0, 200, 400, 500, 600, 700, 800, 1000, 1200, 0]:
    lib.onssen_blstm_rec_set_poll_delay(delay)
    for _ in range(3):
        _lib.blstm_rec_fwd(gates, whh_p, B, T, H, y_h, None, 0.3, 1, 0, ws, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.blstm_rec_fwd(gates, whh_p, B, T, H, y_h, None, 0.3, 1, 0, ws, True)
    e1.record(); torch.cuda.synchronize()
    print(f"poll_delay={delay:5d}: {e0.elapsed_time(e1) / 10 * 1e3 / T:.3f} us/step", flush=True)
