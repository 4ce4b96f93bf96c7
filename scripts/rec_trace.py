"""Timeline of one recurrent step (clock64 stamps of CTA 0, steps 100..103) + launch timing at cfg2 shape."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onssen_b200 import _lib
if os.environ.get("ONSSEN_LIB"):
    _lib.LIB_PATH = os.environ["ONSSEN_LIB"]      # A/B runs against another build of the library

B, T, H = int(os.environ.get("B", 32)), 400, 600
lib = _lib.load()
Hp = _lib.hp_of(H)
torch.manual_seed(0)
k = 1 / np.sqrt(H)
mk = lambda *s: (torch.rand(*s, device="cuda") * 2 - 1) * k
wf = (mk(4 * H, 2 * H), mk(4 * H, H), mk(4 * H), mk(4 * H)); wr = (mk(4 * H, 2 * H), mk(4 * H, H), mk(4 * H), mk(4 * H))
_, whh_p, _ = _lib.lstm_pack_layer(wf, wr, H, 2 * H, True, H)
gates = torch.randn(T * B, 8 * Hp, device="cuda")
y_h = torch.empty(T * B, 2 * Hp, device="cuda", dtype=torch.float16)
ws = _lib.blstm_rec_workspace(B, H, "cuda")
trace = torch.zeros(640, device="cuda", dtype=torch.int64)
if os.environ.get("POLL_DELAY"):
    lib.onssen_blstm_rec_set_poll_delay(int(os.environ["POLL_DELAY"]))
for tc in (True,):
    lib.onssen_blstm_rec_set_trace(ctypes.c_void_p(trace.data_ptr()))
    _lib.blstm_rec_fwd(gates, whh_p, B, T, H, y_h, None, 0.3, 1, 0, ws, tc)
    torch.cuda.synchronize()
    lib.onssen_blstm_rec_set_trace(None)
    full = trace.cpu().numpy()
    tr = full[:64].reshape(4, 16)
    gt = full[64:64 + 19 * 8].reshape(19, 8)
    t0 = gt[:, 0].min()
    print('group trace step 102 (ns rel. to first mma-done): rb: mma_done, publish, gathered')
    for rb in range(19):
        print(f'   rb{rb:2d}: {gt[rb,0]-t0:6d} {gt[rb,1]-t0:6d} {gt[rb,2]-t0:6d}   sm {gt[rb,3]}')
    print(f'   publish spread {gt[:,1].max()-gt[:,1].min()} ns; last publish -> first gathered {gt[:,2].min()-gt[:,1].max()} ns; last gathered {gt[:,2].max()-gt[:,1].max()} ns')
    names = {3: "mma warp: h tile ready (bar3)", 4: "mma warp: issued+commit", 8: "gate: mma done", 9: "gate: tmem loaded",
             10: "gate: act+xchg written", 11: "gate: c/h + LL publish", 12: "gate: h gathered", 13: "gate: fenced+arrived"}
    for s in range(1, 4):
        base = tr[s][3]
        print(f"--- step {100 + s} (cycles relative to 'h tile ready')")
        for slot in sorted(names, key=lambda q: tr[s][q]):
            print(f"   {names[slot]:32s} {tr[s][slot] - base:8d}")
        print(f"   step period: {tr[s][3] - tr[s-1][3]}")
    wt = full[320:320 + 4 * 64].reshape(4, 64)
    if wt.any():
        for s in range(1, 3):
            base = wt[s][8 * 6 + 0]          # MMA warp: first producer group ready
            print(f"--- per-warp trace step {100 + s} (cycles rel. to the MMA warp's first-ready); gate warps: mma-done, "
                  f"published, gather complete, signalled (after fence); MMA warp: first ready, last ready, committed")
            for w in range(8):
                print(f"   warp {w}: " + " ".join(f"{int(wt[s][w * 6 + k] - base):7d}" for k in range(4)))
            print("   mma   : " + " ".join(f"{int(wt[s][8 * 6 + k] - base):7d}" for k in range(3)))
    for _ in range(3):
        _lib.blstm_rec_fwd(gates, whh_p, B, T, H, y_h, None, 0.3, 1, 0, ws, tc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.blstm_rec_fwd(gates, whh_p, B, T, H, y_h, None, 0.3, 1, 0, ws, tc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"tc={tc} B={B}: {ms:.3f} ms per launch = {ms * 1e3 / T:.2f} us/step")
