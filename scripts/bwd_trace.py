"""Persistent BPTT kernel at cfg2 shape: launch timing + clock64 timeline of CTA 0 (steps 100..107)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onssen_b200 import _lib
if os.environ.get("ONSSEN_LIB"):
    _lib.LIB_PATH = os.environ["ONSSEN_LIB"]      # A/B runs against another build of the library

B, T, H = int(os.environ.get("B", 32)), 400, int(os.environ.get("H", 600))
lib = _lib.load()
Hp = _lib.hp_of(H)
torch.manual_seed(0)
k = 1 / np.sqrt(H)
whh_t = _lib.lstm_pack_whh_t((torch.rand(4 * H, H, device="cuda") * 2 - 1) * k, (torch.rand(4 * H, H, device="cuda") * 2 - 1) * k, H)
M = T * B
act0 = torch.rand(M, 8 * Hp, device="cuda")
c = torch.randn(M, 2 * Hp, device="cuda")
dy = torch.randn(M, 2 * Hp, device="cuda") * 1e-3
sc = _lib.amax_scale(dy, target=0.0625)
dg16 = torch.empty(M, 8 * Hp, device="cuda", dtype=torch.float16)
trace = torch.zeros(512, device="cuda", dtype=torch.int64)
names = {0: "step start", 1: "dG fragments fresh (warp 0)", 2: "MMAs done, partials in smem", 3: "after barrier",
         4: "gate math + publish issued", 5: "MMAs done (all chunks)", 6: "publish issued"}
MODE = int(os.environ.get("MODE", 2))
if os.environ.get("RESERVE") is not None:
    lib.onssen_blstm_rec_bwd_set_sm_reserve(int(os.environ["RESERVE"]))
tc_names = {0: "step start", 1: "dG quarter gathered", 2: "fenced + arrived", 3: "mma warp: tile ready", 4: "mma warp: issued + commit",
            5: "mma done", 6: "partials sent (DSMEM)", 7: "4 partials received", 8: "dG published", 9: "dG stored"}
for persistent in (MODE,):
    lib.onssen_blstm_rec_bwd_set_persistent(persistent)
    act = act0.clone()
    lib.onssen_blstm_rec_bwd_set_trace(ctypes.c_void_p(trace.data_ptr()))
    _lib.blstm_rec_bwd(act, dg16, c, dy, whh_t, sc, B, T, H, 0.3, 1, 0)
    torch.cuda.synchronize()
    lib.onssen_blstm_rec_bwd_set_trace(None)
    if persistent == 2:
        tr = trace.cpu().numpy()[:64].reshape(4, 16)
        for s in range(1, 4):
            t0 = tr[s][0]
            print(f"--- step {100 + s}: period {t0 - tr[s - 1][0]} cycles (thread 0 step start = 0)")
            for slot in sorted(tc_names, key=lambda q: tr[s][q]):
                print(f"   {tc_names[slot]:32s} {tr[s][slot] - t0:7d}")
    tr = trace.cpu().numpy().reshape(8, 8, 8)          # [step][slot][warp]
    for s in range(1, 3) if persistent != 2 else ():
        t0 = tr[s][0][0]
        print(f"--- step {100 + s}: period {t0 - tr[s - 1][0][0]} cycles (warp 0 step start = 0)")
        for slot in (0, 1, 5, 2, 3, 6, 4):
            print(f"   {names[slot]:32s} " + " ".join(f"{tr[s][slot][w] - t0:6d}" for w in range(8)))
    for _ in range(2):
        _lib.blstm_rec_bwd(act, dg16, c, dy, whh_t, sc, B, T, H, 0.3, 1, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.blstm_rec_bwd(act, dg16, c, dy, whh_t, sc, B, T, H, 0.3, 1, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"persistent={persistent} B={B} H={H}: {ms:.3f} ms per launch = {ms * 1e3 / T:.2f} us/step")
