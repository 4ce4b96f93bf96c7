"""CUDA-event timing of the featurizer launch at the BASELINE shapes (device-resident waveforms)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onssen_b200 import _lib
if os.environ.get("ONSSEN_LIB"):
    _lib.LIB_PATH = os.environ["ONSSEN_LIB"]
torch.manual_seed(0)
for (B, ns, n_fft, hop, want) in [(32, 32000, 256, 64, ["feature", "mag_mix", "mag_s1", "mag_s2", "feat_max"]),
                                  (64, 32000, 256, 64, ["feature", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2", "feat_max"]),
                                  (16, 64000, 512, 128, ["feature", "mag_mix", "mag_s1", "mag_s2", "ph_mix", "ph_s1", "ph_s2", "feat_max"]),
                                  (32, 64000, 1024, 256, ["feature", "mag_mix", "mag_s1", "mag_s2", "cos_s1", "cos_s2", "feat_max"])]:
    T = 400
    w = [torch.randn(B, ns, device="cuda") * 0.1 for _ in range(3)]
    st = torch.zeros(B, dtype=torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = _lib.stft_features(w[0], w[1], w[2], n_fft, hop, st, T, want)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    F = n_fft // 2 + 1
    out_bytes = sum(o[k].numel() * 4 for k in want)
    in_bytes = 3 * B * min(ns, (T - 1) * hop + n_fft) * 4
    best = min(ms[2:])
    print(f"B={B} n_fft={n_fft} hop={hop}: {best*1e3:.1f} us  ({(in_bytes+out_bytes)/best/1e6:.0f} GB/s algorithmic, "
          f"{(in_bytes+out_bytes)/1e6:.1f} MB)")
