"""Timing target for the float64 FX pass at the default n_fft: spectral_freeze (precision="auto" -> float64).
    python profiles/ncu_target_fx64.py [clips]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 296
x = synth.bass_batch_torch(clips, 480000, 48000, "cuda", seed=0)
r = qd.make_renderer(480000, 48000, quantize_mode="spectral_bins", spectral_freeze=True)
for _ in range(2):
    y, _ = r.render_device(x)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(3):
    y, _ = r.render_device(x)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 3
print("ok freeze", clips, "clips", float(y.abs().max()), "ms per render", ms, "audio-s/s", clips * 10.0 / ms * 1e3)
