"""Timing / ncu target for the float64 plain pass at the default n_fft: the wide-open band mask (sub_cut_hz = air_cut_hz = 0,
what the reference's committed *_multiband.wav renders use) makes precision="auto" choose the float64 kernels.
    python profiles/ncu_target_f64.py [clips] [n_fft]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n_fft = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
x = synth.bass_batch_torch(clips, 480000, 48000, "cuda", seed=0)
r = qd.make_renderer(480000, 48000, n_fft=n_fft, quantize_mode="spectral_bins", sub_cut_hz=0.0, air_cut_hz=0.0)
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
print("ok", n_fft, clips, "clips", float(y.abs().max()), "ms per render", ms, "audio-s/s", clips * 10.0 / ms * 1e3)
