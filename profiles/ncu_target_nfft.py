"""ncu target for the n_fft sweep (BASELINE configs[4]): one wave of whole-clip CTAs of the single-band render.
    python profiles/ncu_target_nfft.py <n_fft> [clips] [repeats]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

n_fft = int(sys.argv[1])
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 296
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x = synth.bass_batch_torch(clips, 480000, 48000, "cuda", seed=0)
r = qd.make_renderer(480000, 48000, n_fft=n_fft, quantize_mode="spectral_bins")
for _ in range(reps):
    y, _ = r.render_device(x)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(3):
    y, _ = r.render_device(x)
t1.record()
torch.cuda.synchronize()
print("ok", n_fft, clips, "clips", float(y.abs().max()), "ms per render", t0.elapsed_time(t1) / 3)
