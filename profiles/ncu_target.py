"""Small, fixed workload for ncu captures: one wave of whole-clip CTAs of the default single-band render.
    python profiles/ncu_target.py [clips] [repeats]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = synth.bass_batch_torch(clips, 480000, 48000, "cuda", seed=0)
r = qd.make_renderer(480000, 48000, quantize_mode="spectral_bins")
for _ in range(reps):
    y, _ = r.render_device(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
