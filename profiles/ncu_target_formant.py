"""Fixed workload for ncu captures of the formant-shift variant of the spectral pass.
    python profiles/ncu_target_formant.py [clips]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 148
x = synth.bass_batch_torch(clips, 480000, 48000, "cuda", seed=0)
r = qd.make_renderer(480000, 48000, formant_shift=3.0)
for _ in range(2):
    y, _ = r.render_device(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
