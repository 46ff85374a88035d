"""Fixed workload for ncu captures of the spectral-FX variant of the pass (config 4: growl + multiband + FX).
    python profiles/ncu_target_fx.py [clips] [fx_mode] [strength]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 148
mode = sys.argv[2] if len(sys.argv) > 2 else "bitcrush"
strength = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
x = synth.bass_batch_torch(clips, 480000, 48000, "cuda", seed=0)
r = qd.make_renderer(480000, 48000, seeds=1234, key="F", scale="minor", snap_strength=0.9, smear=0.3,
                     distortion_params={"fold_amount": 5.0, "bias": 0.1}, use_multiband=True,
                     spectral_fx_mode=mode, spectral_fx_strength=strength)
r.set_fx_seeds(clips, 1234)
for _ in range(2):
    y, _ = r.render_device(x)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
