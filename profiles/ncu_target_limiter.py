"""ncu target for the limiter: a loud batch (every clip crosses the ceiling) through the stage entry point.
    python profiles/ncu_target_limiter.py [clips] [repeats]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x = torch.clamp(6.0 * synth.bass_batch_torch(clips, 480000, 48000, "cuda", seed=0), -1.5, 1.5)
for _ in range(reps):
    y = qd.peak_limiter(x, 48000, ceiling_db=-6.0, lookahead_ms=5.0, release_ms=30.0)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    y = qd.peak_limiter(x, 48000, ceiling_db=-6.0, lookahead_ms=5.0, release_ms=30.0)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 5
print("ok", clips, "clips peak", float(y.abs().max()), "ms", ms, "GB/s (8 B/sample)", 8 * clips * 480000 / ms / 1e6)
