"""Small renders through every instantiation of the team kernel (qd_spec_team.cuh), for compute-sanitizer runs.
    compute-sanitizer --tool racecheck python profiles/sanitizer_team.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

sr = 48000
for n in (24000, 5003):
    x = torch.from_numpy(np.stack([synth.loud_clip(i, n, sr) for i in range(2)])).cuda()
    for n_fft, kw in ((512, {}), (1024, {}), (4096, {}), (4096, dict(precision="float64")), (8192, {}),
                      (4096, dict(passthrough_test=True)), (1024, dict(sub_cut_hz=0.0, air_cut_hz=0.0, precision="float32"))):
        y, _ = qd.process_batch(x, sr, n_fft=n_fft, quantize_mode="spectral_bins", **kw)
        torch.cuda.synchronize()
        print(n, n_fft, sorted(kw), float(y.abs().max()))
print("ok")
