"""Small renders covering every kernel variant, for compute-sanitizer (memcheck / racecheck) runs.
    compute-sanitizer --tool memcheck python profiles/sanitizer_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

n, sr = 16000, 48000
x = torch.from_numpy(np.stack([synth.loud_clip(i, n, sr) for i in range(3)])).cuda()
GROWL = dict(key="F", scale="minor", snap_strength=0.9, smear=0.3, distortion_params={"fold_amount": 5.0, "bias": 0.1})
cases = [({}, 2048), (dict(use_multiband=True), 2048), (dict(GROWL, use_multiband=True, spectral_fx_mode="bitcrush", spectral_fx_strength=0.5), 2048),
         (dict(GROWL, use_multiband=True, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.55), 2048),
         (dict(spectral_freeze=True), 2048), (dict(formant_shift=3.0), 2048), (dict(formant_shift=-4.0, spectral_freeze=True), 1024),
         (dict(GROWL, use_multiband=True, formant_shift=2.0, spectral_fx_mode="phase_dispersal", spectral_fx_strength=0.6), 2048), ({}, 512), ({}, 8192), (dict(precision="float64"), 2048), (dict(passthrough_test=True), 1024)]
for kw, n_fft in cases:
    y, _ = qd.process_batch(x, sr, n_fft=n_fft, seeds=7, quantize_mode="spectral_bins", **kw)
    torch.cuda.synchronize()
    print(n_fft, sorted(kw)[:3], float(y.abs().max()))
xo = torch.from_numpy(np.stack([synth.loud_clip(9, 5003, sr)])).cuda()   # ragged length: scalar load/store paths
y, _ = qd.process_batch(xo, sr, quantize_mode="spectral_bins")
torch.cuda.synchronize()
from quantumdistortion_b200 import analyses
for prec in ("float64", "float32"):
    for nf in (512, 2048, 8192):
        b = analyses.spectral_peak_bins(x, n_fft=nf, topn=5, precision=prec)
        torch.cuda.synchronize()
print("peaks", tuple(b.shape))
print("ok")
