"""Device-resident throughput of the other BASELINE.json configs (one GPU), for the record in profiles/.
    python profiles/bench_configs.py [clips]
Prints one JSON line per configuration: audio-s/s with inputs resident in HBM (CUDA events, 3 warm-up + 3 timed).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import quantumdistortion_b200 as qd
from quantumdistortion_b200 import synth

SR, N = 48000, 480000
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = synth.bass_batch_torch(clips, N, SR, "cuda", seed=3)
GROWL = dict(key="F", scale="minor", snap_strength=0.9, smear=0.3, distortion_mode="wavefold",
             distortion_params={"fold_amount": 5.0, "bias": 0.1, "drive": 1.0, "warmth": 0.5}, limiter_ceiling_db=-1.0)
CONFIGS = [
    ("config2 single-band defaults", {}, 2048, None),
    ("config3 multiband LR4 @300 Hz", dict(use_multiband=True, crossover_hz=300.0), 2048, None),
    ("config4 growl + multiband + bitcrush 0.5", dict(GROWL, use_multiband=True, spectral_fx_mode="bitcrush", spectral_fx_strength=0.5), 2048, 1234),
    ("config4 growl + multiband + phase_dispersal 0.6", dict(GROWL, use_multiband=True, spectral_fx_mode="phase_dispersal", spectral_fx_strength=0.6), 2048, 1234),
    ("config4 growl + multiband + bin_scramble 0.55", dict(GROWL, use_multiband=True, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.55), 2048, 1234),
    ("next(f1) single-band + formant_shift +3 st", dict(formant_shift=3.0), 2048, None),
    ("next(f1) single-band + spectral_freeze", dict(spectral_freeze=True), 2048, None),
    ("config5 n_fft 512", {}, 512, None),
    ("config5 n_fft 1024", {}, 1024, None),
    ("config5 n_fft 4096", {}, 4096, None),
    ("config5 n_fft 8192 (float64 kernels, precision=auto)", {}, 8192, None),
    ("config5 n_fft 8192 float32 kernels", {"precision": "float32"}, 8192, None),
    ("config2 with float64 kernels", {"precision": "float64"}, 2048, None),
]
# the reference's default mode (time-domain)
AT_CLIPS = clips
for name, kw in (("next(f3) autotune_v1 defaults (bass batch)", {}), ("next(f3) autotune_v1, growl distortion, no sub", dict(GROWL, sub_enabled=False))):
    kw = {k: v for k, v in kw.items() if k not in ("smear",)}
    xs = x[:AT_CLIPS]
    r = qd.make_renderer(N, SR, quantize_mode="autotune_v1", **kw)
    for _ in range(2):   # the second warm-up lets torch's allocator create the second output block outside the timing
        y, _ = r.render_device(xs, chunk_clips=2048)
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y, _ = r.render_device(xs, chunk_clips=2048)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[1]
    print(json.dumps({"config": name, "clips": AT_CLIPS, "ms_per_render": round(ms, 3), "ms_all": [round(t, 1) for t in times],
                      "audio_s_per_s": round(AT_CLIPS * 10 / (ms / 1e3)), "finite": bool(torch.isfinite(y).all())}), flush=True)
    del r, y
for name, kw, n_fft, seed in CONFIGS:
    b = clips if n_fft <= 4096 and kw.get("precision") != "float64" else max(64, clips // 4)
    xs = x[:b]
    r = qd.make_renderer(N, SR, n_fft, seeds=seed, quantize_mode="spectral_bins", **kw)
    r.set_fx_seeds(b, seed)
    for _ in range(3):
        y, _ = r.render_device(xs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y, _ = r.render_device(xs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"config": name, "clips": b, "n_fft": n_fft, "ms_per_render": round(ms, 3),
                      "audio_s_per_s": round(b * 10 / (ms / 1e3)), "finite": bool(torch.isfinite(y).all())}))
    del r, y
