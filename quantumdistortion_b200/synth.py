"""Seeded synthetic clips shared by the golden generator, the parity tests and bench.py.

Formulas follow SURVEY.md section 8(d) config #2 (cf. the reference's
scripts/generate_test_fixtures.py:82-89 for the bass-with-LFO idea).
"""
from __future__ import annotations

import numpy as np


def bass_clip(seed: int, n: int, sr: int = 48000) -> np.ndarray:
    """Harmonic 'bass' clip: f0~U(40,110) Hz, 40 harmonics 1/h, LFO amplitude, light noise."""
    rng = np.random.default_rng(1000 + seed)
    t = np.arange(n, dtype=np.float64) / sr
    f0 = rng.uniform(40.0, 110.0)
    fl = rng.uniform(1.0, 5.0)
    ph = rng.uniform(0.0, 2.0 * np.pi, size=40)
    x = np.zeros(n)
    for h in range(1, 41):
        x += np.sin(2.0 * np.pi * f0 * h * t + ph[h - 1]) / h
    x *= 1.0 - 0.4 + 0.4 * (1.0 + np.sin(2.0 * np.pi * fl * t)) / 2.0
    x += 0.01 * rng.standard_normal(n)
    x *= 0.6 / max(np.max(np.abs(x)), 1e-12)
    nf = min(int(0.05 * sr), n // 2)
    if nf > 0:
        ramp = np.linspace(0.0, 1.0, nf)
        x[:nf] *= ramp
        x[-nf:] *= ramp[::-1]
    return x.astype(np.float32)


def noise_clip(seed: int, n: int) -> np.ndarray:
    """0.3*N(0,1) clipped to +-1: stresses limiter and quantizer."""
    rng = np.random.default_rng(5000 + seed)
    return np.clip(0.3 * rng.standard_normal(n), -1.0, 1.0).astype(np.float32)


def loud_clip(seed: int, n: int, sr: int = 48000) -> np.ndarray:
    """Bass clip scaled to exceed the limiter ceiling often."""
    return np.clip(1.6 * bass_clip(seed, n, sr), -1.5, 1.5).astype(np.float32)


def tone_clip(seed: int, n: int, sr: int = 48000) -> np.ndarray:
    """Pitched clip for the autotune path: two consecutive notes (f0~U(150,420) Hz, usually off the scale), eight
    harmonics 1/h, 5.5 Hz vibrato of +-12 cents, a short silence in the middle, light noise, peak 0.5."""
    rng = np.random.default_rng(9000 + seed)
    t = np.arange(n, dtype=np.float64) / sr
    f = np.where(t < t[-1] * 0.55, rng.uniform(150.0, 420.0), rng.uniform(150.0, 420.0))
    f = f * 2.0 ** (0.01 * np.sin(2.0 * np.pi * 5.5 * t))
    ph = 2.0 * np.pi * np.cumsum(f) / sr
    x = np.zeros(n)
    for h in range(1, 9):
        x += np.sin(h * ph + rng.uniform(0.0, 2.0 * np.pi)) / h
    gap = (t > t[-1] * 0.5) & (t < t[-1] * 0.55)
    x[gap] = 0.0
    x += 0.002 * rng.standard_normal(n)
    x *= 0.5 / max(np.max(np.abs(x)), 1e-12)
    return x.astype(np.float32)


def bass_batch_torch(batch: int, n: int, sr: int, device, seed: int = 0, rows_per_step: int = 64):
    """Device-side version of ``bass_clip`` for large benchmark batches (same recipe, torch RNG):
    returns float32 ``[batch, n]`` on ``device``.  Built ``rows_per_step`` clips at a time."""
    import math

    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(1000 + seed)
    out = torch.empty(batch, n, dtype=torch.float32, device=device)
    t = torch.arange(n, dtype=torch.float32, device=device) / float(sr)
    nf = min(int(0.05 * sr), n // 2)
    ramp = torch.linspace(0.0, 1.0, nf, device=device) if nf > 0 else None
    for b0 in range(0, batch, rows_per_step):
        rows = min(rows_per_step, batch - b0)
        f0 = 40.0 + 70.0 * torch.rand(rows, 1, generator=gen, device=device)
        fl = 1.0 + 4.0 * torch.rand(rows, 1, generator=gen, device=device)
        ph = 2.0 * math.pi * torch.rand(rows, 40, generator=gen, device=device)
        x = torch.zeros(rows, n, dtype=torch.float32, device=device)
        base = 2.0 * math.pi * f0 * t[None, :]
        for h in range(1, 41):
            x += torch.sin(base * h + ph[:, h - 1:h]) / h
        x *= 0.6 + 0.4 * (1.0 + torch.sin(2.0 * math.pi * fl * t[None, :])) / 2.0
        x += 0.01 * torch.randn(rows, n, generator=gen, device=device)
        x *= 0.6 / x.abs().amax(dim=1, keepdim=True).clamp_min(1e-12)
        if nf > 0:
            x[:, :nf] *= ramp[None, :]
            x[:, n - nf:] *= ramp.flip(0)[None, :]
        out[b0:b0 + rows] = x
    return out
