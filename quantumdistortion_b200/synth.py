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
