"""Stage-level entry points with the reference's function names, batched on the GPU.

They exist for the parity ladder: each calls one kernel of
libqd_b200.so through the C ABI on ``[B, n]`` float32 CUDA tensors (NumPy arrays are copied
to the device and back).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, tables
from .pipeline import _torch


def _dev(x):
    torch = _torch()
    is_np = isinstance(x, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda() if is_np else x
    single = t.dim() == 1
    if single:
        t = t[None, :]
    if t.dim() != 2:
        raise ValueError("expected mono audio [n] or a batch [B, n]")
    return torch, t.contiguous().float(), is_np, single


def _back(t, is_np, single):
    if single:
        t = t[0]
    return t.cpu().numpy() if is_np else t


def peak_limiter(audio, sr: int, ceiling_db: float = -1.0, lookahead_ms: float = 5.0, release_ms: float = 50.0):
    """dsp/limiter.py:14-80 -> limited audio (float32).  The gain curve is not materialised."""
    torch, x, is_np, single = _dev(audio)
    ceiling, lookahead, coeff = tables.limiter_constants(sr, ceiling_db, lookahead_ms, release_ms)
    y = torch.empty_like(x)
    lib = _lib.load()
    _lib.check(lib.qd_limiter_device(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], lookahead, ceiling, coeff,
                                     torch.cuda.current_stream().cuda_stream))
    return _back(y, is_np, single)


def linkwitz_riley_split(audio, sr: int, crossover_hz: float):
    """dsp/crossover.py:71-118 -> (low, high) float32."""
    torch, x, is_np, single = _dev(audio)
    sos_lo, sos_hi = tables.design_linkwitz_riley_sos(sr, crossover_hz)
    arr_t = (C.c_double * 6) * 2
    lo_c, hi_c = arr_t(), arr_t()
    for r in range(2):
        for c in range(6):
            lo_c[r][c] = float(sos_lo[r, c])
            hi_c[r][c] = float(sos_hi[r, c])
    low, high = torch.empty_like(x), torch.empty_like(x)
    lib = _lib.load()
    _lib.check(lib.qd_crossover_device(x.data_ptr(), low.data_ptr(), high.data_ptr(), x.shape[0], x.shape[1],
                                       C.byref(lo_c), C.byref(hi_c), torch.cuda.current_stream().cuda_stream))
    return _back(low, is_np, single), _back(high, is_np, single)


def apply_distortion(audio, mode: str, fold_amount: float = 1.0, bias: float = 0.0, drive: float = 1.0,
                     warmth: float = 0.5):
    """dsp/distortion.py:93-114."""
    if mode not in ("wavefold", "tube"):
        raise ValueError(f"Unsupported distortion mode: {mode}")
    torch, x, is_np, single = _dev(audio)
    a = 1.0 + 4.0 * float(np.clip(warmth, 0.0, 1.0))
    y = torch.empty_like(x)
    lib = _lib.load()
    _lib.check(lib.qd_distort_device(x.data_ptr(), y.data_ptr(), x.numel(), 0 if mode == "wavefold" else 1,
                                     float(fold_amount), float(bias), a * max(float(drive), 0.0),
                                     1.0 / float(np.tanh(a)), torch.cuda.current_stream().cuda_stream))
    return _back(y, is_np, single)
