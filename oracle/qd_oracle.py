"""
oracle/qd_oracle.py -- CPU oracle for the STFT processing path.  TEST INFRASTRUCTURE.

This file restates, in float64 NumPy, the algorithm of the reference's
``process_audio(..., quantize_mode="spectral_bins")`` path so that parity can be
checked on the GPU box, where ``/root/reference`` does not exist.  It is NOT the
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``quantumdistortion_b200``) never imports anything from ``oracle/``.

Parity pin: ``tests/golden/make_golden.py`` imports the live reference in the
build container and stores its outputs; ``tests/test_oracle.py`` checks this
restatement against those fixtures and against the reference's own known-answer
tests (tests/test_quantizer.py:10-131 etc.).  See DESIGN.md "Oracle".

Third-party arithmetic the reference delegates to (un-pinned in its
requirements.txt; this image carries numpy 2.3.5 / scipy 1.18.1) is called
directly, as the reference does: ``np.fft.rfft/irfft`` (pocketfft),
``scipy.signal.windows.hann``, ``scipy.signal.butter``.  The two per-sample loops
(limiter, SOS cascade) are restated in C in ``oracle/qd_seq.c``.

All ``file:line`` citations are relative to ``/root/reference/quantum_distortion``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Any, Dict, Optional, Tuple

import numpy as np
from scipy.signal import butter
from scipy.signal import windows as _windows

_HERE = os.path.dirname(os.path.abspath(__file__))
_SEQ_LIB = None

N_FFT_DEFAULT = 2048  # dsp/pipeline.py:149

NOTE_NAMES_SHARP = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]  # dsp/quantizer.py:29
SCALE_INTERVALS = {  # dsp/quantizer.py:32-39
    "major": (0, 2, 4, 5, 7, 9, 11),
    "minor": (0, 2, 3, 5, 7, 8, 10),
    "pentatonic": (0, 2, 4, 7, 9),
    "dorian": (0, 2, 3, 5, 7, 9, 10),
    "mixolydian": (0, 2, 4, 5, 7, 9, 10),
    "harmonic_minor": (0, 2, 3, 5, 7, 8, 11),
}
ROLE_WEIGHTS = {"root": 1.0, "fifth": 0.8, "third": 0.7, "seventh": 0.6, "other": 0.5}  # dsp/quantizer.py:42-48


# --------------------------------------------------------------------------- C helper
def _seq_lib() -> ctypes.CDLL:
    """Load (building on first use) oracle/libqd_oracle_seq.so."""
    global _SEQ_LIB
    if _SEQ_LIB is not None:
        return _SEQ_LIB
    so = os.path.join(_HERE, "libqd_oracle_seq.so")
    src = os.path.join(_HERE, "qd_seq.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.qd_oracle_limiter_gain.argtypes = [dp, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.c_double, ctypes.c_double, dp]
    lib.qd_oracle_limiter_gain.restype = None
    lib.qd_oracle_sosfilt.argtypes = [dp, ctypes.c_int, dp, ctypes.c_int64, dp]
    lib.qd_oracle_sosfilt.restype = None
    _SEQ_LIB = lib
    return lib


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------- STFT / iSTFT
def stft(x: np.ndarray, sr: int, n_fft: int = 2048, center: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """dsp/stft_utils.py:11-97.  Returns (S[bins, frames] complex128, freqs[bins])."""
    x = np.asarray(x, dtype=float)
    if x.ndim != 1:
        raise ValueError("stft_mono expects mono (1D) audio")  # :47
    hop = n_fft // 4  # :51
    w = _windows.hann(n_fft, sym=False)  # :56
    if center:
        x = np.pad(x, n_fft // 2, mode="constant", constant_values=0.0)  # :59-62
    n = len(x)
    n_frames = max(1, 1 + (n - n_fft) // hop)  # :67
    need = (n_frames - 1) * hop + n_fft
    if need > n:  # last frame zero-filled (:80-83)
        x = np.concatenate([x, np.zeros(need - n)])
    frames = np.lib.stride_tricks.sliding_window_view(x, n_fft)[::hop][:n_frames]
    S = np.fft.rfft(frames * w[None, :], n=n_fft, axis=1).T  # :89-92 (per-frame rfft)
    freqs = np.fft.rfftfreq(n_fft, d=1.0 / sr)  # :95
    return np.ascontiguousarray(S), freqs


def istft(S: np.ndarray, sr: int, n_fft: int = 2048, length: Optional[int] = None,
          center: bool = True) -> np.ndarray:
    """dsp/stft_utils.py:100-234.  Returns float32[length]."""
    hop = n_fft // 4
    w = _windows.hann(n_fft, sym=False)  # :141
    n_bins, n_frames = S.shape
    if n_bins != n_fft // 2 + 1:
        raise ValueError(f"STFT matrix has {n_bins} bins, expected {n_fft // 2 + 1}")  # :146
    natural = (n_frames - 1) * hop + n_fft  # :150
    if length is not None:
        out_len = max(natural, length + 2 * (n_fft // 2)) if center else max(natural, length)  # :154-162
    else:
        out_len = natural
    y = np.zeros(out_len)
    wss = np.zeros(out_len)
    w2 = w ** 2
    frames = np.fft.irfft(S.T, n=n_fft, axis=1) * w[None, :]  # :193-197
    for t in range(n_frames):  # same accumulation order as :175-183 and :190-209
        a = t * hop
        b = min(a + n_fft, out_len)
        wss[a:b] += w2[: b - a]
        y[a:b] += frames[t, : b - a]
    wss = np.maximum(wss, 1e-10)  # :186
    y = y / wss  # :214
    if center:
        p = n_fft // 2
        if len(y) > 2 * p:
            y = y[p:-p]  # :219-222
    if length is not None and len(y) != length:  # :227-232
        y = y[:length] if len(y) > length else np.pad(y, (0, length - len(y)))
    return y.astype(np.float32)  # :234


# --------------------------------------------------------------------------- scale tables
def note_name_to_pitch_class(name: str) -> int:
    """dsp/quantizer.py:59-69."""
    name = name.strip().upper()
    for flat, sharp in (("DB", "C#"), ("EB", "D#"), ("GB", "F#"), ("AB", "G#"), ("BB", "A#")):
        name = name.replace(flat, sharp)
    if name not in NOTE_NAMES_SHARP:
        raise ValueError(f"Unsupported key name: {name}")
    return NOTE_NAMES_SHARP.index(name)


def _freq_to_midi(f: float) -> float:
    return -np.inf if f <= 0.0 else 69.0 + 12.0 * np.log2(f / 440.0)  # dsp/quantizer.py:76-79


def scale_notes(key: str, scale: str, fmin: float, fmax: float) -> Tuple[np.ndarray, np.ndarray]:
    """dsp/quantizer.py:98-124 -> (note_freqs, note_weights)."""
    root = note_name_to_pitch_class(key)
    intervals = SCALE_INTERVALS[scale]  # KeyError for unknown scale, as the reference
    lo = int(np.floor(_freq_to_midi(max(fmin, 20.0)))) - 12
    hi = int(np.ceil(_freq_to_midi(min(fmax, 22050.0)))) + 12
    fr, wt = [], []
    for midi in range(lo, hi + 1):
        iv = (midi % 12 - root) % 12
        if iv not in intervals:
            continue
        f = 440.0 * (2.0 ** ((float(midi) - 69.0) / 12.0))  # :72-73
        if f < fmin * 0.5 or f > fmax * 2.0:
            continue
        role = ("root" if iv == 0 else "fifth" if iv == 7 else "third" if iv in (3, 4)
                else "seventh" if iv in (10, 11) else "other")  # :82-95
        fr.append(f)
        wt.append(ROLE_WEIGHTS[role])
    return np.array(fr, dtype=float), np.array(wt, dtype=float)


def target_bins_for_freqs(freqs: np.ndarray, key: str, scale: str) -> np.ndarray:
    """dsp/quantizer.py:127-196 (row-blocked so the bins x bins matrix stays small)."""
    freqs = np.asarray(freqs, dtype=float)
    if freqs.ndim != 1:
        raise ValueError("freqs must be 1D array")
    ident = np.arange(len(freqs), dtype=int)
    valid = np.isfinite(freqs) & (freqs > 0.0)
    if not np.any(valid):
        return ident
    nf, nw = scale_notes(key, scale, float(np.min(freqs[valid])), float(np.max(freqs[valid])))
    if nf.size == 0:
        return ident
    nw = np.clip(nw, 1e-3, None)
    cost = np.abs(freqs[:, None] - nf[None, :]) / nw[None, :]  # :171-174
    tf = nf[np.argmin(cost, axis=1)]  # first-index tie break (:178)
    tb = np.empty(len(freqs), dtype=int)
    for a in range(0, len(freqs), 512):  # :186-191
        tb[a:a + 512] = np.argmin(np.abs(tf[a:a + 512, None] - freqs[None, :]), axis=1)
    return np.where(valid, tb, ident)  # :194


def harmonic_target_bins(freqs: np.ndarray, fundamental_hz: float, n_harmonics: int = 32) -> np.ndarray:
    """dsp/quantizer.py:199-250."""
    freqs = np.asarray(freqs, dtype=float)
    if fundamental_hz <= 0.0:
        return np.arange(len(freqs), dtype=int)
    h = fundamental_hz * np.arange(1, n_harmonics + 1, dtype=float)
    fmax = float(np.max(freqs[freqs > 0])) if np.any(freqs > 0) else 24000.0
    h = h[h <= fmax * 1.1]
    if len(h) == 0:
        return np.arange(len(freqs), dtype=int)
    nh = h[np.argmin(np.abs(freqs[:, None] - h[None, :]), axis=1)]
    tb = np.argmin(np.abs(nh[:, None] - freqs[None, :]), axis=1)
    tb[0] = 0
    return tb


def quantize_band_mask(freqs: np.ndarray, min_hz: float, max_hz: float) -> np.ndarray:
    """dsp/pipeline.py:164-177."""
    freqs = np.asarray(freqs, dtype=float)
    m = np.ones_like(freqs, dtype=bool)
    if min_hz > 0.0:
        m &= freqs >= min_hz
    if max_hz > 0.0:
        m &= freqs <= max_hz
    if m.size:
        m[0] = False
    return m


# --------------------------------------------------------------------------- quantizer (all frames at once)
def smear_kernel(radius: int = 2) -> np.ndarray:
    """dsp/quantizer.py:460-465."""
    k = np.arange(2 * radius + 1, dtype=float) - radius
    sigma = max(1.0, radius / 2.0)
    g = np.exp(-0.5 * (k / sigma) ** 2)
    return g / np.sum(g)


def quantize_frames(mags: np.ndarray, phases: np.ndarray, target_bins: np.ndarray,
                    active_mask: Optional[np.ndarray], snap_strength: float, smear: float,
                    bin_smoothing: bool, smear_radius: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    """
    dsp/quantizer.py:343-529 applied to every frame (rows of mags/phases: [frames, bins]).

    The reference scatters with np.add.at in ascending source order (:446-451) and then
    smears source by source, tap by tap (:304-340).  Here the outer loops run over
    source bins and the arithmetic is vectorised over frames, so every destination
    receives its float64 addends in the reference's order (a source that is not
    "valid" in some frame contributes an exact 0.0 there).
    """
    mags = np.array(mags, dtype=float)
    phases = np.array(phases, dtype=float)
    single = mags.ndim == 1
    if single:
        mags, phases = mags[None, :], phases[None, :]
    snap = float(np.clip(snap_strength, 0.0, 1.0))  # :405
    smear = float(np.clip(smear, 0.0, 1.0))  # :406
    if snap <= 0.0 and not bin_smoothing:  # :409
        return (mags[0], phases[0]) if single else (mags, phases)
    T, n = mags.shape
    tb = np.asarray(target_bins)
    am = np.ones(n, dtype=bool) if active_mask is None else np.asarray(active_mask, dtype=bool)
    E = mags * snap  # :424
    valid = (E > 0.0) & ((tb >= 0) & (tb < n) & am)[None, :]  # :426-431
    new = mags - np.where(valid, E, 0.0)  # :434
    base = E * (1.0 - smear)  # :437
    sme = E * smear  # :438
    tE = np.zeros((T, n))
    tP = np.zeros((T, n), dtype=complex)
    src = np.nonzero(((tb >= 0) & (tb < n) & am))[0]
    ph = np.exp(1j * phases)  # :450
    for i in src:  # :446-451, ascending source index
        v = valid[:, i]
        b = np.where(v, base[:, i], 0.0)
        d = tb[i]
        new[:, d] += b
        tE[:, d] += b
        tP[:, d] += b * ph[:, i]
    if smear > 0.0 and smear_radius > 0:  # :458
        k = smear_kernel(smear_radius)
        for i in src:  # :304-340
            v = valid[:, i] & (sme[:, i] > 0.0)  # :468
            e = np.where(v, sme[:, i], 0.0)
            t = int(tb[i])
            a, b = max(0, t - smear_radius), min(n, t + smear_radius + 1)
            k0 = max(0, smear_radius - t)
            ksum = 0.0
            for q in range(k0, k0 + (b - a)):  # :321-323
                ksum += k[q]
            if b - a <= 0 or not ksum > 0.0:
                continue
            cs = np.cos(phases[:, i]) + 1j * np.sin(phases[:, i])  # :337-339
            for j in range(b - a):
                le = e * (k[k0 + j] / ksum)  # :330-331
                new[:, a + j] += le
                tE[:, a + j] += le
                tP[:, a + j] += le * cs
    pm = tE > 0.0  # :518
    out_ph = phases.copy()
    out_ph[pm] = np.angle(tP[pm])  # :520
    if bin_smoothing and n > 2:  # :523-527, scipy convolve1d mode="nearest"
        p = np.concatenate([new[:, :1], new, new[:, -1:]], axis=1)
        # scipy.ndimage.correlate1d accumulates centre first, then symmetric pairs;
        # tests pin this against the live reference output.
        new = 0.5 * p[:, 1:-1] + 0.25 * (p[:, :-2] + p[:, 2:])
    return (new[0], out_ph[0]) if single else (new, out_ph)


# --------------------------------------------------------------------------- spectral FX (one frame)
def fx_bitcrush(mag, phase, *, method="uniform", step=0.02, step_db=1.5, threshold=None):
    """dsp/spectral_fx.py:198-260."""
    mag = np.clip(np.asarray(mag, dtype=float), 0.0, None)
    if method == "uniform":
        out = mag.copy() if step <= 0 else np.clip(np.round(mag / step) * step, 0.0, None)
    elif method == "log":
        if step_db <= 0:
            out = mag.copy()
        else:
            db = 20.0 * np.log10(np.clip(mag, 1e-12, None))
            out = 10.0 ** ((np.round(db / step_db) * step_db) / 20.0)
    else:
        out = mag.copy()
    if threshold is not None and threshold > 0.0:
        out = np.where(out < threshold, 0.0, out)
    return out, np.asarray(phase, dtype=float)


def fx_phase_dispersal(mag, phase, *, thresh=0.01, amount=0.5, randomized=False, rand_amt=0.3):
    """dsp/spectral_fx.py:263-323 (draws np.random.rand(n) from the global state, :314)."""
    mag = np.asarray(mag, dtype=float).copy()
    phase = np.asarray(phase, dtype=float)
    if amount <= 0 and not randomized:
        return mag, phase
    mask = (mag > thresh) if (thresh is not None and thresh > 0.0) else np.ones_like(mag, dtype=bool)
    rot = np.where(mask, amount * (mag / (np.max(mag) + 1e-12)), 0.0)
    if randomized:
        jit = (np.random.rand(*phase.shape) * 2.0 - 1.0) * rand_amt
        rot = rot + np.where(mask, jit, 0.0)
    out = (phase + rot + np.pi) % (2.0 * np.pi) - np.pi
    return mag, out


def fx_bin_scramble(mag, phase, *, window=5, mode="random_pick"):
    """dsp/spectral_fx.py:326-390 (global np.random draws at :364 / :371)."""
    mag = np.asarray(mag, dtype=float)
    phase = np.asarray(phase, dtype=float)
    if window < 2 or window % 2 == 0:
        window = max(3, window if window % 2 == 1 else window + 1)
    if mode == "random_pick":
        half = window // 2
        n = mag.shape[0]
        idx = np.clip(np.arange(n) + np.random.randint(-half, half + 1, size=n), 0, n - 1)
        out = mag[idx]
    elif mode == "swap":
        out = mag.copy()
        s = np.where(np.random.rand(mag.size - 1) < 0.25)[0]
        if len(s) > 1:
            s = s[np.concatenate([[True], np.diff(s) > 1])]
        if len(s) > 0:
            out[s], out[s + 1] = out[s + 1].copy(), out[s].copy()
    else:
        out = mag.copy()
    out = out * (np.sum(mag) / (np.sum(out) + 1e-12))
    return out, phase


def apply_spectral_fx(mag, phase, mode, s, params):
    """dsp/pipeline.py:64-141: strength -> concrete FX parameters, one frame."""
    s = float(s)
    params = params or {}
    if not mode or s <= 0.0:
        return mag, phase
    if mode == "bitcrush":
        threshold = params.get("threshold", None)
        if threshold is None and s >= 0.4:
            threshold = 0.02 * (s ** 1.5) * float(mag.max() if mag.size else 1.0)
        return fx_bitcrush(mag, phase, method=params.get("method", "log"),
                           step=params.get("step", 0.01 + 0.09 * (s ** 1.2)),
                           step_db=params.get("step_db", 0.5 + 7.5 * (s ** 1.3)),
                           threshold=threshold)
    if mode == "phase_dispersal":
        randomized = params.get("randomized", s > 0.35)
        return fx_phase_dispersal(
            mag, phase,
            thresh=params.get("thresh", 0.01 * float(mag.max() if mag.size else 1.0)),
            amount=params.get("amount", (s ** 1.7) * np.pi),
            randomized=randomized,
            rand_amt=params.get("rand_amt", 0.0 if not randomized else 0.2 * (s ** 1.3) * np.pi))
    if mode == "bin_scramble":
        w = params.get("window", None)
        if w is None:
            w = int(3 + (12 * (s ** 1.2)))
        if w < 3:
            w = 3
        if w % 2 == 0:
            w += 1
        return fx_bin_scramble(mag, phase, window=w,
                               mode=params.get("mode", "swap" if s < 0.4 else "random_pick"))
    return mag, phase


def formant_shift_frame(mag, shift_semitones: float, lifter_order: int = 30):
    """dsp/spectral_fx.py:116-195: cepstral envelope (low-quefrency lifter) moved along the bin axis by
    2**(semitones/12), fine structure kept, total energy preserved."""
    mag = np.asarray(mag, dtype=float)
    if len(mag) < 4 or shift_semitones == 0.0:  # :146-147
        return mag.copy()
    eps = 1e-12
    n_bins = len(mag)
    log_mag = np.log(np.clip(mag, eps, None))  # :153
    cep = np.fft.irfft(log_mag, n=2 * (n_bins - 1))  # :156
    lifter = np.zeros_like(cep)  # :159-162
    order = min(lifter_order, len(lifter) // 2)
    lifter[:order] = 1.0
    lifter[-order + 1:] = 1.0
    log_env = np.fft.rfft(cep * lifter)[:n_bins].real  # :168
    log_fine = np.fft.rfft(cep * (1.0 - lifter))[:n_bins].real  # :169
    ratio = 2.0 ** (shift_semitones / 12.0)  # :173
    idx = np.arange(n_bins, dtype=float)
    env_s = np.interp(idx / ratio, idx, log_env, left=log_env[0], right=log_env[-1])  # :176-183
    out = np.exp(env_s + log_fine)  # :186-187
    e0, e1 = np.sum(mag ** 2), np.sum(out ** 2)  # :190-193
    if e1 > eps:
        out *= np.sqrt(e0 / e1)
    return out


# --------------------------------------------------------------------------- spectral stage on an STFT matrix
def spectral_quantize_stft(S, freqs, key, scale, snap_strength, smear, bin_smoothing, *,
                           is_high_band=False, spectral_fx_mode=None, spectral_fx_strength=0.0,
                           spectral_fx_params=None, spectral_freeze=False, harmonic_lock_hz=0.0,
                           quantize_min_hz=110.0, quantize_max_hz=5000.0, formant_shift=0.0):
    """dsp/pipeline.py:228-342."""
    mags = np.abs(S).T.copy()  # [frames, bins]   :269
    phases = np.angle(S).T.copy()  # :270
    tb = (harmonic_target_bins(freqs, harmonic_lock_hz) if harmonic_lock_hz > 0.0
          else target_bins_for_freqs(freqs, key, scale))  # :277-280
    am = quantize_band_mask(freqs, quantize_min_hz, quantize_max_hz)  # :282
    if spectral_freeze and mags.shape[0] > 0:  # :285-287, :303-304
        mags[:] = mags[0][None, :]
    if formant_shift != 0.0:  # :306-310, any band, after the freeze and before the FX
        for t in range(mags.shape[0]):
            mags[t] = formant_shift_frame(mags[t], formant_shift)
    if is_high_band and spectral_fx_mode is not None:  # :291, :313-314 -- per frame, RNG order preserved
        for t in range(mags.shape[0]):
            mags[t], phases[t] = apply_spectral_fx(mags[t], phases[t], spectral_fx_mode,
                                                   spectral_fx_strength, spectral_fx_params)
    mags, phases = quantize_frames(mags, phases, tb, am, snap_strength, smear, bin_smoothing)  # :316-327
    return (mags * np.exp(1j * phases)).T  # :341


# --------------------------------------------------------------------------- time-domain stages
def wavefold(x, fold_amount=1.0, bias=0.0, threshold=1.0):
    """dsp/distortion.py:18-58."""
    y = (np.asarray(x, dtype=float) + bias) * fold_amount
    if threshold <= 0.0:
        threshold = 1.0
    y = y.copy()
    p = y > threshold
    y[p] = 2.0 * threshold - y[p]
    q = y < -threshold
    y[q] = -2.0 * threshold - y[q]
    return np.clip(y, -threshold, threshold).astype(np.float32)


def soft_tube(x, drive=1.0, warmth=0.5):
    """dsp/distortion.py:61-90."""
    y = np.asarray(x, dtype=float) * max(drive, 0.0)
    a = 1.0 + 4.0 * float(np.clip(warmth, 0.0, 1.0))
    y = np.tanh(a * y)
    y /= np.tanh(a) if a != 0.0 else 1.0
    return y.astype(np.float32)


def apply_distortion(x, mode, fold_amount=1.0, bias=0.0, drive=1.0, warmth=0.5):
    """dsp/distortion.py:93-114."""
    if np.asarray(x).ndim != 1:
        raise ValueError("Distortion currently expects mono (1D) audio")
    if mode == "wavefold":
        return wavefold(x, fold_amount=fold_amount, bias=bias, threshold=1.0)
    if mode == "tube":
        return soft_tube(x, drive=drive, warmth=warmth)
    raise ValueError(f"Unsupported distortion mode: {mode}")


def limiter_constants(sr: int, ceiling_db: float, lookahead_ms: float, release_ms: float):
    """dsp/limiter.py:52-60 (Python round() = half-to-even)."""
    ceiling = 10.0 ** (ceiling_db / 20.0)
    L = int(max(1, round(sr * (lookahead_ms / 1000.0))))
    R = max(1, int(round(sr * (release_ms / 1000.0))))
    return ceiling, L, float(np.exp(-1.0 / R))


def peak_limiter(x, sr, ceiling_db=-1.0, lookahead_ms=5.0, release_ms=50.0):
    """dsp/limiter.py:14-80 (per-sample loop in oracle/qd_seq.c)."""
    x = np.ascontiguousarray(x, dtype=float)
    if x.ndim != 1:
        raise ValueError("peak_limiter currently expects mono (1D) audio")
    n = x.shape[0]
    if n == 0:
        return x.astype(np.float32), np.ones_like(x, dtype=np.float32)
    ceiling, L, c = limiter_constants(sr, ceiling_db, lookahead_ms, release_ms)
    gain = np.empty(n)
    _seq_lib().qd_oracle_limiter_gain(_dptr(x), n, L, ceiling, c, _dptr(gain))
    return (x * gain).astype(np.float32), gain.astype(np.float32)


def sosfilt(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """scipy.signal.sosfilt arithmetic (DF2T cascade, float64), see oracle/qd_seq.c."""
    sos = np.ascontiguousarray(sos, dtype=float)
    x = np.ascontiguousarray(x, dtype=float)
    y = np.empty_like(x)
    _seq_lib().qd_oracle_sosfilt(_dptr(sos), int(sos.shape[0]), _dptr(x), x.shape[0], _dptr(y))
    return y


def linkwitz_riley_sos(sr: int, crossover_hz: float, order_per_side: int = 2):
    """dsp/crossover.py:9-68."""
    nyq = sr / 2.0
    wn = crossover_hz / nyq
    if wn <= 0.0 or wn >= 1.0:
        raise ValueError(f"Crossover frequency {crossover_hz} Hz must be between 0 and Nyquist ({nyq} Hz)")
    lp = butter(N=order_per_side, Wn=wn, btype="low", output="sos")
    hp = butter(N=order_per_side, Wn=wn, btype="high", output="sos")
    return np.concatenate([lp, lp], axis=0), np.concatenate([hp, hp], axis=0)


def linkwitz_riley_split(x, sr, crossover_hz):
    """dsp/crossover.py:71-118 (mono)."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim != 1:
        raise ValueError("oracle handles mono only")
    lo, hi = linkwitz_riley_sos(sr, crossover_hz)
    return sosfilt(lo, x).astype(np.float32), sosfilt(hi, x).astype(np.float32)


def saturate_lowband(x, drive=1.0):
    """dsp/saturation.py:6-62 (tanh evaluated in float32, divided by float64 tanh(3))."""
    x = np.asarray(x, dtype=np.float32)
    xf = x * max(float(drive), 0.0)
    y = np.tanh(3.0 * xf)
    y = y / np.tanh(3.0)
    return y.astype(np.float32)


def null_test_db(a, b) -> float:
    """tests/utils/audio_test_utils.py:34-96 (1-D case): RMS of (a-b) in dB, floor 1e-10."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    n = min(len(a), len(b))
    r = a[:n] - b[:n]
    return float(20.0 * np.log10(max(float(np.sqrt(np.mean(r ** 2))) if n else 0.0, 1e-10)))


# --------------------------------------------------------------------------- QA metric
def _nearest_scale_midi(midi_value: float, key: str, scale: str) -> float:
    """dsp/analyses.py:18-50."""
    if not np.isfinite(midi_value):
        return np.nan
    freq = 440.0 * (2.0 ** ((midi_value - 69.0) / 12.0))
    fmin, fmax = float(max(20.0, freq / 4.0)), float(min(20000.0, freq * 4.0))
    root = note_name_to_pitch_class(key)
    intervals = SCALE_INTERVALS[scale]
    lo = int(np.floor(_freq_to_midi(max(fmin, 20.0)))) - 12
    hi = int(np.ceil(_freq_to_midi(min(fmax, 22050.0)))) + 12
    midis = []
    for midi in range(lo, hi + 1):  # dsp/quantizer.py:113-123
        if (midi % 12 - root) % 12 in intervals:
            f = 440.0 * (2.0 ** ((float(midi) - 69.0) / 12.0))
            if f < fmin * 0.5 or f > fmax * 2.0:
                continue
            midis.append(midi)
    if not midis:
        return np.nan
    m = np.array(midis, dtype=float)
    return float(m[int(np.argmin(np.abs(m - midi_value)))])


def avg_cents_offset_from_scale(audio, sr, key, scale, frame_length=2048, hop_length=None, topn_peaks=3,
                                min_db=-60.0, return_bins=False):
    """dsp/analyses.py:53-142.  ``return_bins`` additionally returns the chosen bins [frames, topn] (-1 = none),
    which is what the CUDA kernel is compared with."""
    x = np.asarray(audio, dtype=float)
    if x.ndim != 1:
        raise ValueError("avg_cents_offset_from_scale expects mono (1D) audio")
    S, freqs = stft(x, sr, n_fft=frame_length)  # :90-96 (hop_length is unused by the reference)
    mags_db = 20.0 * np.log10(np.maximum(np.abs(S), 1e-12))  # :97-101
    per_peak = []
    bins = np.full((mags_db.shape[1], topn_peaks), -1, dtype=np.int16)
    for t in range(mags_db.shape[1]):
        frame_db = mags_db[:, t]
        if np.max(frame_db) < min_db:  # :108-110
            continue
        count = 0
        for idx in np.argsort(frame_db)[::-1]:  # :113
            if frame_db[idx] < min_db:
                break
            freq = float(freqs[idx])
            if freq <= 0.0:
                continue
            midi_est = _freq_to_midi(freq)
            if not np.isfinite(midi_est):
                continue
            scale_midi = _nearest_scale_midi(midi_est, key, scale)
            if not np.isfinite(scale_midi):
                continue
            per_peak.append(abs(float(100.0 * (midi_est - scale_midi))))
            bins[t, count] = idx
            count += 1
            if count >= topn_peaks:
                break
    arr = np.array(per_peak, dtype=float)
    avg = float(np.mean(arr)) if arr.size else float("nan")
    return (avg, arr, bins) if return_bins else (avg, arr)


# --------------------------------------------------------------------------- pipeline
def _fit(x: np.ndarray, n: int) -> np.ndarray:
    if x.shape[0] == n:
        return x
    return x[:n] if x.shape[0] > n else np.concatenate([x, np.zeros(n - x.shape[0], dtype=np.float32)])


def process_single_band(x_in, sr, *, key, scale, snap_strength, smear, bin_smoothing, pre_quant,
                        post_quant, distortion_mode, distortion_params, limiter_on, limiter_ceiling_db,
                        dry_wet, tap_input, passthrough_test=False, is_high_band=False,
                        spectral_fx_mode=None, spectral_fx_strength=0.0, spectral_fx_params=None,
                        spectral_freeze=False, harmonic_lock_hz=0.0, output_trim_db=0.0,
                        sub_cut_hz=110.0, air_cut_hz=5000.0, n_fft=N_FFT_DEFAULT, formant_shift=0.0,
                        no_spectral=False):
    """dsp/pipeline.py:419-920.  Of the autotune branch (:537-601) only the form with its pitch stage gated off
    (``no_spectral``: x_pre = x_in, distortion, limiter, mix) is restated here; the pitch stage itself is
    oracle/qd_autotune.py."""
    n = x_in.shape[0]
    if passthrough_test:  # :477-535
        S, _ = stft(x_in, sr, n_fft)
        y = _fit(istft(S, sr, n_fft, length=n).astype(np.float32), n)
        return y, {"pre_quant": x_in.copy(), "post_dist": y.copy(), "output": y.copy()}

    def spec(S, freqs):
        return spectral_quantize_stft(
            S, freqs, key, scale, snap_strength, smear, bin_smoothing, is_high_band=is_high_band,
            spectral_fx_mode=spectral_fx_mode, spectral_fx_strength=spectral_fx_strength,
            spectral_fx_params=spectral_fx_params, spectral_freeze=spectral_freeze,
            harmonic_lock_hz=harmonic_lock_hz, quantize_min_hz=sub_cut_hz, quantize_max_hz=air_cut_hz,
            formant_shift=formant_shift)

    pre = bool(pre_quant and snap_strength > 0.0) and not no_spectral  # :635
    post = bool(post_quant and snap_strength > 0.0) and not no_spectral  # :728
    S, freqs = (None, None) if no_spectral else stft(x_in, sr, n_fft)  # :615
    if pre:
        S = spec(S, freqs)
        x_pre = istft(S, sr, n_fft, length=n).astype(np.float32)  # :668 / :690
    else:
        x_pre = x_in.copy()  # :700
    tap_pre = x_pre.copy()
    dp = distortion_params or {}
    x_dist = apply_distortion(x_pre, distortion_mode or "wavefold",
                              fold_amount=float(dp.get("fold_amount", 1.0)), bias=float(dp.get("bias", 0.0)),
                              drive=float(dp.get("drive", 1.0)), warmth=float(dp.get("warmth", 0.5)))  # :705-719
    tap_dist = x_dist.copy()
    if post:
        if pre:  # :729-801
            S2, f2 = stft(x_dist, sr, n_fft)
            x_pq = istft(spec(S2, f2), sr, n_fft, length=n).astype(np.float32)
        else:  # :802-848 -- post-quant of the UNDISTORTED spectrum
            x_pq = istft(spec(S, freqs), sr, n_fft, length=n).astype(np.float32)
    else:
        x_pq = x_dist.copy() if (pre or no_spectral) else istft(S, sr, n_fft, length=n).astype(np.float32)  # :849-879, :591
    if limiter_on:  # :882-891
        x_lim, _ = peak_limiter(x_pq, sr, ceiling_db=limiter_ceiling_db, lookahead_ms=5.0, release_ms=30.0)
    else:
        x_lim = x_pq.copy()
    dw = float(np.clip(dry_wet, 0.0, 1.0))  # :894
    y = (dw * x_lim) + ((1.0 - dw) * tap_input)  # :895 float32 arithmetic (weak Python scalars)
    if output_trim_db != 0.0:  # :898-900
        y = y * (10.0 ** (output_trim_db / 20.0))
    y = _fit(y.astype(np.float32), n)  # :902-910
    return y, {"pre_quant": tap_pre, "post_dist": tap_dist, "output": y.copy()}


def process_audio(audio, sr=48000, key="D", scale="minor", quantize_mode="spectral_bins",
                  snap_strength=1.0, smear=0.1, bin_smoothing=True, pre_quant=True, post_quant=True,
                  distortion_mode="wavefold", distortion_params=None, limiter_on=True,
                  limiter_ceiling_db=-1.0, dry_wet=1.0, use_multiband=False, crossover_hz=300.0,
                  lowband_drive=1.0, passthrough_test=False, spectral_fx_mode=None,
                  spectral_fx_strength=0.0, spectral_fx_params=None, spectral_freeze=False,
                  harmonic_lock_hz=0.0, delta_listen=False, mono_strength=1.0, output_trim_db=0.0,
                  sub_cut_hz=110.0, air_cut_hz=5000.0, low_trim_db=0.0,
                  n_fft=N_FFT_DEFAULT, formant_shift=0.0) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    """
    dsp/pipeline.py:1113-1407 for the STFT path.  ``quantize_mode`` must resolve to
    "spectral_bins" (the reference's default "autotune_v1" is a different, out-of-scope
    algorithm; it flips to spectral_bins by itself when an FX/freeze/lock option is set,
    :1315-1324).  ``n_fft`` restates the module global N_FFT_DEFAULT (:149).
    """
    if quantize_mode == "autotune_v1" and (spectral_fx_mode is not None or spectral_freeze
                                            or formant_shift != 0.0 or harmonic_lock_hz > 0.0):
        quantize_mode = "spectral_bins"
    if quantize_mode == "autotune_v1" and snap_strength > 0.0:
        use_multiband = False  # :1326-1327, ahead of every branch (a passthrough_test render is single band then too)
    # autotune_v1 survives inside a multiband render only with snap_strength <= 0 (:1326-1327), where its pitch stage
    # is gated off (:538): the high band is distorted, limited and mixed without any STFT
    no_spectral = quantize_mode == "autotune_v1" and use_multiband and not snap_strength > 0.0 and not passthrough_test
    # passthrough_test (:477) returns the STFT -> iSTFT round trip before the mode is looked at
    if quantize_mode != "spectral_bins" and not no_spectral and not passthrough_test:
        raise NotImplementedError("this oracle covers quantize_mode='spectral_bins'; autotune_v1 is oracle/qd_autotune.py")
    x = np.asarray(audio, dtype=np.float32)  # config.py:19-24
    if x.ndim == 2:
        x = x.mean(axis=1).astype(np.float32)
    tap_in = x.copy()
    n = x.shape[0]
    common: Dict[str, Any] = dict(
        key=key, scale=scale, snap_strength=snap_strength, smear=smear, bin_smoothing=bin_smoothing,
        pre_quant=pre_quant, post_quant=post_quant, distortion_mode=distortion_mode,
        distortion_params=distortion_params or {}, limiter_on=limiter_on,
        limiter_ceiling_db=limiter_ceiling_db, dry_wet=dry_wet, passthrough_test=passthrough_test,
        spectral_fx_mode=spectral_fx_mode, spectral_fx_strength=spectral_fx_strength,
        spectral_fx_params=spectral_fx_params or {}, spectral_freeze=spectral_freeze,
        harmonic_lock_hz=harmonic_lock_hz, output_trim_db=output_trim_db, sub_cut_hz=sub_cut_hz,
        air_cut_hz=air_cut_hz, n_fft=n_fft, formant_shift=formant_shift, no_spectral=no_spectral)
    if use_multiband:  # dsp/pipeline.py:1011-1110
        low, high = linkwitz_riley_split(x, sr, crossover_hz)
        d = n_fft // 2  # :1056, filter-delay term cancels (:380-386)
        low = np.concatenate([np.zeros(d, dtype=low.dtype), low])[: high.shape[0]] if d > 0 else low
        low_p = saturate_lowband(low, drive=lowband_drive)
        # mono-maker is the identity for mono input (dsp/saturation.py:88-90); the blend at
        # :1066-1068 then yields mono_strength*l + (1-mono_strength)*l, evaluated in float32.
        if 0.0 < mono_strength < 1.0:
            low_p = (mono_strength * low_p) + ((1.0 - mono_strength) * low_p)
        if low_trim_db != 0.0:
            low_p = low_p * (10.0 ** (low_trim_db / 20.0))
        low_p = low_p.astype(np.float32)
        yh, th = process_single_band(high, sr, tap_input=high, is_high_band=True, **common)
        y = _fit((low_p + yh).astype(np.float32), n)
        taps = {"input": tap_in, "pre_quant": th["pre_quant"], "post_dist": low_p + th["post_dist"],
                "output": y.copy()}
    else:
        y, tb = process_single_band(x, sr, tap_input=tap_in, is_high_band=False, **common)
        taps = {"input": tap_in, **tb}
    if delta_listen:  # :1371-1375
        m = min(len(tap_in), len(y))
        y = (tap_in[:m] - y[:m]).astype(np.float32)
        taps["output"] = y.copy()
    return y, taps
