"""Host-side table and parameter builder (NumPy float64).

Everything integer the reference derives on the CPU -- target bins, band mask, sample counts
with Python's half-to-even ``round``, crossover coefficients -- is computed here with the
reference's own formulas so that it is bit-exact, then handed to the CUDA library as plain
numbers (include/qd_b200.h).  ``file:line`` citations are relative to
``/root/reference/quantum_distortion``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np
from scipy.signal import butter

from . import _lib
from .config import N_FFT_DEFAULT

NOTE_NAMES_SHARP = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]
SCALE_INTERVALS: Dict[str, Tuple[int, ...]] = {
    "major": (0, 2, 4, 5, 7, 9, 11),
    "minor": (0, 2, 3, 5, 7, 8, 10),
    "pentatonic": (0, 2, 4, 7, 9),
    "dorian": (0, 2, 3, 5, 7, 9, 10),
    "mixolydian": (0, 2, 4, 5, 7, 9, 10),
    "harmonic_minor": (0, 2, 3, 5, 7, 8, 11),
}
HARMONIC_WEIGHTS = {"root": 1.0, "fifth": 0.8, "third": 0.7, "seventh": 0.6, "other": 0.5}
SUPPORTED_N_FFT = (512, 1024, 2048, 4096, 8192)
_FLAT_TO_SHARP = {"DB": "C#", "EB": "D#", "GB": "F#", "AB": "G#", "BB": "A#"}


# ----------------------------------------------------------------------------- scale tables
def note_name_to_pitch_class(name: str) -> int:
    """dsp/quantizer.py:59-69; ValueError for an unsupported key."""
    n = name.strip().upper()
    for flat, sharp in _FLAT_TO_SHARP.items():
        n = n.replace(flat, sharp)
    if n not in NOTE_NAMES_SHARP:
        raise ValueError(f"Unsupported key name: {n}")
    return NOTE_NAMES_SHARP.index(n)


def _role_weight(interval: int) -> float:
    """dsp/quantizer.py:82-95 role classification -> attraction weight (:42-48)."""
    if interval == 0:
        return HARMONIC_WEIGHTS["root"]
    if interval == 7:
        return HARMONIC_WEIGHTS["fifth"]
    if interval in (3, 4):
        return HARMONIC_WEIGHTS["third"]
    if interval in (10, 11):
        return HARMONIC_WEIGHTS["seventh"]
    return HARMONIC_WEIGHTS["other"]


def build_scale_notes(key: str, scale: str, min_freq: float, max_freq: float) -> Tuple[np.ndarray, np.ndarray]:
    """dsp/quantizer.py:98-124 -> (frequencies, weights) of the in-scale notes around the band."""
    root_pc = note_name_to_pitch_class(key)
    intervals = SCALE_INTERVALS[scale]  # KeyError for an unknown scale, like the reference

    def midi_of(f: float) -> float:
        return 69.0 + 12.0 * np.log2(f / 440.0)

    first = int(np.floor(midi_of(max(min_freq, 20.0)))) - 12
    last = int(np.ceil(midi_of(min(max_freq, 22050.0)))) + 12
    freqs, weights = [], []
    for midi in range(first, last + 1):
        interval = ((midi % 12) - root_pc) % 12
        if interval in intervals:
            f = 440.0 * (2.0 ** ((float(midi) - 69.0) / 12.0))
            if min_freq * 0.5 <= f <= max_freq * 2.0:
                freqs.append(f)
                weights.append(_role_weight(interval))
    return np.asarray(freqs, dtype=np.float64), np.asarray(weights, dtype=np.float64)


def build_target_bins_for_freqs(freqs: np.ndarray, key: str, scale: str) -> np.ndarray:
    """dsp/quantizer.py:127-196: per bin, nearest weighted scale note, then nearest bin to that note.
    ``np.argmin`` keeps the reference's first-index tie break."""
    freqs = np.asarray(freqs, dtype=np.float64)
    if freqs.ndim != 1:
        raise ValueError("freqs must be 1D array")
    identity = np.arange(freqs.shape[0], dtype=np.int64)
    ok = np.isfinite(freqs) & (freqs > 0.0)
    if not ok.any():
        return identity
    note_f, note_w = build_scale_notes(key, scale, float(freqs[ok].min()), float(freqs[ok].max()))
    if note_f.size == 0:
        return identity
    note_w = np.clip(note_w, 1e-3, None)
    nearest_note = np.argmin(np.abs(freqs[:, None] - note_f[None, :]) / note_w[None, :], axis=1)
    wanted = note_f[nearest_note]
    out = np.empty(freqs.shape[0], dtype=np.int64)
    step = 256  # row blocks keep the bins x bins distance matrix small at n_fft = 8192
    for lo in range(0, freqs.shape[0], step):
        out[lo:lo + step] = np.argmin(np.abs(wanted[lo:lo + step, None] - freqs[None, :]), axis=1)
    return np.where(ok, out, identity)


def build_harmonic_target_bins(freqs: np.ndarray, fundamental_hz: float, n_harmonics: int = 32) -> np.ndarray:
    """dsp/quantizer.py:199-250 (harmonic lock)."""
    freqs = np.asarray(freqs, dtype=np.float64)
    identity = np.arange(freqs.shape[0], dtype=np.int64)
    if fundamental_hz <= 0.0:
        return identity
    series = fundamental_hz * np.arange(1, n_harmonics + 1, dtype=np.float64)
    top = float(freqs[freqs > 0].max()) if (freqs > 0).any() else 24000.0
    series = series[series <= top * 1.1]
    if series.size == 0:
        return identity
    wanted = series[np.argmin(np.abs(freqs[:, None] - series[None, :]), axis=1)]
    out = np.argmin(np.abs(wanted[:, None] - freqs[None, :]), axis=1).astype(np.int64)
    out[0] = 0
    return out


def build_quantize_band_mask(freqs: np.ndarray, min_hz: float, max_hz: float) -> np.ndarray:
    """dsp/pipeline.py:164-177."""
    freqs = np.asarray(freqs, dtype=np.float64)
    mask = np.ones(freqs.shape, dtype=bool)
    if min_hz > 0.0:
        mask &= freqs >= min_hz
    if max_hz > 0.0:
        mask &= freqs <= max_hz
    if mask.size:
        mask[0] = False
    return mask


def smear_kernel(radius: int = 2) -> np.ndarray:
    """dsp/quantizer.py:460-465."""
    idx = np.arange(2 * radius + 1, dtype=np.float64) - radius
    sigma = max(1.0, radius / 2.0)
    k = np.exp(-0.5 * (idx / sigma) ** 2)
    return k / np.sum(k)


# ----------------------------------------------------------------------------- time-domain constants
def limiter_constants(sr: int, ceiling_db: float, lookahead_ms: float = 5.0, release_ms: float = 30.0):
    """dsp/limiter.py:52-60 with the pipeline's 5 ms / 30 ms (dsp/pipeline.py:883-889)."""
    ceiling = 10.0 ** (ceiling_db / 20.0)
    lookahead = int(max(1, round(sr * (lookahead_ms / 1000.0))))
    release = max(1, int(round(sr * (release_ms / 1000.0))))
    return ceiling, lookahead, float(np.exp(-1.0 / release))


def design_linkwitz_riley_sos(sr: int, crossover_hz: float, order_per_side: int = 2):
    """dsp/crossover.py:9-68."""
    nyquist = sr / 2.0
    wn = crossover_hz / nyquist
    if wn <= 0.0 or wn >= 1.0:
        raise ValueError(f"Crossover frequency {crossover_hz} Hz must be between 0 and Nyquist ({nyquist} Hz)")
    lp = butter(N=order_per_side, Wn=wn, btype="low", output="sos")
    hp = butter(N=order_per_side, Wn=wn, btype="high", output="sos")
    return np.concatenate([lp, lp], axis=0), np.concatenate([hp, hp], axis=0)


# ----------------------------------------------------------------------------- spectral FX (high band)
def resolve_spectral_fx(mode: Optional[str], strength: float, params: Optional[Dict[str, Any]]) -> Dict[str, Any]:
    """dsp/pipeline.py:64-141: strength -> concrete FX parameters.  Returns the qd_params fx_* values plus
    ``rng``: None, ("jitter",), ("pick", half) or ("swap",) -- the np.random draws one frame consumes."""
    s = float(strength)
    params = params or {}
    none = {"fx_mode": 0, "a": 0.0, "b": 0.0, "c": 0.0, "rng": None}
    if not mode or s <= 0.0:
        return none
    if mode == "bitcrush":
        method = params.get("method", "log")
        step_db = params.get("step_db", 0.5 + 7.5 * (s ** 1.3))
        step = params.get("step", 0.01 + 0.09 * (s ** 1.2))
        threshold = params.get("threshold", None)
        rel = 0.0
        absolute = 0.0
        if threshold is None and s >= 0.4:
            rel = 0.02 * (s ** 1.5)            # x frame max (:90)
        elif threshold is not None and threshold > 0.0:
            absolute = float(threshold)
        if method == "log":
            return {"fx_mode": _lib.QD_FX["bitcrush_log"], "a": float(step_db), "b": rel, "c": absolute, "rng": None}
        if method == "uniform":
            return {"fx_mode": _lib.QD_FX["bitcrush_uniform"], "a": float(step), "b": rel, "c": absolute, "rng": None}
        # unknown method: magnitudes unchanged, threshold still applies (dsp/spectral_fx.py:252-258)
        return {"fx_mode": _lib.QD_FX["bitcrush_uniform"], "a": 0.0, "b": rel, "c": absolute, "rng": None}
    if mode == "phase_dispersal":
        amount = params.get("amount", (s ** 1.7) * np.pi)
        thresh = params.get("thresh", None)      # default 0.01 * frame max (:105)
        randomized = bool(params.get("randomized", s > 0.35))
        rand_amt = params.get("rand_amt", 0.0 if not randomized else 0.2 * (s ** 1.3) * np.pi)
        if amount <= 0 and not randomized:       # dsp/spectral_fx.py:297-298
            return none
        c = -1.0 if thresh is None else max(float(thresh), 0.0)
        return {"fx_mode": _lib.QD_FX["phase_dispersal"], "a": float(amount), "b": float(rand_amt) if randomized else 0.0,
                "c": c, "rng": ("jitter",) if randomized else None}
    if mode == "bin_scramble":
        window = params.get("window", None)
        if window is None:
            window = int(3 + (12 * (s ** 1.2)))
        if window < 3:
            window = 3
        if window % 2 == 0:
            window += 1
        mode_name = params.get("mode", "swap" if s < 0.4 else "random_pick")
        if mode_name == "random_pick":
            return {"fx_mode": _lib.QD_FX["scramble_pick"], "a": float(window // 2), "b": 0.0, "c": 0.0,
                    "rng": ("pick", window // 2)}
        if mode_name == "swap":
            return {"fx_mode": _lib.QD_FX["scramble_swap"], "a": 0.0, "b": 0.0, "c": 0.0, "rng": ("swap",)}
        # unknown scramble mode: copy + energy rescale = identity (dsp/spectral_fx.py:382-388)
        return none
    return none  # unknown FX mode is a silent no-op (:140-141)


def replay_fx_table(rng_kind, n_passes: int, n_frames: int, n_bins: int) -> np.ndarray:
    """Replay, from the GLOBAL np.random state, the draws one reference ``process_audio`` call makes:
    pass 1 frames 0..T-1 then pass 2, one call per frame.  Returns
    ``[2, n_frames, n_bins]`` (int16 source bins, or float32 jitter in [-1, 1)); unused passes stay identity/zero."""
    kind = rng_kind[0]
    if kind == "jitter":
        out = np.zeros((2, n_frames, n_bins), dtype=np.float32)
    else:
        out = np.empty((2, n_frames, n_bins), dtype=np.int16)
        out[:] = np.arange(n_bins, dtype=np.int16)[None, None, :]
    base = np.arange(n_bins)
    for p in range(n_passes):
        for t in range(n_frames):
            if kind == "jitter":      # dsp/spectral_fx.py:314
                out[p, t] = (np.random.rand(n_bins) * 2.0 - 1.0).astype(np.float32)
            elif kind == "pick":      # :360-367
                half = rng_kind[1]
                out[p, t] = np.clip(base + np.random.randint(-half, half + 1, size=n_bins), 0, n_bins - 1)
            else:                     # swap, :368-381
                sw = np.where(np.random.rand(n_bins - 1) < 0.25)[0]
                if len(sw) > 1:
                    sw = sw[np.concatenate([[True], np.diff(sw) > 1])]
                idx = base.copy()
                idx[sw], idx[sw + 1] = sw + 1, sw
                out[p, t] = idx
    return out


# ----------------------------------------------------------------------------- precision of the spectral pass
F64_FAN_IN = 64   # more sources than this per target: their phasor sum can cancel below float32 resolution


def max_fan_in(target_bins: Optional[np.ndarray], active_mask: Optional[np.ndarray]) -> int:
    if target_bins is None or active_mask is None or not np.any(active_mask):
        return 0
    return int(np.max(np.bincount(np.asarray(target_bins)[np.asarray(active_mask, dtype=bool)])))


def choose_precision(precision: str, n_fft: int, target_bins, active_mask, fx_active: bool = False,
                     formant_active: bool = False, fx_mode: int = 0, freeze_active: bool = False) -> int:
    """0 = float32 kernels, 1 = float64 kernels.  "auto" keeps the fast float32 path for the reference's
    defaults and switches to float64 where float32 cannot hold the 1e-4 parity bound: n_fft 8192
     and quantiser tables whose targets gather more than F64_FAN_IN source bins
    (e.g. sub_cut_hz = air_cut_hz = 0), where the phase of a near-cancelling phasor sum is decided below
    float32 resolution; bitcrush on a linear grid, the formant shift and the spectral freeze (reasons below)."""
    if precision in ("float32", "f32", "fp32"):
        return 0
    if precision in ("float64", "f64", "fp64"):
        return 1
    if precision != "auto":
        raise ValueError("precision must be 'auto', 'float32' or 'float64'")
    if n_fft >= 8192:
        return 1   # every variant of the pass: the float32 FFT alone is 1e-4 ... 3e-4 off there
    if fx_mode == _lib.QD_FX["bitcrush_uniform"]:
        # round(m / step) on a LINEAR grid flips wherever float32 rounding moves a magnitude across a boundary, and a
        # flip moves the bin by a whole step whatever its size (measured: 1.3e-4 at step 0.004, 4e-6 in float64); the
        # default log-domain method is not affected (a flip there is proportional to the bin)
        return 1
    if formant_active:
        # the cepstral envelope takes log(max(m, 1e-12)) of EVERY bin: bins under the float32 FFT's noise floor (1e-7 of
        # the frame's strongest bin -- the far skirts of any DC-heavy or tonal frame) come out as noise, the envelope
        # follows, and the render moves by 2e-3 (measured on the fade-in of a wavefolded clip); float64 has 1e-16
        return 1
    if freeze_active:
        # every frame gets the magnitudes of frame 0 (half of it centre padding: a broadband spectrum) on its OWN phases;
        # on a tonal clip most bins of the later frames sit under the float32 FFT's noise floor, their float32 phases are
        # noise, and the frozen magnitudes make that noise audible: 4e-4 after one pass for 1e-7 of spectral noise on the
        # reference's own test signal (tests/test_oracle_ref_scenarios.py), 2e-3 measured on the GPU; float64: 2e-9
        return 1
    return 1 if max_fan_in(target_bins, active_mask) > F64_FAN_IN else 0


# ----------------------------------------------------------------------------- resolved render
@dataclass
class Resolved:
    """C structs plus the NumPy arrays that back their pointers (kept alive here)."""
    params: _lib.QdParams
    tables: Optional[_lib.QdTables]
    keepalive: tuple
    target_bins: Optional[np.ndarray]
    active_mask: Optional[np.ndarray]
    fx_rng: Optional[tuple] = None       # np.random draws per frame (see replay_fx_table)
    fx_passes: int = 0                   # quantised passes that consume them
    n_frames: int = 0
    n_bins: int = 0


def _sos_fill(dst, sos: np.ndarray) -> None:
    if sos.shape != (2, 6):
        raise ValueError("the CUDA crossover expects two cascaded second-order sections per band")
    for r in range(2):
        for c in range(6):
            dst[r][c] = float(sos[r, c])


def resolve(*, sr: int, n_samples: int, n_fft: int = N_FFT_DEFAULT, key: str, scale: str, snap_strength: float,
            smear: float, bin_smoothing: bool, pre_quant: bool, post_quant: bool, distortion_mode: Optional[str],
            distortion_params: Optional[Dict[str, Any]], limiter_on: bool, limiter_ceiling_db: float,
            dry_wet: float, use_multiband: bool, crossover_hz: float, lowband_drive: float,
            passthrough_test: bool, harmonic_lock_hz: float, delta_listen: bool, mono_strength: float,
            output_trim_db: float, low_trim_db: float, sub_cut_hz: float, air_cut_hz: float,
            spectral_fx_mode: Optional[str] = None, spectral_fx_strength: float = 0.0,
            spectral_fx_params: Optional[Dict[str, Any]] = None, precision: str = "auto",
            spectral_freeze: bool = False, formant_shift: float = 0.0, no_spectral: bool = False) -> Resolved:
    """Turn the reference's keyword arguments into qd_params / qd_tables.  Raises the reference's
    exceptions (SURVEY.md section 8(b) "Errors") before anything is launched."""
    if n_fft not in SUPPORTED_N_FFT:
        raise ValueError(f"n_fft must be one of {SUPPORTED_N_FFT}")
    p = _lib.QdParams()
    p.struct_size = C.sizeof(_lib.QdParams)
    p.sample_rate, p.n_fft, p.n_samples = int(sr), int(n_fft), int(n_samples)
    p.passthrough = int(bool(passthrough_test))
    p.pre_quant = int(bool(pre_quant and snap_strength > 0.0))    # dsp/pipeline.py:635 (unclipped gate)
    p.post_quant = int(bool(post_quant and snap_strength > 0.0))  # :728
    if no_spectral and not passthrough_test:   # autotune_v1 with its pitch stage gated off: no STFT at all (:537-601)
        p.no_spectral, p.pre_quant, p.post_quant = 1, 0, 0
    p.bin_smoothing = int(bool(bin_smoothing))
    # distortion (dsp/pipeline.py:705-710, dsp/distortion.py:93-114)
    mode = distortion_mode or "wavefold"
    dp = distortion_params or {}
    if mode == "wavefold":
        p.distortion_mode = 0
    elif mode == "tube":
        p.distortion_mode = 1
    elif not passthrough_test:
        raise ValueError(f"Unsupported distortion mode: {mode}")
    p.fold_amount = float(dp.get("fold_amount", 1.0))
    p.bias = float(dp.get("bias", 0.0))
    a = 1.0 + 4.0 * float(np.clip(float(dp.get("warmth", 0.5)), 0.0, 1.0))
    p.tube_gain = a * max(float(dp.get("drive", 1.0)), 0.0)
    p.tube_norm = 1.0 / float(np.tanh(a))
    # limiter
    ceiling, lookahead, coeff = limiter_constants(sr, limiter_ceiling_db)
    p.limiter_on, p.lookahead, p.ceiling_lin, p.release_coeff = int(bool(limiter_on)), lookahead, ceiling, coeff
    # mix (float32 arithmetic with weak Python scalars, dsp/pipeline.py:894-900)
    dw = float(np.clip(dry_wet, 0.0, 1.0))
    p.wet = float(np.float32(dw))
    p.dry = float(np.float32(1.0 - dw))
    p.apply_trim = int(output_trim_db != 0.0)
    p.trim_gain = float(np.float32(10.0 ** (output_trim_db / 20.0)))
    p.delta_listen = int(bool(delta_listen))
    # multiband (dsp/pipeline.py:1011-1110)
    p.multiband = int(bool(use_multiband))
    keep = []
    if use_multiband:
        sos_lo, sos_hi = design_linkwitz_riley_sos(sr, crossover_hz)
        _sos_fill(p.sos_low, sos_lo)
        _sos_fill(p.sos_high, sos_hi)
        p.low_delay = n_fft // 2                      # :1056; the filter-delay term cancels (:380-386)
        p.low_gain = max(float(lowband_drive), 0.0)   # dsp/saturation.py:44
        p.low_norm = 1.0 / float(np.tanh(3.0))        # dsp/saturation.py:54
        p.apply_low_trim = int(low_trim_db != 0.0)
        p.low_trim_gain = float(np.float32(10.0 ** (low_trim_db / 20.0)))
        blend = 0.0 < mono_strength < 1.0             # :1064-1070, mono-maker is the identity for 1-D audio
        p.apply_mono_blend = int(blend)
        p.mono_a = float(np.float32(mono_strength))
        p.mono_b = float(np.float32(1.0 - mono_strength))
    p.fx_mode = 0
    n_frames = 1 + int(n_samples) // (n_fft // 4)
    fx_rng = None
    fx_passes = 0
    quant_on = (not passthrough_test) and (p.pre_quant or p.post_quant)
    if use_multiband and quant_on:   # FX runs inside the quantiser call of the HIGH band only (:291, :313-314)
        fx = resolve_spectral_fx(spectral_fx_mode, spectral_fx_strength, spectral_fx_params)
        p.fx_mode, p.fx_a, p.fx_b, p.fx_c = fx["fx_mode"], fx["a"], fx["b"], fx["c"]
        if p.fx_mode:
            fx_rng = fx["rng"]
            fx_passes = int(p.pre_quant) + int(p.post_quant)
            p.fx_table_frames = n_frames if fx_rng else 0
    tables = None
    tb = mask = None
    if not passthrough_test and (p.pre_quant or p.post_quant):
        freqs = np.fft.rfftfreq(n_fft, d=1.0 / sr)  # same call as dsp/stft_utils.py:95
        if harmonic_lock_hz > 0.0:
            tb = build_harmonic_target_bins(freqs, harmonic_lock_hz)     # dsp/pipeline.py:277-278
        else:
            tb = build_target_bins_for_freqs(freqs, key, scale)          # :280
        mask = build_quantize_band_mask(freqs, sub_cut_hz, air_cut_hz)   # :282
        tb32 = np.ascontiguousarray(tb, dtype=np.int32)
        m8 = np.ascontiguousarray(mask, dtype=np.uint8)
        kw = np.ascontiguousarray(smear_kernel(2), dtype=np.float64)
        tables = _lib.QdTables()
        tables.n_bins = int(freqs.shape[0])
        tables.target_bins = tb32.ctypes.data_as(C.POINTER(C.c_int32))
        tables.active_mask = m8.ctypes.data_as(C.POINTER(C.c_uint8))
        tables.snap = float(np.clip(snap_strength, 0.0, 1.0))   # dsp/quantizer.py:405
        tables.smear = float(np.clip(smear, 0.0, 1.0))          # :406
        tables.smear_radius = 2
        tables.smear_w = kw.ctypes.data_as(C.POINTER(C.c_double))
        keep += [tb32, m8, kw]
    else:
        # still validate key / scale like the reference would on its first quantizer call
        pass
    p.spectral_freeze = int(bool(spectral_freeze) and quant_on)   # dsp/pipeline.py:285-287 (any band)
    formant_on = bool(quant_on and float(formant_shift) != 0.0)  # dsp/pipeline.py:306-310 (any band)
    p.formant_ratio = float(2.0 ** (float(formant_shift) / 12.0)) if formant_on else 0.0  # dsp/spectral_fx.py:173
    p.formant_order = 30                                                                  # dsp/spectral_fx.py:120
    p.precision = choose_precision(precision, n_fft, tb, mask, fx_active=bool(p.fx_mode) or bool(p.spectral_freeze),
                                   formant_active=formant_on, fx_mode=int(p.fx_mode),
                                   freeze_active=bool(p.spectral_freeze))
    return Resolved(params=p, tables=tables, keepalive=tuple(keep), target_bins=tb, active_mask=mask,
                    fx_rng=fx_rng, fx_passes=fx_passes, n_frames=n_frames, n_bins=n_fft // 2 + 1)
