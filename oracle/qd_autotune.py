"""oracle/qd_autotune.py -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the reference's ``autotune_v1`` mode (quantum_distortion/dsp/autotune.py and the branch
at dsp/pipeline.py:537-601): zero-phase band split, YIN pitch track with a note-hold state machine, granular
two-tap pitch shifter, envelope-followed sub oscillator, then the shared distortion / limiter / mix tail.
Only tests/ and __graft_entry__.smoke() may import this file.

Pinned: tests/golden/autotune.npz is produced by the LIVE reference (tests/golden/make_golden.py) and
tests/test_oracle.py checks every stage and the whole path against it bit for bit.  The NumPy scalar
semantics matter here (NEP 50, NumPy >= 2: a Python float combined with a float32 value computes in
float32), so the loops below keep NumPy scalars where the reference has them; ``file:line`` citations
are relative to /root/reference/quantum_distortion.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.signal

from . import qd_oracle as orc


@dataclass(frozen=True)
class AutotuneConfig:
    """dsp/autotune.py:18-45 (only the fields process_audio can reach, plus the fixed detector constants)."""
    key: str = "D"
    scale: str = "minor"
    strength: float = 1.0
    sub_enabled: bool = True
    sub_source: str = "root"
    sub_note: str = "C"
    sub_scale_degree: int = 0
    sub_octave: int = 2
    sub_level: float = 0.35
    sub_preserve_mix: float = 0.15
    sub_cut_hz: float = 110.0
    air_cut_hz: float = 5000.0
    air_mix: float = 1.0
    detector_low_hz: float = 110.0
    detector_high_hz: float = 3000.0
    detector_frame_size: int = 4096
    detector_hop_size: int = 512
    detector_min_confidence: float = 0.72
    detector_rms_threshold: float = 0.01
    detector_flatness_threshold: float = 0.55
    note_change_cents: float = 40.0
    note_confirm_frames: int = 3
    note_release_frames: int = 2
    grain_size: int = 1024
    buffer_size: int = 4096
    warm_sub: bool = False


# --------------------------------------------------------------------------- zero-phase filters
def butter_sos(sr: int, cutoff_hz: float, btype: str, order: int = 4) -> Optional[np.ndarray]:
    """dsp/autotune.py:88-93."""
    if cutoff_hz <= 0.0:
        return None
    wn = float(np.clip(cutoff_hz / (sr / 2.0), 1e-5, 0.999))
    return scipy.signal.butter(order, wn, btype=btype, output="sos")


def sosfiltfilt_restated(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """What scipy.signal.sosfiltfilt(sos, x) computes with its defaults (padtype="odd", padlen=None), spelled out
    so that the CUDA kernel has a sample-level specification:
      ntaps = 2*n_sections + 1 - min(#(b2 == 0), #(a2 == 0));  edge = 3*ntaps
      ext   = odd extension of x by `edge` samples on both sides, computed in x's dtype, then float64
      zi    = sosfilt_zi(sos);  forward sosfilt over ext from the state zi*ext[0], backward from zi*y[-1]
    """
    sos = np.asarray(sos, dtype=float)
    x = np.asarray(x)   # the odd extension is formed in the INPUT dtype (float32 for the reference's calls) ...
    n_sections = sos.shape[0]
    ntaps = 2 * n_sections + 1
    ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    edge = 3 * ntaps
    if x.shape[0] <= edge:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % edge)
    left = 2.0 * x[0] - x[edge:0:-1]
    right = 2.0 * x[-1] - x[-2:-(edge + 2):-1]
    ext = np.concatenate([left, x, right]).astype(float)   # ... and only then promoted to float64
    zi = scipy.signal.sosfilt_zi(sos)

    def run(v, z0):
        z = z0.copy()
        out = np.empty_like(v)
        for i in range(v.shape[0]):
            s = v[i]
            for k in range(n_sections):
                b0, b1, b2, _, a1, a2 = sos[k]
                o = b0 * s + z[k, 0]
                z[k, 0] = b1 * s - a1 * o + z[k, 1]
                z[k, 1] = b2 * s - a2 * o
                s = o
            out[i] = s
        return out

    y = run(ext, zi * ext[0])
    y = run(y[::-1], zi * y[-1])[::-1]
    return y[edge:-edge]


def zero_phase(audio: np.ndarray, sr: int, cutoff_hz: float, btype: str) -> np.ndarray:
    """dsp/autotune.py:96-100."""
    sos = butter_sos(sr, cutoff_hz, btype)
    if sos is None:
        return audio.astype(np.float32)
    return scipy.signal.sosfiltfilt(sos, audio).astype(np.float32)


def split_sub_body_air(audio, sr, sub_cut_hz, air_cut_hz):
    """dsp/autotune.py:103-113."""
    audio = np.asarray(audio, dtype=np.float32)
    sub = zero_phase(audio, sr, sub_cut_hz, "low")
    air = zero_phase(audio, sr, air_cut_hz, "high")
    body = (audio - sub - air).astype(np.float32)
    return sub, body, air


def detector_sidechain(body, sr, low_hz, high_hz):
    """dsp/autotune.py:116-127."""
    band = np.asarray(body, dtype=np.float32)
    if low_hz > 0.0:
        band = zero_phase(band, sr, low_hz, "high")
    if high_hz > 0.0:
        band = zero_phase(band, sr, high_hz, "low")
    return band.astype(np.float32)


# --------------------------------------------------------------------------- pitch detector
def spectral_flatness(frame: np.ndarray) -> float:
    """dsp/autotune.py:130-137 (np.hanning is the SYMMETRIC Hann window)."""
    spec = np.abs(np.fft.rfft(frame * np.hanning(len(frame)))) + 1e-8
    geo = float(np.exp(np.mean(np.log(spec))))
    ari = float(np.mean(spec))
    return 1.0 if ari <= 1e-8 else geo / ari


def yin_tables(frame: np.ndarray, sr: int, min_freq: float, max_freq: float):
    """Difference function and cumulative-mean-normalised difference of dsp/autotune.py:149-170;
    returns (min_tau, max_tau, diff, cmnd) or None for a silent / degenerate frame."""
    frame = np.asarray(frame, dtype=np.float64)
    if np.max(np.abs(frame)) < 1e-6:
        return None
    c = frame - np.mean(frame)
    max_tau = min(int(sr / max(min_freq, 1e-6)), max(2, len(c) // 2 - 1))
    min_tau = max(2, int(sr / max(max_freq, 1e-6)))
    if max_tau <= min_tau:
        return None
    diff = np.zeros(max_tau + 1)
    for tau in range(1, max_tau + 1):
        d = c[:-tau] - c[tau:]
        diff[tau] = np.dot(d, d)
    cmnd = np.ones_like(diff)
    run = 0.0
    for tau in range(1, max_tau + 1):
        run += diff[tau]
        if run > 0.0:
            cmnd[tau] = diff[tau] * tau / run
    return min_tau, max_tau, diff, cmnd


def yin_pick(cmnd: np.ndarray, min_tau: int, max_tau: int, sr: int, min_freq: float, max_freq: float,
             threshold: float = 0.15) -> Tuple[float, float]:
    """dsp/autotune.py:172-197: first dip below the threshold, walked to its local minimum, parabolic refinement."""
    est = -1
    for tau in range(min_tau, max_tau + 1):
        if cmnd[tau] < threshold:
            while tau + 1 <= max_tau and cmnd[tau + 1] < cmnd[tau]:
                tau += 1
            est = tau
            break
    if est == -1:
        return 0.0, 0.0
    better = float(est)
    if min_tau < est < max_tau:
        s0, s1, s2 = cmnd[est - 1], cmnd[est], cmnd[est + 1]
        den = 2.0 * (s0 - 2.0 * s1 + s2)
        if abs(den) > 1e-12:
            better = est + (s0 - s2) / den
    pitch = float(sr / better) if better > 0 else 0.0
    if pitch < min_freq or pitch > max_freq:
        return 0.0, 0.0
    return pitch, float(np.clip(1.0 - cmnd[est], 0.0, 1.0))


def detect_pitch_yin(frame, sr, min_freq=70.0, max_freq=1200.0, threshold=0.15):
    """dsp/autotune.py:140-198."""
    t = yin_tables(frame, sr, min_freq, max_freq)
    if t is None:
        return 0.0, 0.0
    min_tau, max_tau, _, cmnd = t
    return yin_pick(cmnd, min_tau, max_tau, sr, min_freq, max_freq, threshold)


def nearest_scale_freq(freq: float, key: str, scale: str) -> float:
    """dsp/autotune.py:65-85."""
    if freq <= 0.0:
        return freq
    root = orc.note_name_to_pitch_class(key)
    midi = orc._freq_to_midi(freq)
    in_oct = ((midi - root) % 12.0 + 12.0) % 12.0
    base = midi - in_oct
    best, best_d = round(midi), np.inf
    for octv in (-1, 0, 1):
        for iv in orc.SCALE_INTERVALS[scale]:
            cand = base + iv + octv * 12.0
            d = abs(midi - cand)
            if d < best_d:
                best_d, best = d, cand
    return 440.0 * (2.0 ** ((float(best) - 69.0) / 12.0))


def frame_features(det: np.ndarray, sr: int, cfg: AutotuneConfig):
    """Per detector frame (dsp/autotune.py:217-236): (centers, rms, flatness, pitch, confidence)."""
    det = np.asarray(det, dtype=np.float32)
    fs = int(max(1024, cfg.detector_frame_size))
    hop = int(max(128, cfg.detector_hop_size))
    lo = max(60.0, cfg.detector_low_hz * 0.65)
    hi = max(cfg.detector_high_hz, 400.0)
    centers, rms, flat, pitch, conf = [], [], [], [], []
    for start in range(0, len(det), hop):
        fr = det[start:start + fs]
        if len(fr) < fs:
            fr = np.pad(fr, (0, fs - len(fr)), mode="constant")
        centers.append(min(len(det) - 1, start + fs // 2))
        rms.append(float(np.sqrt(np.mean(fr * fr))))
        flat.append(spectral_flatness(fr))
        p, c = detect_pitch_yin(fr, sr, min_freq=lo, max_freq=hi)
        pitch.append(p)
        conf.append(c)
    return (np.array(centers, dtype=np.float64), np.array(rms), np.array(flat), np.array(pitch), np.array(conf))


def note_hold_ratios(rms, flat, pitch, conf, cfg: AutotuneConfig) -> np.ndarray:
    """The per-frame state machine of dsp/autotune.py:210-277 -> correction ratio per frame."""
    held = cand = 0.0
    cand_n = 0
    rel = cfg.note_release_frames + 1
    last = 1.0
    out = np.empty(len(pitch))
    for i in range(len(pitch)):
        p = float(pitch[i])
        voiced = (p > 0.0 and rms[i] >= cfg.detector_rms_threshold and flat[i] <= cfg.detector_flatness_threshold
                  and conf[i] >= cfg.detector_min_confidence)
        if voiced:
            tgt = nearest_scale_freq(p, cfg.key, cfg.scale)
            if held <= 0.0:
                held, cand, cand_n = tgt, 0.0, 0
            else:
                dc = abs(1200.0 * np.log2(max(tgt, 1e-6) / max(held, 1e-6)))
                if dc >= cfg.note_change_cents:
                    cd = abs(1200.0 * np.log2(max(tgt, 1e-6) / max(cand, 1e-6))) if cand > 0.0 else np.inf
                    if cand > 0.0 and cd < 20.0:
                        cand_n += 1
                    else:
                        cand, cand_n = tgt, 1
                    if cand_n >= cfg.note_confirm_frames:
                        held, cand, cand_n = cand, 0.0, 0
                else:
                    cand, cand_n = 0.0, 0
            r = float(np.clip(1.0 + cfg.strength * ((held / p) - 1.0), 0.5, 2.0))
            rel = 0
            last = r
        else:
            rel += 1
            if held > 0.0 and rel <= cfg.note_release_frames:
                r = last
            else:
                held = cand = 0.0
                cand_n = 0
                r = last = 1.0
        out[i] = r
    return out


def ratio_track(det: np.ndarray, sr: int, cfg: AutotuneConfig) -> np.ndarray:
    """dsp/autotune.py:200-298, the ratio track only (float32 per sample)."""
    n = len(det)
    centers, rms, flat, pitch, conf = frame_features(det, sr, cfg)
    if centers.size == 0:
        return np.ones(n, dtype=np.float32)
    ratios = note_hold_ratios(rms, flat, pitch, conf, cfg)
    return np.interp(np.arange(n, dtype=np.float64), centers, ratios).astype(np.float32)


# --------------------------------------------------------------------------- granular shifter
def granular_pitch_shift(audio: np.ndarray, ratios: np.ndarray, grain_size: int = 1024, buffer_size: int = 4096):
    """dsp/autotune.py:301-360.  The circular buffer only ever returns one of the last `buffer_size` inputs (zero
    before the start), so it is addressed as a plain index here; the float32 / float64 mix of the reference's
    scalar arithmetic is kept: the two-point interpolation rounds to float32, weights and sums are float64."""
    audio = np.asarray(audio, dtype=np.float32)
    ratios = np.asarray(ratios, dtype=np.float32)
    if len(audio) == 0 or np.allclose(ratios, 1.0, atol=1e-3):
        return audio.copy()
    max_delay = int(max(256, grain_size))
    size = int(max(max_delay * 2, buffer_size))
    if size & (size - 1):
        size = 1 << int(np.ceil(np.log2(size)))
    n = len(audio)
    taps = [0.25 * max_delay, 0.75 * max_delay]
    out = np.zeros(n, dtype=np.float32)
    zero = np.float32(0.0)

    def held(i, wpos_age):  # sample written `wpos_age` positions ago relative to index i (0 = the current one)
        j = i - wpos_age
        return audio[j] if j >= 0 else zero

    for i in range(n):
        slope = 1.0 - float(np.clip(ratios[i], 0.5, 2.0))
        mixed = 0.0
        wsum = 0.0
        w = i % size  # write position
        for t in range(2):
            taps[t] += slope
            while taps[t] < 0.0:
                taps[t] += max_delay
            while taps[t] >= max_delay:
                taps[t] -= max_delay
            phase = taps[t] / max_delay
            weight = 0.5 * (1.0 - np.cos(2.0 * np.pi * phase))
            rp = (w - taps[t] + size) % size
            b = int(rp) & (size - 1)
            nx = (b + 1) & (size - 1)
            fr = rp - int(rp)
            s0 = held(i, (w - b) % size)
            s1 = held(i, (w - nx) % size)
            sample = s0 * (1.0 - fr) + s1 * fr          # float32 (NEP 50)
            mixed += sample * weight                     # float64
            wsum += weight
        out[i] = mixed / wsum if wsum > 1e-6 else 0.0
    lat = max_delay // 2
    if lat > 0 and n > lat:
        out = np.concatenate([out[lat:], np.zeros(lat, dtype=np.float32)])
    return out.astype(np.float32)


# --------------------------------------------------------------------------- sub layer
def envelope_follow(audio: np.ndarray, sr: int, attack_ms: float = 8.0, release_ms: float = 90.0) -> np.ndarray:
    """dsp/autotune.py:363-377.  `current` turns into a float32 scalar after the first sample (NEP 50), so the
    whole recurrence runs in float32 with the coefficients rounded to float32."""
    a = np.abs(np.asarray(audio, dtype=np.float32))
    att = float(np.exp(-1.0 / max(1.0, attack_ms * 0.001 * sr)))
    rel = float(np.exp(-1.0 / max(1.0, release_ms * 0.001 * sr)))
    env = np.zeros_like(a, dtype=np.float32)
    cur = 0.0
    for i, s in enumerate(a):
        k = att if s > cur else rel
        cur = s + k * (cur - s)
        env[i] = cur
    return env


def sub_frequency(cfg: AutotuneConfig) -> float:
    """dsp/autotune.py:380-396."""
    if cfg.sub_source == "manual":
        pc = orc.note_name_to_pitch_class(cfg.sub_note)
    elif cfg.sub_source == "scale_degree":
        iv = orc.SCALE_INTERVALS[cfg.scale]
        pc = (orc.note_name_to_pitch_class(cfg.key) + iv[int(np.clip(cfg.sub_scale_degree, 0, len(iv) - 1))]) % 12
    else:
        pc = orc.note_name_to_pitch_class(cfg.key)
    midi = 12 * (int(np.clip(cfg.sub_octave, 0, 6)) + 1) + pc
    return float(440.0 * (2.0 ** ((float(midi) - 69.0) / 12.0)))


def sub_layer(reference_audio: np.ndarray, sr: int, cfg: AutotuneConfig) -> np.ndarray:
    """dsp/autotune.py:399-423."""
    if not cfg.sub_enabled or cfg.sub_level <= 0.0:
        return np.zeros_like(reference_audio, dtype=np.float32)
    f = sub_frequency(cfg)
    if f <= 0.0:
        return np.zeros_like(reference_audio, dtype=np.float32)
    ref = np.asarray(reference_audio, dtype=np.float32)
    env = envelope_follow(ref, sr)
    top = float(np.max(env))
    if top > 1e-6:
        env = env / top
    phase = 2.0 * np.pi * f * np.arange(len(ref), dtype=np.float32) / float(sr)   # float32 ramp
    osc = np.sin(phase)
    if cfg.warm_sub:
        osc += 0.15 * np.sin(2.0 * phase)
        osc /= max(float(np.max(np.abs(osc))), 1.0)
    return (cfg.sub_level * env * osc).astype(np.float32)


# --------------------------------------------------------------------------- the mode
def apply_autotune_v1(audio: np.ndarray, sr: int, cfg: AutotuneConfig) -> Dict[str, np.ndarray]:
    """dsp/autotune.py:426-447."""
    audio = np.asarray(audio, dtype=np.float32)
    sub, body, air = split_sub_body_air(audio, sr, cfg.sub_cut_hz, cfg.air_cut_hz)
    det = detector_sidechain(body, sr, cfg.detector_low_hz, cfg.detector_high_hz)
    ratios = ratio_track(det, sr, cfg)
    corrected = granular_pitch_shift(body, ratios, cfg.grain_size, cfg.buffer_size)
    layer = sub_layer(audio, sr, cfg)
    low = (cfg.sub_preserve_mix * sub) + layer if cfg.sub_enabled else sub
    out = low + corrected + (cfg.air_mix * air)
    return {"output": out.astype(np.float32), "sub": sub, "body": body, "air": air, "detector": det,
            "ratio_track": ratios, "corrected_body": corrected.astype(np.float32), "sub_layer": layer}


def process_audio_autotune(audio, sr=48000, key="D", scale="minor", snap_strength=1.0, pre_quant=True,
                           distortion_mode="wavefold", distortion_params=None, limiter_on=True,
                           limiter_ceiling_db=-1.0, dry_wet=1.0, output_trim_db=0.0, delta_listen=False,
                           sub_enabled=True, sub_source="root", sub_note="C", sub_scale_degree=0, sub_octave=2,
                           sub_level=0.35, sub_cut_hz=110.0, air_cut_hz=5000.0, air_mix=1.0):
    """process_audio(quantize_mode="autotune_v1") without FX / freeze / formant / lock options
    (dsp/pipeline.py:1315-1327 forces single band; :537-601 is the branch; :180-223 the shared tail)."""
    x = np.asarray(audio, dtype=np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1).astype(np.float32)
    tap_in = x.copy()
    n = x.shape[0]
    if pre_quant and snap_strength > 0.0:
        cfg = AutotuneConfig(key=key, scale=scale, strength=float(np.clip(snap_strength, 0.0, 1.0)),
                             sub_enabled=sub_enabled, sub_source=sub_source, sub_note=sub_note,
                             sub_scale_degree=sub_scale_degree, sub_octave=sub_octave, sub_level=sub_level,
                             sub_cut_hz=sub_cut_hz, air_cut_hz=air_cut_hz, air_mix=air_mix)
        x_pre = apply_autotune_v1(x, sr, cfg)["output"].astype(np.float32)
    else:
        x_pre = x.copy()
    dp = distortion_params or {}
    x_dist = orc.apply_distortion(x_pre, distortion_mode or "wavefold", fold_amount=float(dp.get("fold_amount", 1.0)),
                                  bias=float(dp.get("bias", 0.0)), drive=float(dp.get("drive", 1.0)),
                                  warmth=float(dp.get("warmth", 0.5)))
    if limiter_on:
        x_lim, _ = orc.peak_limiter(x_dist, sr, ceiling_db=limiter_ceiling_db, lookahead_ms=5.0, release_ms=30.0)
    else:
        x_lim = x_dist.copy()
    dw = float(np.clip(dry_wet, 0.0, 1.0))
    y = (dw * x_lim) + ((1.0 - dw) * tap_in)
    if output_trim_db != 0.0:
        y = y * (10.0 ** (output_trim_db / 20.0))
    y = orc._fit(y.astype(np.float32), n)
    taps = {"input": tap_in, "pre_quant": x_pre.copy(), "post_dist": x_dist.copy(), "output": y.copy()}
    if delta_listen:
        m = min(len(tap_in), len(y))
        y = (tap_in[:m] - y[:m]).astype(np.float32)
        taps["output"] = y.copy()
    return y, taps
