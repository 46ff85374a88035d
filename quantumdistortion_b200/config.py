"""Parameter surface of the STFT path, mirroring the reference field for field.

Reference: quantum_distortion/config.py:30-58 (defaults), :64-130 (PipelineConfig),
:132-151 (from_preset), :19-24 (ensure_mono_float32).  Same names, same defaults, so code
written against the reference's PipelineConfig keeps working.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

DEFAULT_SAMPLE_RATE = 48000
DEFAULT_KEY = "D"
DEFAULT_SCALE = "minor"
DEFAULT_SNAP_STRENGTH = 1.0
DEFAULT_SMEAR = 0.1
DEFAULT_BIN_SMOOTHING = True
DEFAULT_DISTORTION_MODE = "wavefold"
DEFAULT_LIMITER_ON = True
DEFAULT_LIMITER_CEILING_DB = -1.0
DEFAULT_DRY_WET = 1.0
# config.py:44 -- the reference's default is the time-domain pitch-tracking mode (quantumdistortion_b200/autotune.py);
# the STFT path (this package's headline) is quantize_mode="spectral_bins", or any spectral FX / freeze / formant /
# harmonic-lock option (dsp/pipeline.py:1315-1324).  Same default here, so a bare process_audio(x, sr),
# PipelineConfig() and PipelineConfig.from_preset() run what the same call runs on the reference.
DEFAULT_QUANTIZE_MODE = "autotune_v1"
DEFAULT_SUB_CUT_HZ = 110.0
DEFAULT_AIR_CUT_HZ = 5000.0
PREVIEW_ENABLED_DEFAULT = False
PREVIEW_MAX_SECONDS = 10.0
N_FFT_DEFAULT = 2048  # dsp/pipeline.py:149


def ensure_mono_float32(audio: np.ndarray) -> np.ndarray:
    """config.py:19-24: float32, stereo averaged to mono."""
    x = np.asarray(audio, dtype=np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1).astype(np.float32)
    return x


@dataclass
class PipelineConfig:
    """Same fields and defaults as the reference's PipelineConfig (config.py:64-130), quantize_mode="autotune_v1"
    included (:77); the ``sub_*`` / ``air_mix`` fields only matter for that mode."""
    key: str = DEFAULT_KEY
    scale: str = DEFAULT_SCALE
    quantize_mode: str = DEFAULT_QUANTIZE_MODE
    snap_strength: float = DEFAULT_SNAP_STRENGTH
    smear: float = DEFAULT_SMEAR
    bin_smoothing: bool = DEFAULT_BIN_SMOOTHING
    pre_quant: bool = True
    post_quant: bool = True
    sub_enabled: bool = True
    sub_source: str = "root"
    sub_note: str = "C"
    sub_scale_degree: int = 0
    sub_octave: int = 2
    sub_level: float = 0.35
    sub_cut_hz: float = DEFAULT_SUB_CUT_HZ
    air_cut_hz: float = DEFAULT_AIR_CUT_HZ
    air_mix: float = 1.0
    distortion_mode: str = DEFAULT_DISTORTION_MODE
    distortion_params: Dict[str, Any] = field(default_factory=dict)
    limiter_on: bool = DEFAULT_LIMITER_ON
    limiter_ceiling_db: float = DEFAULT_LIMITER_CEILING_DB
    dry_wet: float = DEFAULT_DRY_WET
    preview_enabled: Optional[bool] = None
    use_multiband: bool = False
    crossover_hz: float = 300.0
    lowband_drive: float = 1.0
    passthrough_test: bool = False
    spectral_fx_mode: Optional[str] = None
    spectral_fx_strength: float = 0.0
    spectral_fx_params: Dict[str, Any] = field(default_factory=dict)
    spectral_freeze: bool = False
    formant_shift: float = 0.0
    harmonic_lock_hz: float = 0.0
    delta_listen: bool = False
    mono_strength: float = 1.0
    output_trim_db: float = 0.0

    @classmethod
    def from_preset(cls, preset_name: str) -> "PipelineConfig":
        """config.py:132-151 (quantize_mode stays the default "autotune_v1", :140)."""
        from .presets import get_preset
        p = get_preset(preset_name)
        return cls(key=str(p["key"]), scale=str(p["scale"]), quantize_mode=DEFAULT_QUANTIZE_MODE,
                   snap_strength=float(p["snap_strength"]), smear=float(p["smear"]),
                   bin_smoothing=bool(p["bin_smoothing"]), pre_quant=bool(p["pre_quant"]),
                   post_quant=bool(p["post_quant"]), distortion_mode=str(p["distortion_mode"]),
                   distortion_params=dict(p["distortion_params"]), limiter_on=bool(p["limiter_on"]),
                   limiter_ceiling_db=float(p["limiter_ceiling_db"]), dry_wet=float(p["dry_wet"]))
