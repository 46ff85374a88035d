"""The reference's four presets (values of presets/quantum_distortion_presets.json, accessor API of
quantum_distortion/presets.py:63-80)."""
from __future__ import annotations

from typing import Any, Dict, List


def _p(desc, key, scale, snap, smear, mode, fold, bias, drive, warmth, ceiling, dry_wet, post=True) -> Dict[str, Any]:
    return {"description": desc, "key": key, "scale": scale, "snap_strength": snap, "smear": smear,
            "bin_smoothing": True, "pre_quant": True, "post_quant": post, "distortion_mode": mode,
            "distortion_params": {"fold_amount": fold, "bias": bias, "drive": drive, "warmth": warmth},
            "limiter_on": True, "limiter_ceiling_db": ceiling, "dry_wet": dry_wet}


_PRESETS: Dict[str, Dict[str, Any]] = {
    "Chordal Noise Wash": _p("Turn noisy or wideband content into an in-key, smeared harmonic wash.",
                             "C", "minor", 0.85, 0.6, "wavefold", 3.5, 0.0, 1.0, 0.5, -1.0, 1.0),
    "Controlled Dubstep Growl": _p("Aggressive, folded bass textures that stay locked to root + fifth.",
                                   "F", "minor", 0.9, 0.3, "wavefold", 5.0, 0.1, 1.0, 0.5, -1.0, 1.0),
    "Perc To Tonal Clang": _p("Push percussive hits into in-key metallic pitched impacts.",
                              "D", "minor", 0.75, 0.4, "tube", 1.0, 0.0, 4.0, 0.7, -2.0, 1.0),
    "Subtle Tube Glue": _p("Gentle saturation with light quantization to keep things musical.",
                           "C", "major", 0.4, 0.2, "tube", 1.0, 0.0, 2.0, 0.3, -1.0, 0.7, post=False),
}


def list_presets() -> List[str]:
    return sorted(_PRESETS.keys())


def get_preset(name: str) -> Dict[str, Any]:
    if name not in _PRESETS:
        raise KeyError(f"Preset not found: {name}")
    out = dict(_PRESETS[name])
    out["distortion_params"] = dict(out["distortion_params"])
    return out
