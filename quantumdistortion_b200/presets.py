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


# Spectral-FX presets of the high band (dsp/spectral_fx.py:34-113): name -> FX mode, strength (0..1) and the
# spectral_fx_params overrides, as scripts/quick_regression_suite.py:65-84 hands them to the renderer.
def _fx(fx_mode: str, strength: float, description: str, **params) -> Dict[str, Any]:
    return {"mode": fx_mode, "distortion_strength": strength, "params": params, "description": description}


SPECTRAL_FX_PRESETS: Dict[str, Dict[str, Any]] = {
    "sub_safe_glue": _fx("bitcrush", 0.35, "Sub-safe subtle log-domain bitcrush for main bass growls.",
                         method="log", step_db=2.0, threshold=0.0),
    "digital_growl": _fx("bitcrush", 0.55, "More obvious digital grit for aggressive growls.", method="log", step_db=3.0),
    "hard_crush_fx": _fx("bitcrush", 0.8, "Heavy bitcrush for stabs and FX, not main bass.", method="uniform", step=0.07),
    "gentle_movement": _fx("phase_dispersal", 0.25, "Small phase rotation for subtle shimmer.", randomized=False),
    "laser_zap": _fx("phase_dispersal", 0.6, "Laser/zap phase dispersal for neuro-style tops.", randomized=True),
    "phase_chaos_fx": _fx("phase_dispersal", 0.9, "Extreme phase chaos for risers and FX beds.", randomized=True),
    "stereo_smear": _fx("bin_scramble", 0.3, "Subtle local swaps for smeared high-end texture.", mode="swap"),
    "grainy_top": _fx("bin_scramble", 0.55, "Noticeable granular smear on the high band.", mode="random_pick"),
    "granular_shred": _fx("bin_scramble", 0.85, "Heavily shredded high band for FX-only usage.", mode="random_pick"),
}


def spectral_fx_preset_kwargs(name: str) -> Dict[str, Any]:
    """process_audio keyword arguments of one spectral-FX preset (KeyError for an unknown name)."""
    p = SPECTRAL_FX_PRESETS[name]
    return {"spectral_fx_mode": p["mode"], "spectral_fx_strength": float(p["distortion_strength"]),
            "spectral_fx_params": dict(p["params"])}
