"""Named pipeline cases used by make_golden.py (live reference) and the parity tests.

Every case is (signal kind, seed, n_samples, sr, n_fft, np.random seed or None, kwargs for
process_audio).  quantize_mode="spectral_bins" is always passed.
"""
from __future__ import annotations

GROWL = dict(key="F", scale="minor", snap_strength=0.9, smear=0.3, bin_smoothing=True,
             pre_quant=True, post_quant=True, distortion_mode="wavefold",
             distortion_params={"fold_amount": 5.0, "bias": 0.1, "drive": 1.0, "warmth": 0.5},
             limiter_on=True, limiter_ceiling_db=-1.0, dry_wet=1.0)
CLANG = dict(key="D", scale="minor", snap_strength=0.75, smear=0.4, bin_smoothing=True,
             pre_quant=True, post_quant=True, distortion_mode="tube",
             distortion_params={"fold_amount": 1.0, "bias": 0.0, "drive": 4.0, "warmth": 0.7},
             limiter_on=True, limiter_ceiling_db=-2.0, dry_wet=1.0)
GLUE = dict(key="C", scale="major", snap_strength=0.4, smear=0.2, bin_smoothing=True,
            pre_quant=True, post_quant=False, distortion_mode="tube",
            distortion_params={"fold_amount": 1.0, "bias": 0.0, "drive": 2.0, "warmth": 0.3},
            limiter_on=True, limiter_ceiling_db=-1.0, dry_wet=0.7)

N = 12000  # 0.25 s @ 48 kHz -> 24 frames at n_fft 2048

# name: (kind, seed, n, sr, n_fft, rng_seed, kwargs)
CASES = {
    "sb_default_bass": ("bass", 0, N, 48000, 2048, None, {}),
    "sb_default_noise": ("noise", 1, N, 48000, 2048, None, {}),
    "sb_default_loud": ("loud", 2, N, 48000, 2048, None, {}),
    "sb_default_441": ("bass", 3, 11025, 44100, 2048, None, {}),
    "sb_ragged_len": ("bass", 4, 5003, 48000, 2048, None, {}),
    "sb_short": ("noise", 5, 700, 48000, 2048, None, {}),
    "sb_passthrough": ("bass", 6, N, 48000, 2048, None, {"passthrough_test": True}),
    "sb_growl": ("loud", 7, N, 48000, 2048, None, dict(GROWL)),
    "sb_clang_tube": ("bass", 8, N, 48000, 2048, None, dict(CLANG)),
    "sb_glue_pre_only_drywet": ("bass", 9, N, 48000, 2048, None, dict(GLUE)),
    "sb_post_only": ("bass", 10, N, 48000, 2048, None, {"pre_quant": False}),
    "sb_no_quant": ("bass", 11, N, 48000, 2048, None, {"snap_strength": 0.0}),
    "sb_no_limiter_trim_delta": ("loud", 12, N, 48000, 2048, None,
                                 {"limiter_on": False, "output_trim_db": -3.0, "delta_listen": True,
                                  "dry_wet": 0.5}),
    "sb_no_smooth_no_smear": ("noise", 13, N, 48000, 2048, None, {"bin_smoothing": False, "smear": 0.0}),
    "sb_wide_mask": ("noise", 14, N, 48000, 2048, None, {"sub_cut_hz": 0.0, "air_cut_hz": 0.0,
                                                          "key": "A", "scale": "pentatonic"}),
    "sb_harmonic_lock": ("bass", 15, N, 48000, 2048, None, {"harmonic_lock_hz": 55.0}),
    "mb_default": ("bass", 20, N, 48000, 2048, None, {"use_multiband": True, "crossover_hz": 300.0}),
    "mb_loud_drive": ("loud", 21, N, 48000, 2048, None, {"use_multiband": True, "crossover_hz": 300.0,
                                                          "lowband_drive": 2.5, "dry_wet": 0.8}),
    "mb_growl_bitcrush": ("loud", 22, N, 48000, 2048, 1234,
                          dict(GROWL, use_multiband=True, spectral_fx_mode="bitcrush",
                               spectral_fx_strength=0.5)),
    "mb_growl_bitcrush_lo": ("bass", 23, N, 48000, 2048, 1234,
                             dict(GROWL, use_multiband=True, spectral_fx_mode="bitcrush",
                                  spectral_fx_strength=0.3)),
    "mb_growl_dispersal": ("loud", 24, N, 48000, 2048, 1234,
                           dict(GROWL, use_multiband=True, spectral_fx_mode="phase_dispersal",
                                spectral_fx_strength=0.6)),
    "mb_growl_dispersal_det": ("bass", 25, N, 48000, 2048, 1234,
                               dict(GROWL, use_multiband=True, spectral_fx_mode="phase_dispersal",
                                    spectral_fx_strength=0.3)),
    "mb_growl_scramble_pick": ("loud", 26, N, 48000, 2048, 1234,
                               dict(GROWL, use_multiband=True, spectral_fx_mode="bin_scramble",
                                    spectral_fx_strength=0.55)),
    "mb_growl_scramble_swap": ("bass", 27, N, 48000, 2048, 1234,
                               dict(GROWL, use_multiband=True, spectral_fx_mode="bin_scramble",
                                    spectral_fx_strength=0.3)),
    "sb_freeze": ("bass", 28, N, 48000, 2048, None, {"spectral_freeze": True}),
    "mb_freeze_bitcrush": ("loud", 29, N, 48000, 2048, 1234,
                           dict(GROWL, use_multiband=True, spectral_freeze=True, spectral_fx_mode="bitcrush",
                                spectral_fx_strength=0.5)),
    "sb_formant_up": ("bass", 34, N, 48000, 2048, None, {"formant_shift": 3.0}),
    "sb_formant_down_growl": ("loud", 35, N, 48000, 2048, None, dict(GROWL, formant_shift=-5.0)),
    "mb_formant_bitcrush": ("loud", 36, N, 48000, 2048, 1234,
                            dict(GROWL, use_multiband=True, formant_shift=2.0, spectral_fx_mode="bitcrush",
                                 spectral_fx_strength=0.5)),
    "sb_formant_freeze": ("bass", 37, N, 48000, 2048, None, {"formant_shift": 7.0, "spectral_freeze": True}),
    "nfft1024_formant": ("bass", 38, 6000, 48000, 1024, None, {"formant_shift": -2.0}),
    "nfft512": ("bass", 30, 6000, 48000, 512, None, {}),
    "nfft1024": ("bass", 31, 6000, 48000, 1024, None, {}),
    "nfft4096": ("bass", 32, N, 48000, 4096, None, {}),
    "nfft8192": ("bass", 33, 20000, 48000, 8192, None, {}),
}


def make_signal(kind: str, seed: int, n: int, sr: int):
    from quantumdistortion_b200 import synth as signals
    if kind == "bass":
        return signals.bass_clip(seed, n, sr)
    if kind == "noise":
        return signals.noise_clip(seed, n)
    if kind == "loud":
        return signals.loud_clip(seed, n, sr)
    if kind == "tone":
        return signals.tone_clip(seed, n, sr)
    raise KeyError(kind)


# quantize_mode="autotune_v1" (dsp/autotune.py, dsp/pipeline.py:537-601): name -> (kind, seed, n, sr, kwargs)
AT_N = 24000  # 0.5 s @ 48 kHz -> 47 detector frames
AUTOTUNE_CASES = {
    "at_tone_default": ("tone", 0, AT_N, 48000, {}),
    "at_tone_half_strength_441": ("tone", 1, 22050, 44100, {"snap_strength": 0.5, "key": "F", "scale": "major"}),
    "at_bass_growl_nosub": ("bass", 80, AT_N, 48000, dict(GROWL, sub_enabled=False)),
    "at_tone_tube_drywet": ("tone", 2, AT_N, 48000, dict(CLANG, dry_wet=0.6, sub_source="scale_degree",
                                                          sub_scale_degree=4, sub_octave=1, sub_level=0.5,
                                                          air_mix=0.5, output_trim_db=-2.0)),
    "at_noise": ("noise", 81, 9000, 48000, {"sub_source": "manual", "sub_note": "G", "delta_listen": True}),
    "at_no_prequant": ("tone", 3, 9000, 48000, {"pre_quant": False}),
    "at_ragged_short": ("tone", 4, 5003, 48000, {"sub_cut_hz": 0.0, "air_cut_hz": 8000.0}),
}

# avg_cents_offset_from_scale (dsp/analyses.py:53-142): name -> (kind, seed, n, sr, key, scale, kwargs)
ANALYSIS_CASES = {
    "an_bass_dminor": ("bass", 60, N, 48000, "D", "minor", {}),
    "an_noise_top5": ("noise", 61, 6000, 48000, "A", "pentatonic", {"topn_peaks": 5, "min_db": -20.0}),
    "an_loud_1024": ("loud", 62, 9000, 44100, "F#", "dorian", {"frame_length": 1024}),
    "an_quiet": ("bass", 63, 6000, 48000, "C", "major", {"min_db": 40.0}),
}

# Streamlit V2 UI dicts (dsp/pipeline.py:923-1008): name -> (kind, seed, n, sr, rng_seed, config dict, extra kwargs)
UI_CASES = {
    "ui_full": ("loud", 50, N, 48000, 77,
                {"quantization": {"key": "E", "scale": "dorian", "sub_cut_hz": 90.0, "air_cut_hz": 6000.0},
                 "crossover_freq": 250.0,
                 "low_band": {"saturation_amount": 0.5, "mono_strength": 0.6, "output_trim_db": -1.5},
                 "high_band": {"bin_scrambling": 0.0, "phase_dispersal": 0.45, "mag_decimation": 0.7, "output_trim_db": -2.0},
                 "delta_listen": False}, {"snap_strength": 0.8}),
    "ui_defaults_scramble": ("bass", 51, N, 48000, 78, {"high_band": {}, "quantum_fx": {"fundamental_hz": 0.0}}, {}),
    "ui_bitcrush_delta": ("loud", 52, N, 48000, 79,
                          {"high_band": {"bin_scrambling": 0.0, "phase_dispersal": 0.0, "mag_decimation": 0.6},
                           "delta_listen": True}, {}),
}
