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

# Round-2 additions (fixtures in tests/golden/round2.npz; same tuple layout plus an optional `precision`):
# spectral_fx_params overrides, uniform bitcrush, spectral FX away from n_fft 2048, the float64 FX kernels,
# n_fft 8192 with every option, and autotune_v1 kept on the high band of a multiband render (snap_strength = 0).
_MB = dict(GROWL, use_multiband=True)
CASES_R2 = {
    "mb_bitcrush_uniform": ("loud", 100, N, 48000, 2048, 1234,
                            dict(_MB, spectral_fx_mode="bitcrush", spectral_fx_strength=0.5,
                                 spectral_fx_params={"method": "uniform"})),
    "mb_bitcrush_uniform_step_thr": ("bass", 101, N, 48000, 2048, 1234,
                                     dict(_MB, spectral_fx_mode="bitcrush", spectral_fx_strength=0.3,
                                          spectral_fx_params={"method": "uniform", "step": 0.004, "threshold": 0.003})),
    "mb_bitcrush_log_stepdb_thr": ("loud", 102, N, 48000, 2048, 1234,
                                   dict(_MB, spectral_fx_mode="bitcrush", spectral_fx_strength=0.7,
                                        spectral_fx_params={"step_db": 3.0, "threshold": 0.02})),
    "mb_bitcrush_unknown_method": ("bass", 103, N, 48000, 2048, 1234,
                                   dict(_MB, spectral_fx_mode="bitcrush", spectral_fx_strength=0.6,
                                        spectral_fx_params={"method": "cubic"})),
    "mb_dispersal_thresh_det": ("loud", 104, N, 48000, 2048, 1234,
                                dict(_MB, spectral_fx_mode="phase_dispersal", spectral_fx_strength=0.6,
                                     spectral_fx_params={"thresh": 0.05, "randomized": False})),
    "mb_dispersal_forced_random": ("bass", 105, N, 48000, 2048, 1234,
                                   dict(_MB, spectral_fx_mode="phase_dispersal", spectral_fx_strength=0.2,
                                        spectral_fx_params={"randomized": True, "rand_amt": 0.4, "amount": 1.1,
                                                            "thresh": 0.0})),
    "mb_scramble_window_swap": ("loud", 106, N, 48000, 2048, 1234,
                                dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.7,
                                     spectral_fx_params={"window": 9, "mode": "swap"})),
    "mb_scramble_window_pick": ("bass", 107, N, 48000, 2048, 1234,
                                dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.2,
                                     spectral_fx_params={"window": 4, "mode": "random_pick"})),
    "mb_scramble_unknown_mode": ("bass", 108, N, 48000, 2048, 1234,
                                 dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.5,
                                      spectral_fx_params={"mode": "shuffle"})),
    "nfft512_bitcrush": ("loud", 110, 6000, 48000, 512, 1234,
                         dict(_MB, spectral_fx_mode="bitcrush", spectral_fx_strength=0.5)),
    "nfft512_scramble_pick": ("bass", 111, 6000, 48000, 512, 1234,
                              dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.55)),
    "nfft1024_dispersal": ("loud", 112, 6000, 48000, 1024, 1234,
                           dict(_MB, spectral_fx_mode="phase_dispersal", spectral_fx_strength=0.6)),
    "nfft4096_bitcrush": ("loud", 113, N, 48000, 4096, 1234,
                          dict(_MB, spectral_fx_mode="bitcrush", spectral_fx_strength=0.5)),
    "nfft4096_dispersal": ("bass", 114, N, 48000, 4096, 1234,
                           dict(_MB, spectral_fx_mode="phase_dispersal", spectral_fx_strength=0.6)),
    "nfft4096_scramble_pick": ("loud", 115, N, 48000, 4096, 1234,
                               dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.55)),
    "nfft4096_scramble_swap": ("bass", 116, N, 48000, 4096, 1234,
                               dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.3)),
    "nfft4096_formant": ("bass", 117, N, 48000, 4096, None, {"formant_shift": 4.0}),
    "nfft8192_bitcrush": ("loud", 120, 20000, 48000, 8192, 1234,
                          dict(_MB, spectral_fx_mode="bitcrush", spectral_fx_strength=0.5)),
    "nfft8192_dispersal": ("bass", 121, 20000, 48000, 8192, 1234,
                           dict(_MB, spectral_fx_mode="phase_dispersal", spectral_fx_strength=0.6)),
    "nfft8192_scramble_pick": ("loud", 122, 20000, 48000, 8192, 1234,
                               dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.55)),
    "nfft8192_scramble_swap": ("bass", 123, 20000, 48000, 8192, 1234,
                               dict(_MB, spectral_fx_mode="bin_scramble", spectral_fx_strength=0.3)),
    "nfft8192_freeze": ("bass", 124, 20000, 48000, 8192, None, {"spectral_freeze": True}),
    "nfft8192_formant": ("bass", 125, 20000, 48000, 8192, None, {"formant_shift": 3.0}),
    "nfft8192_multiband": ("loud", 126, 20000, 48000, 8192, None, {"use_multiband": True}),
}
# cases of CASES / CASES_R2 that are ALSO rendered with precision="float64" against the same fixture
F64_CASES = ("mb_growl_bitcrush", "mb_growl_dispersal", "mb_growl_scramble_pick", "nfft4096_bitcrush",
             "nfft4096_dispersal", "nfft4096_scramble_pick", "nfft4096_scramble_swap", "nfft4096_formant",
             "nfft512_scramble_pick")

# quantize_mode left at the reference default ("autotune_v1") with snap_strength = 0 inside a multiband render:
# the only combination that keeps both (dsp/pipeline.py:1326-1327, :1076, :537-601)
AT_MB_CASES = {
    "at_mb_snap0": ("loud", 130, N, 48000, {"snap_strength": 0.0, "use_multiband": True, "lowband_drive": 1.7}),
    "at_mb_snap0_tube": ("bass", 131, 11025, 44100, dict(CLANG, snap_strength=0.0, use_multiband=True, crossover_hz=180.0,
                                                          dry_wet=0.6, output_trim_db=-1.5, delta_listen=True)),
}

# the reference's own audio files (BASELINE configs[0]): name -> path below /root/reference
REF_WAVS = {
    "example_bass": "examples/example_bass.wav",
    "sub_sweep": "tests/data/sub_sweep.wav",
    "wobble_bass": "tests/data/wobble_bass.wav",
    "kick_sub_combo": "tests/data/kick_sub_combo.wav",
    "midrange_growl_like": "tests/data/midrange_growl_like.wav",
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
# UI dicts that keep quantize_mode="autotune_v1" (no high_band / quantum_fx section, so no FX switches the mode):
# the quantization.* sub-layer keys must reach the autotune render (dsp/pipeline.py:964-977)
UI_AT_CASES = {
    "ui_at_sub_keys": ("tone", 53, AT_N, 48000, None,
                       {"quantization": {"mode": "autotune_v1", "key": "G", "scale": "major", "sub_enabled": True,
                                         "sub_source": "scale_degree", "sub_scale_degree": 2, "sub_octave": 1,
                                         "sub_level": 0.6, "sub_cut_hz": 95.0, "air_cut_hz": 7000.0, "air_mix": 0.4},
                        "crossover_freq": 200.0}, {}),
    "ui_at_manual_sub": ("tone", 54, AT_N, 48000, None,
                         {"quantization": {"sub_source": "manual", "sub_note": "A", "sub_level": 0.8},
                          "low_band": {"saturation_amount": 0.2}}, {"snap_strength": 0.7}),
}
