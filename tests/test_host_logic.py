"""CPU: host-side logic of the product (tables, parameter resolution, error behaviour, C ABI surface)."""
import ctypes
import os
import re

import numpy as np
import pytest

from quantumdistortion_b200 import PipelineConfig, get_preset, list_presets, tables
from quantumdistortion_b200 import _lib as qlib
from quantumdistortion_b200 import build as qbuild
from quantumdistortion_b200.pipeline import _resolve_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_integer_tables_bit_exact_vs_reference_fixtures():
    t = np.load(os.path.join(G, "tables.npz"))
    n = 0
    for k in t.files:
        if k.startswith("tb/") and k != "tb/kat4":
            sr, n_fft, key, scale = k[3:].split("_", 3)
            freqs = np.fft.rfftfreq(int(n_fft), d=1.0 / int(sr))
            assert np.array_equal(tables.build_target_bins_for_freqs(freqs, key, scale), t[k]), k
            assert np.array_equal(tables.build_quantize_band_mask(freqs, 110.0, 5000.0), t["mask/" + k[3:]]), k
            n += 1
    assert n == 10
    freqs = np.fft.rfftfreq(2048, d=1.0 / 48000)
    assert np.array_equal(tables.build_quantize_band_mask(freqs, 0.0, 0.0), t["mask/wide"])
    for f0 in (55.0, 110.0, 441.3):
        assert np.array_equal(tables.build_harmonic_target_bins(freqs, f0), t[f"htb/{f0}"])
    assert list(tables.build_target_bins_for_freqs(np.array([0.0, 430.0, 440.0, 450.0]), "A", "minor")) == [0, 2, 2, 2]


def test_default_table_structure_matches_survey():
    freqs = np.fft.rfftfreq(2048, d=1.0 / 48000)
    tb = tables.build_target_bins_for_freqs(freqs, "D", "minor")
    mask = tables.build_quantize_band_mask(freqs, 110.0, 5000.0)
    assert mask.sum() == 209 and mask[5] and mask[213] and not mask[4] and not mask[214]
    assert len(np.unique(tb[mask])) == 36 and np.max(np.abs(tb[mask] - np.nonzero(mask)[0])) == 13
    assert np.max(np.bincount(tb[mask])) == 27


def test_constants_half_even_rounding_and_sos():
    assert tables.limiter_constants(44100, -1.0)[1] == 220       # round(220.5) -> 220
    assert tables.limiter_constants(48000, -1.0)[1] == 240
    c = tables.limiter_constants(44100, -1.0)[2]
    assert c == float(np.exp(-1.0 / 1323))
    lo, hi = tables.design_linkwitz_riley_sos(48000, 300.0)
    assert lo.shape == (2, 6) and hi.shape == (2, 6)
    assert abs(lo[0, 4] + 1.9444776577670935) < 1e-15 and abs(hi[0, 0] - 0.9726138984998438) < 1e-15
    st = np.load(os.path.join(G, "stages.npz"))
    assert np.array_equal(lo, st["td/xo/sos_low"]) and np.array_equal(hi, st["td/xo/sos_high"])


def test_reference_error_behaviour_before_launch():
    base = dict(sr=48000, n_samples=1000, key="D", scale="minor", snap_strength=1.0, smear=0.1, bin_smoothing=True,
                pre_quant=True, post_quant=True, distortion_mode="wavefold", distortion_params={}, limiter_on=True,
                limiter_ceiling_db=-1.0, dry_wet=1.0, use_multiband=False, crossover_hz=300.0, lowband_drive=1.0,
                passthrough_test=False, harmonic_lock_hz=0.0, delta_listen=False, mono_strength=1.0,
                output_trim_db=0.0, low_trim_db=0.0, sub_cut_hz=110.0, air_cut_hz=5000.0)
    tables.resolve(**base)
    with pytest.raises(ValueError, match="Unsupported key name"):
        tables.resolve(**dict(base, key="H"))
    with pytest.raises(KeyError):
        tables.resolve(**dict(base, scale="lydian"))
    with pytest.raises(ValueError, match="Unsupported distortion mode"):
        tables.resolve(**dict(base, distortion_mode="fuzz"))
    with pytest.raises(ValueError, match="Crossover frequency"):
        tables.resolve(**dict(base, use_multiband=True, crossover_hz=24000.0))
    with pytest.raises(ValueError):
        tables.resolve(**dict(base, n_fft=3000))


def test_resolved_flags_follow_reference_gates():
    SB = {"quantize_mode": "spectral_bins"}
    r, _ = _resolve_kwargs(1000, 48000, 2048, dict(SB, snap_strength=0.0))
    assert r.params.pre_quant == 0 and r.params.post_quant == 0 and r.tables is None and r.params.no_spectral == 0
    r, _ = _resolve_kwargs(1000, 48000, 2048, dict(SB, dry_wet=1.7, output_trim_db=-6.0, snap_strength=2.0))
    assert r.params.wet == 1.0 and r.params.dry == 0.0 and r.params.apply_trim == 1
    assert r.tables.snap == 1.0 and r.params.pre_quant == 1
    assert abs(r.params.trim_gain - np.float32(10 ** (-6 / 20))) == 0.0
    pc = PipelineConfig.from_preset("Subtle Tube Glue")
    assert pc.quantize_mode == "autotune_v1" and PipelineConfig().quantize_mode == "autotune_v1"   # config.py:77, :140
    from quantumdistortion_b200 import autotune as host_at
    assert isinstance(_resolve_kwargs(1000, 44100, 2048, {"pipeline_config": pc})[0], host_at.QdAutotuneParams)
    assert isinstance(_resolve_kwargs(1000, 44100, 2048, {})[0], host_at.QdAutotuneParams)          # bare call: config.py:44
    pc.quantize_mode = "spectral_bins"
    r, _ = _resolve_kwargs(1000, 44100, 2048, {"pipeline_config": pc})
    assert r.params.post_quant == 0 and r.params.distortion_mode == 1 and abs(r.params.wet - np.float32(0.7)) == 0
    assert r.params.lookahead == 220
    with pytest.raises(NotImplementedError):
        _resolve_kwargs(1000, 48000, 2048, {"quantize_mode": "granular_v9"})
    r, _ = _resolve_kwargs(1000, 48000, 2048, {"quantize_mode": "autotune_v1", "harmonic_lock_hz": 55.0})
    assert r.params.pre_quant == 1          # an FX / freeze / formant / lock option flips the mode (:1315-1324)
    with pytest.raises(TypeError):
        _resolve_kwargs(1000, 48000, 2048, {"bogus": 1})
    # formant shift (dsp/pipeline.py:306-310): ratio 2^(st/12), flips autotune_v1 to the STFT path; precision="auto" takes
    # the float64 kernels (bins under the float32 FFT's noise floor would reach the cepstral envelope as noise)
    r, _ = _resolve_kwargs(1000, 48000, 2048, {"quantize_mode": "autotune_v1", "formant_shift": 12.0})
    assert r.params.formant_ratio == 2.0 and r.params.formant_order == 30 and r.params.precision == 1
    r, _ = _resolve_kwargs(1000, 48000, 2048, {"formant_shift": 12.0, "precision": "float32"})
    assert r.params.precision == 0
    r, _ = _resolve_kwargs(1000, 48000, 2048, {"formant_shift": 3.0, "snap_strength": 0.0})   # formant flips the mode
    assert r.params.formant_ratio == 0.0   # the spectral stage does not run at all (:635, :728)
    r, _ = _resolve_kwargs(1000, 48000, 8192, {"formant_shift": 3.0})
    assert r.params.formant_ratio > 1.0 and r.params.precision == 1   # n_fft 8192: always the float64 kernels


def test_autotune_params_follow_reference_rules():
    """quantize_mode="autotune_v1" (dsp/pipeline.py:537-601, dsp/autotune.py): host-resolved numbers and struct layout."""
    import subprocess
    import tempfile
    from quantumdistortion_b200 import autotune as host_at
    from oracle import qd_autotune as at
    p, _ = _resolve_kwargs(24000, 48000, 2048, {"quantize_mode": "autotune_v1", "key": "F", "scale": "dorian",
                                                 "snap_strength": 1.7, "sub_source": "scale_degree", "sub_scale_degree": 4,
                                                 "sub_octave": 1, "use_multiband": True})
    assert isinstance(p, host_at.QdAutotuneParams) and p.apply == 1 and p.strength == 1.0
    assert (p.min_tau, p.max_tau, p.frame_size, p.hop) == (16, 671, 4096, 512)
    assert (p.max_delay, p.buffer_size) == (1024, 4096) and p.root_pc == 5 and list(p.intervals[:7]) == [0, 2, 3, 5, 7, 9, 10]
    cfg = at.AutotuneConfig(key="F", scale="dorian", sub_source="scale_degree", sub_scale_degree=4, sub_octave=1)
    assert p.phase_k == float(np.float32(2.0 * np.pi * at.sub_frequency(cfg)))
    sos = at.butter_sos(48000, 110.0, "low")
    assert [p.sos[0][1][c] for c in range(6)] == list(sos[1])
    p44, _ = _resolve_kwargs(100, 44100, 2048, {"quantize_mode": "autotune_v1", "sub_cut_hz": 0.0})
    assert p44.filt_on[0] == 0 and p44.filt_on[1] == 1 and p44.max_tau == int(44100 / 71.5) and p44.lookahead == 220
    with pytest.raises(ValueError):     # scipy's sosfiltfilt: input shorter than the padding
        _resolve_kwargs(12, 48000, 2048, {"quantize_mode": "autotune_v1"})
    # the one combination that keeps autotune_v1 inside a multiband render (:1326-1327): pitch stage gated off, no STFT
    r, _ = _resolve_kwargs(1000, 48000, 2048, {"snap_strength": 0.0, "use_multiband": True})
    assert r.params.no_spectral == 1 and r.params.multiband == 1 and r.params.pre_quant == 0 and r.tables is None
    r, _ = _resolve_kwargs(1000, 48000, 2048, {"quantize_mode": "autotune_v1", "passthrough_test": True})
    assert r.params.passthrough == 1     # the passthrough branch comes first (:477)
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "qd_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", '
           'sizeof(qd_autotune_params), offsetof(qd_autotune_params, zi), offsetof(qd_autotune_params, strength), '
           'offsetof(qd_autotune_params, phase_k), offsetof(qd_autotune_params, ceiling_lin));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    P = host_at.QdAutotuneParams
    assert got == [ctypes.sizeof(P), P.zi.offset, P.strength.offset, P.phase_k.offset, P.ceiling_lin.offset]


def test_presets_mirror_reference():
    assert list_presets() == ["Chordal Noise Wash", "Controlled Dubstep Growl", "Perc To Tonal Clang",
                              "Subtle Tube Glue"]
    g = get_preset("Controlled Dubstep Growl")
    assert g["key"] == "F" and g["snap_strength"] == 0.9 and g["distortion_params"]["fold_amount"] == 5.0
    with pytest.raises(KeyError):
        get_preset("nope")
    pc = PipelineConfig()
    assert pc.key == "D" and pc.sub_cut_hz == 110.0 and pc.air_cut_hz == 5000.0 and pc.crossover_hz == 300.0


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    """No compute call here: only dlopen + symbol lookup (no GPU needed)."""
    lib_path = qbuild.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "qd_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", header))
    assert {"qd_plan_create", "qd_render_device", "qd_render_host", "qd_limiter_device", "qd_crossover_device",
            "qd_distort_device", "qd_plan_read_timing"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/qd_b200.h but not exported"
    lib.qd_abi_version.restype = ctypes.c_int
    assert lib.qd_abi_version() == qlib.QD_ABI_VERSION
    assert ctypes.sizeof(qlib.QdParams) % 8 == 0


def test_params_struct_layout_matches_header():
    """ctypes mirror vs the C compiler's view of qd_params (offsets of a few late fields)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "qd_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(qd_params), offsetof(qd_params, ceiling_lin), offsetof(qd_params, sos_low), offsetof(qd_params, low_norm), offsetof(qd_params, fx_a));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    P = qlib.QdParams
    assert got == [ctypes.sizeof(P), P.ceiling_lin.offset, P.sos_low.offset, P.low_norm.offset, P.fx_a.offset]


def test_ui_config_dict_mapping_matches_reference_rules():
    from quantumdistortion_b200.pipeline import _parse_ui_config
    out = _parse_ui_config({"high_band": {}}, {})
    assert out["use_multiband"] is True and out["spectral_fx_mode"] == "bin_scramble" and out["spectral_fx_strength"] == 0.2
    out = _parse_ui_config({"high_band": {"bin_scrambling": 0.0}}, {})
    assert out["spectral_fx_mode"] == "phase_dispersal" and out["spectral_fx_strength"] == 0.3
    out = _parse_ui_config({"high_band": {"bin_scrambling": 0.0, "phase_dispersal": 0.0, "output_trim_db": -3.0}}, {})
    assert out["spectral_fx_mode"] == "bitcrush" and out["spectral_fx_strength"] == 0.5 and out["output_trim_db"] == -3.0
    out = _parse_ui_config({"low_band": {"saturation_amount": 0.5}, "crossover_freq": 250}, {"mono_strength": 0.9})
    assert out["lowband_drive"] == 3.0 and out["mono_strength"] == 0.9 and out["crossover_hz"] == 250.0
    r, _ = _resolve_kwargs(12000, 48000, 2048, {"config": {"quantization": {"key": "E", "scale": "dorian"},
                                                          "high_band": {"bin_scrambling": 0.0, "phase_dispersal": 0.0}}})
    assert r.params.multiband == 1 and r.params.fx_mode == qlib.QD_FX["bitcrush_log"]


def test_wav_io_roundtrip_and_pcm_rules(tmp_path):
    from quantumdistortion_b200.audio_io import float_to_pcm16, load_audio, save_audio
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 1.5, 0.25], dtype=np.float32)
    pcm = float_to_pcm16(x)
    assert list(pcm) == [0, 16384, -16384, 32767, -32767, 32767, 8192]   # lrint(x * 0x7FFF) half-to-even, clipped
    p = tmp_path / "a" / "t.wav"
    save_audio(p, x, 44100)
    y, sr = load_audio(p)
    assert sr == 44100 and y.dtype == np.float32 and np.array_equal(y, pcm.astype(np.float32) / 32768.0)
    from quantumdistortion_b200 import harness
    with pytest.raises(FileNotFoundError):
        harness.process_file_to_file(tmp_path / "missing.wav", tmp_path / "o.wav")
    cfg = harness._config("Subtle Tube Glue", {"dry_wet": 0.25, "not_a_field": 1})
    assert cfg.dry_wet == 0.25 and cfg.distortion_mode == "tube" and not hasattr(cfg, "not_a_field")
    with pytest.raises(KeyError):
        harness._config("nope", None)


def test_streamlit_default_dict_raises_like_the_reference():
    """ui/app_streamlit.py:74-163 with harmonic lock "Off" hands over quantum_fx.fundamental_hz = None, and the
    reference's parser calls float() on it (dsp/pipeline.py:1003): the reference raises TypeError on its own UI's
    default dict (checked against the live reference).  Same exception here, before any launch; a numeric
    fundamental (lock mode "G1" = 49.0 Hz) resolves to the STFT path."""
    from quantumdistortion_b200.pipeline import _resolve_kwargs
    cfg = {"quantization": {"key": "D", "scale": "minor", "mode": "autotune_v1"}, "crossover_freq": 300,
           "low_band": {"saturation_amount": 0.3, "saturation_type": "Tube", "mono_strength": 1.0, "output_trim_db": 0},
           "high_band": {"fft_size": 2048, "window_type": "hann", "precision_mode": "Quantized", "mag_decimation": 0.5,
                         "phase_dispersal": 0.3, "bin_scrambling": 0.2, "output_trim_db": 0},
           "quantum_fx": {"spectral_freeze": False, "formant_shift": 0, "harmonic_lock_mode": "Off", "fundamental_hz": None},
           "delta_listen": False}
    with pytest.raises(TypeError):
        _resolve_kwargs(4800, 48000, 2048, dict(config=cfg))
    cfg["quantum_fx"].update(harmonic_lock_mode="G1", fundamental_hz=49.0)
    r, _ = _resolve_kwargs(4800, 48000, 2048, dict(config=cfg))
    assert r.params.multiband == 1 and r.params.fx_mode != 0   # bin_scrambling 0.2 wins (:986-994), mode flipped (:1315)
