"""CPU: the autotune_v1 oracle (oracle/qd_autotune.py) against outputs of the LIVE reference stored in
tests/golden/autotune.npz (tests/golden/make_golden.py autotune).  Bit-exact on every stage and on whole renders."""
import os

import numpy as np
import pytest
import scipy.signal

import qd_cases
from oracle import qd_autotune as at

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "autotune.npz"))


def test_band_split_and_detector(gold):
    x = gold["st/x"]
    sub, body, air = at.split_sub_body_air(x, 48000, 110.0, 5000.0)
    assert np.array_equal(sub, gold["st/sub"]) and np.array_equal(body, gold["st/body"]) and np.array_equal(air, gold["st/air"])
    assert np.array_equal(at.detector_sidechain(body, 48000, 110.0, 3000.0), gold["st/det"])


def test_sosfiltfilt_restatement_is_scipy():
    """The sample-level specification the CUDA kernel follows equals scipy.signal.sosfiltfilt."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal(700).astype(np.float32)
    for cutoff, bt in ((110.0, "low"), (5000.0, "high"), (3000.0, "low")):
        sos = at.butter_sos(48000, cutoff, bt)
        assert np.array_equal(at.sosfiltfilt_restated(sos, x), scipy.signal.sosfiltfilt(sos, x))
    with pytest.raises(ValueError):
        at.sosfiltfilt_restated(at.butter_sos(48000, 110.0, "low"), np.zeros(15))


def test_frame_features_and_ratio_track(gold):
    cfg = at.AutotuneConfig()
    centers, rms, flat, pitch, conf = at.frame_features(gold["st/det"], 48000, cfg)
    f = gold["st/features"]
    assert np.array_equal(rms, f[:, 0]) and np.array_equal(flat, f[:, 1])
    assert np.array_equal(pitch, f[:, 2]) and np.array_equal(conf, f[:, 3])
    assert np.array_equal(at.ratio_track(gold["st/det"], 48000, cfg), gold["st/ratio_track"])
    got = np.array([at.nearest_scale_freq(v, "D", "minor") for v in (97.3, 233.1, 440.0, 1234.5)])
    assert np.array_equal(got, gold["st/nearest"])


def test_granular_shifter(gold):
    assert np.array_equal(at.granular_pitch_shift(gold["st/body"], gold["st/ratio_track"]), gold["st/corrected"])
    assert np.array_equal(at.granular_pitch_shift(gold["st/body"], gold["st/ratio_manual"]), gold["st/shift_manual"])
    ones = np.ones(100, dtype=np.float32)
    assert np.array_equal(at.granular_pitch_shift(gold["st/body"][:100], ones), gold["st/body"][:100])  # early out, no latency


def test_sub_layer_and_mode_output(gold):
    x = gold["st/x"]
    assert np.array_equal(at.envelope_follow(x, 48000), gold["st/env"])
    assert np.array_equal(at.sub_layer(x, 48000, at.AutotuneConfig()), gold["st/sub_layer"])
    assert np.array_equal(at.apply_autotune_v1(x, 48000, at.AutotuneConfig())["output"], gold["st/output"])


@pytest.mark.parametrize("name", list(qd_cases.AUTOTUNE_CASES))
def test_autotune_pipeline_matches_reference(gold, name):
    kind, seed, n, sr, kw = qd_cases.AUTOTUNE_CASES[name]
    x = qd_cases.make_signal(kind, seed, n, sr)
    assert np.array_equal(x, gold[f"{name}/x"])
    y, taps = at.process_audio_autotune(x, sr, **{k: v for k, v in kw.items()
                                                  if k not in ("smear", "bin_smoothing", "post_quant")})
    for got, key in ((y, "y"), (taps["pre_quant"], "pre_quant"), (taps["post_dist"], "post_dist")):
        ref = gold[f"{name}/{key}"]
        assert np.array_equal(got, ref), f"{name}/{key}: max abs diff {np.max(np.abs(got.astype(float) - ref)):.3e}"
