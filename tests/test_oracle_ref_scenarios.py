"""CPU: the oracle on the reference's own pipeline-level test scenarios (qd_cases.REF_SCENARIOS; fixtures
tests/golden/scenarios.npz from the LIVE reference, generator tests/golden/make_golden.py scenarios).

The signals and keyword arguments are the ones the reference's test suite passes to process_audio
(tests/test_pipeline.py, test_passthrough_null.py, test_multiband_alignment.py, test_pipeline_multiband_identity.py,
test_m12_quantum_fx.py, test_quantization_integration.py); quantize_mode stays at the reference default wherever
the reference test leaves it there, so these cases also pin the mode-resolution rules of dsp/pipeline.py:1315-1327.
"""
import os

import numpy as np
import pytest

import qd_cases
from oracle import qd_autotune as at
from oracle import qd_oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_AT_DROP = ("smear", "bin_smoothing", "post_quant", "use_multiband", "crossover_hz", "lowband_drive", "quantize_mode",
            # options the reference test passes at their "off" value: they do not leave autotune_v1 (:1315-1324)
            "spectral_fx_mode", "spectral_fx_strength", "spectral_freeze", "formant_shift", "harmonic_lock_hz")


@pytest.fixture(scope="module")
def sc():
    return np.load(os.path.join(G, "scenarios.npz"))


def oracle_render(x, sr, kw):
    """The two oracles behind the reference's dispatch: qd_oracle.process_audio resolves the mode like
    dsp/pipeline.py:1315-1327 (FX / freeze / formant / lock options and passthrough_test run the STFT path, autotune_v1
    with snap_strength > 0 drops multiband) and refuses a render that stays in the time-domain autotune_v1 chain,
    which is oracle/qd_autotune.py."""
    kw = dict(kw)
    mode = kw.pop("quantize_mode", "autotune_v1")
    try:
        return orc.process_audio(x, sr, quantize_mode=mode, **{k: v for k, v in kw.items() if k != "sub_enabled"})
    except NotImplementedError:
        return at.process_audio_autotune(x, sr, **{k: v for k, v in kw.items() if k not in _AT_DROP})


@pytest.mark.parametrize("name", list(qd_cases.REF_SCENARIOS))
def test_oracle_on_reference_test_scenarios(sc, name):
    spec, sr, rng_seed, kw, prop, cite = qd_cases.REF_SCENARIOS[name]
    x = qd_cases.scenario_signal(spec, sr)
    assert np.array_equal(x, sc[f"{name}/x"]), "scenario signal drifted from the fixture"
    if rng_seed is not None:
        np.random.seed(rng_seed)
    y, taps = oracle_render(x, sr, kw)
    for got, key in ((y, "y"), (taps["pre_quant"], "pre_quant"), (taps["post_dist"], "post_dist")):
        ref = sc[f"{name}/{key}"]
        got = np.asarray(got, dtype=np.float32)
        assert got.shape == ref.shape, f"{name}/{key} ({cite})"
        err = float(np.max(np.abs(got.astype(np.float64) - ref))) if ref.size else 0.0
        assert err <= 2e-7, f"{name}/{key} ({cite}): max abs err {err:.3e}"


def test_reference_assertions_hold_on_the_fixtures(sc):
    """The fixtures themselves satisfy what the reference's tests assert (a guard for the generator)."""
    for name in qd_cases.REF_SCENARIOS:
        qd_cases.check_scenario_property(name, sc[f"{name}/x"], sc[f"{name}/y"], lambda other: sc[f"{other}/y"])
