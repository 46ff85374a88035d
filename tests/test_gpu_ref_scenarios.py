"""GPU (B200): the reference's own pipeline-level test scenarios through the drop-in ``process_audio``.

Every process_audio call of the reference's test suite (qd_cases.REF_SCENARIOS cites file:line) with the reference's
signals and keyword arguments; quantize_mode stays at the reference default wherever the reference test leaves it
there.  Each case is checked twice: against the LIVE reference's output and taps (tests/golden/scenarios.npz,
generator tests/golden/make_golden.py scenarios) at the north-star tolerance, and against the assertion the
reference's test makes (qd_cases.check_scenario_property), evaluated on OUR output.
"""
import os

import numpy as np
import pytest

import qd_cases
from oracle import qd_oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAX_ABS = 1e-4
NULL_DB = -80.0


@pytest.fixture(scope="module")
def qd():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import quantumdistortion_b200 as q
    from quantumdistortion_b200 import _lib
    assert _lib.load().qd_device_count() >= 1, "libqd_b200.so sees no sm_100 device"
    return q


@pytest.fixture(scope="module")
def sc():
    return np.load(os.path.join(G, "scenarios.npz"))


@pytest.fixture(scope="module")
def rendered(qd, sc):
    """Our output of every scenario (several reference assertions compare two renders)."""
    out = {}
    for name, (spec, sr, rng_seed, kw, prop, cite) in qd_cases.REF_SCENARIOS.items():
        if rng_seed is not None:
            np.random.seed(rng_seed)   # the random spectral FX replay the global np.random state like the reference
        out[name] = qd.process_audio(sc[f"{name}/x"], sr=sr, **kw)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(qd_cases.REF_SCENARIOS))
def test_reference_test_scenarios(rendered, sc, name):
    spec, sr, rng_seed, kw, prop, cite = qd_cases.REF_SCENARIOS[name]
    x = sc[f"{name}/x"]
    y, taps = rendered[name]
    assert set(taps.keys()) == {"input", "pre_quant", "post_dist", "output"}, cite
    assert np.array_equal(taps["input"], x) and np.array_equal(taps["output"], y)
    for got, key in ((y, "y"), (taps["pre_quant"], "pre_quant"), (taps["post_dist"], "post_dist")):
        ref = sc[f"{name}/{key}"]
        what = f"{name}/{key} ({cite})"
        assert got.dtype == np.float32 and got.shape == ref.shape == x.shape, what
        err = float(np.max(np.abs(got.astype(np.float64) - ref)))
        if key == "y" and name in qd_cases.REF_SCENARIO_Y_TOL:   # the reference itself is ill-conditioned there
            assert err <= qd_cases.REF_SCENARIO_Y_TOL[name], f"{what}: max abs err {err:.3e}"
            continue
        assert err <= MAX_ABS, f"{what}: max abs err {err:.3e}"
        assert orc.null_test_db(got, ref) <= NULL_DB, what
    qd_cases.check_scenario_property(name, x, y, lambda other: rendered[other][0])
