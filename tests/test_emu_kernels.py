"""CPU: single-step the CUDA kernel SOURCE through tests/emu/cuda_emu.h (g++ -DQD_EMU) and compare
with the oracle.  This checks index math / table layout / branch shapes of
quantumdistortion_b200/csrc/qd_spec.cuh in the GPU-less container; it is not the product path
(the shipped library is the nvcc build of the same source) and proves nothing about speed."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import qd_oracle as orc
from quantumdistortion_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(os.path.dirname(HERE), "quantumdistortion_b200", "csrc")


def _build_emu(exe, srcs, main):
    """Compile an emulator once per source change; link under a private name and rename, so that pytest-xdist workers
    (or two suites on one box) never execute a half-written binary."""
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        tmp = f"{exe}.{os.getpid()}"
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-DQD_EMU", "-I", EMU_DIR, "-o", tmp, main, "-lpthread"])
        os.replace(tmp, exe)
    return exe


@pytest.fixture(scope="module")
def emu_spec():
    exe = os.path.join(tempfile.gettempdir(), "qd_emu_spec")
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_spec.cpp", "cuda_emu.h")]
    srcs += [os.path.join(CSRC, f) for f in ("qd_spec.cuh", "qd_spec_team.cuh", "qd_common.cuh", "qd_host_tables.hpp")]
    return _build_emu(exe, srcs, os.path.join(EMU_DIR, "emu_spec.cpp"))


def run_emu(exe, x, sr, n_fft, nw, tile_blocks, quant, smoothing, snap, smear, epilogue=0, fold=1.0,
            bias=0.0, tg=1.0, tn=1.0, key="D", scale="minor", lo=110.0, hi=5000.0, fx=None, prec="f32",
            formant_ratio=None):
    freqs = np.fft.rfftfreq(n_fft, d=1.0 / sr)
    tb = orc.target_bins_for_freqs(freqs, key, scale).astype(np.int32)
    mask = orc.quantize_band_mask(freqs, lo, hi).astype(np.uint8)
    with tempfile.TemporaryDirectory() as d:
        p = lambda n: os.path.join(d, n)
        x.astype(np.float32).tofile(p("x"))
        tb.tofile(p("tb"))
        mask.tofile(p("mask"))
        n_per_clip = x.shape[-1]   # x may be [clips, n]: the grouped kernel renders several clips per CTA
        cmd = [exe, prec, str(n_fft), str(nw), str(n_per_clip), str(tile_blocks), str(int(quant)),
               str(int(smoothing)), repr(float(snap)), repr(float(smear)), str(epilogue),
               repr(float(fold)), repr(float(bias)), repr(float(tg)), repr(float(tn)),
               p("x"), p("tb"), p("mask"), p("y"), p("tap")]
        if fx is not None:  # (fx_mode, a, b, c, table ndarray or None, pass index)
            tab = "-"
            if fx[4] is not None:
                fx[4].tofile(p("fxt"))
                tab = p("fxt")
            cmd += [str(fx[0]), repr(float(fx[1])), repr(float(fx[2])), repr(float(fx[3])), tab, str(fx[5])]
        env = dict(os.environ)
        env.pop("QD_EMU_FORMANT_RATIO", None)
        if formant_ratio is not None:
            env["QD_EMU_FORMANT_RATIO"] = repr(float(formant_ratio))
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, env=env)
        return np.fromfile(p("y"), dtype=np.float32), np.fromfile(p("tap"), dtype=np.float32)


def oracle_pass(x, sr, n_fft, quant, smoothing, snap, smear, key="D", scale="minor", lo=110.0, hi=5000.0):
    S, freqs = orc.stft(x, sr, n_fft)
    if quant:
        S = orc.spectral_quantize_stft(S, freqs, key, scale, snap, smear, smoothing,
                                       quantize_min_hz=lo, quantize_max_hz=hi)
    return orc.istft(S, sr, n_fft, length=len(x))


@pytest.mark.parametrize("n_fft,nw,n,tile", [(2048, 4, 6000, 64), (2048, 8, 5003, 5), (512, 4, 1500, 7),
                                             (1024, 4, 2100, 64), (4096, 4, 9000, 3), (8192, 4, 17000, 64),
                                             (2048, 4, 700, 64)])
def test_emu_passthrough(emu_spec, n_fft, nw, n, tile):
    x = synth.bass_clip(3, n)
    y, tap = run_emu(emu_spec, x, 48000, n_fft, nw, tile, False, False, 1.0, 0.1)
    ref = oracle_pass(x, 48000, n_fft, False, False, 1.0, 0.1)
    assert np.max(np.abs(y - ref)) < 2e-6
    assert np.array_equal(y, tap)
    assert orc.null_test_db(y, x) < -110.0


@pytest.mark.parametrize("n_fft,nw,n,tile,snap,smear,smooth", [
    (2048, 4, 6000, 64, 1.0, 0.1, True), (2048, 8, 5003, 4, 0.9, 0.3, True), (2048, 16, 9000, 64, 1.0, 0.1, True),
    (2048, 4, 24000, 64, 1.0, 0.1, True),   # long enough for interior batches: staged by the bulk-copy path
    (2048, 4, 24000, 13, 1.0, 0.1, True),   # the same cut into tiles
    (512, 4, 24002, 64, 1.0, 0.1, True),    # n % 4 != 0: bulk copies are never used (2048, 4, 4000, 64, 0.75, 0.0, False),
    (512, 4, 1500, 64, 1.0, 0.1, True), (1024, 4, 2100, 6, 1.0, 0.1, True), (4096, 4, 9000, 64, 1.0, 0.1, True)])
def test_emu_quantized_pass(emu_spec, n_fft, nw, n, tile, snap, smear, smooth):
    x = synth.noise_clip(5, n)
    y, _ = run_emu(emu_spec, x, 48000, n_fft, nw, tile, True, smooth, snap, smear)
    ref = oracle_pass(x, 48000, n_fft, True, smooth, snap, smear)
    err = float(np.max(np.abs(y - ref)))
    assert err < 1e-5, err


def test_emu_wide_mask_and_epilogue(emu_spec):
    x = synth.loud_clip(6, 5000)
    y, tap = run_emu(emu_spec, x, 48000, 2048, 4, 64, True, True, 0.9, 0.3, epilogue=1, fold=5.0, bias=0.1,
                     key="A", scale="pentatonic", lo=0.0, hi=0.0)
    ref_tap = oracle_pass(x, 48000, 2048, True, True, 0.9, 0.3, key="A", scale="pentatonic", lo=0.0, hi=0.0)
    assert np.max(np.abs(tap - ref_tap)) < 1e-5
    ref = orc.apply_distortion(tap, "wavefold", fold_amount=5.0, bias=0.1)
    assert np.max(np.abs(y - ref)) < 1e-5


@pytest.fixture(scope="module")
def emu_time():
    exe = os.path.join(tempfile.gettempdir(), "qd_emu_time")
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_time.cpp", "cuda_emu.h")]
    srcs += [os.path.join(CSRC, f) for f in ("qd_time.cuh", "qd_common.cuh", "qd_host_time.hpp")]
    return _build_emu(exe, srcs, os.path.join(EMU_DIR, "emu_time.cpp"))


@pytest.mark.parametrize("n,sr,kind", [(6000, 48000, "loud"), (2048, 44100, "noise"), (5000, 48000, "noise"),
                                       (300, 48000, "noise"), (4097, 1000, "noise")])
def test_emu_limiter(emu_time, n, sr, kind):
    x = (synth.loud_clip(1, n) if kind == "loud" else 2.0 * synth.noise_clip(2, n)).astype(np.float32)
    ceiling, L, c = orc.limiter_constants(sr, -1.0, 5.0, 30.0)
    with tempfile.TemporaryDirectory() as d:
        x.tofile(os.path.join(d, "x"))
        subprocess.check_call([emu_time, "limiter", str(n), str(L), repr(ceiling), repr(c),
                               os.path.join(d, "x"), os.path.join(d, "y")])
        y = np.fromfile(os.path.join(d, "y"), dtype=np.float32)
    ref, _ = orc.peak_limiter(x, sr, -1.0, 5.0, 30.0)
    assert np.max(np.abs(ref)) > 0 and np.max(np.abs(ref - x)) > 1e-3, "limiter must engage in this test"
    assert np.max(np.abs(y.astype(np.float64) - ref)) <= 6e-8, float(np.max(np.abs(y - ref)))


@pytest.mark.parametrize("n,delay,tiling", [(6000, 0, None), (5003, 1024, None), (100, 0, None), (2048, 300, None),
                                            (6000, 1024, (256, 1632)), (5003, 0, (512, 2048)), (9000, 64, (32, 1632))])
def test_emu_crossover(emu_time, n, delay, tiling):
    """`tiling` = (tile, halo) forces many tiles per clip: every tile but the first starts from a zero state `halo`
    samples early and must land on the sequential filter's output."""
    x = synth.loud_clip(3, n)
    sl, sh = orc.linkwitz_riley_sos(48000, 300.0)
    with tempfile.TemporaryDirectory() as d:
        p = lambda f: os.path.join(d, f)
        x.tofile(p("x"))
        sl.astype(np.float64).tofile(p("lp"))
        sh.astype(np.float64).tofile(p("hp"))
        subprocess.check_call([emu_time, "crossover", str(n), str(delay), p("lp"), p("hp"), p("x"), p("lo"), p("hi")] +
                              ([str(tiling[0]), str(tiling[1])] if tiling else []))
        lo = np.fromfile(p("lo"), dtype=np.float32)
        hi = np.fromfile(p("hi"), dtype=np.float32)
    rlo, rhi = orc.linkwitz_riley_split(x, 48000, 300.0)
    rlo = np.concatenate([np.zeros(delay, dtype=np.float32), rlo])[:n]
    assert np.max(np.abs(hi - rhi)) <= 1.2e-7
    assert np.max(np.abs(lo - rlo)) <= 1.2e-7


@pytest.mark.parametrize("mode,strength,seed", [("bitcrush", 0.5, None), ("bitcrush", 0.3, None),
                                                ("phase_dispersal", 0.6, 7), ("phase_dispersal", 0.3, None),
                                                ("bin_scramble", 0.55, 11), ("bin_scramble", 0.3, 12)])
def test_emu_spectral_fx_pass(emu_spec, mode, strength, seed):
    """One high-band pass with a spectral FX: kernel source (emulated) vs the oracle with the same np.random seed."""
    from quantumdistortion_b200 import tables
    n, sr, n_fft = 5000, 48000, 2048
    x = synth.loud_clip(8, n)
    fxr = tables.resolve_spectral_fx(mode, strength, {})
    tab = None
    if fxr["rng"]:
        np.random.seed(seed)
        tab = tables.replay_fx_table(fxr["rng"], 1, 1 + n // (n_fft // 4), n_fft // 2 + 1)
    nw = 12 if mode == "bin_scramble" else 8   # 12 warps per CTA is the production FX variant for n_fft 2048
    y, _ = run_emu(emu_spec, x, sr, n_fft, nw, 64, True, True, 0.9, 0.3, key="F",
                   fx=(fxr["fx_mode"], fxr["a"], fxr["b"], fxr["c"], tab, 0))
    S, freqs = orc.stft(x, sr, n_fft)
    if seed is not None:
        np.random.seed(seed)
    Sq = orc.spectral_quantize_stft(S, freqs, "F", "minor", 0.9, 0.3, True, is_high_band=True,
                                    spectral_fx_mode=mode, spectral_fx_strength=strength, spectral_fx_params={})
    ref = orc.istft(Sq, sr, n_fft, length=n)
    err = float(np.max(np.abs(y - ref)))
    assert err < 2e-5, err


@pytest.mark.parametrize("n_fft,nw,n", [(2048, 4, 6000), (512, 4, 1500), (8192, 2, 17000)])
def test_emu_float64_pass(emu_spec, n_fft, nw, n):
    """The float64 instantiation of the same kernel source (parity path for ill-conditioned configurations)."""
    x = synth.noise_clip(5, n)
    y, _ = run_emu(emu_spec, x, 48000, n_fft, nw, 64, True, True, 1.0, 0.1, prec="f64")
    ref = oracle_pass(x, 48000, n_fft, True, True, 1.0, 0.1)
    assert float(np.max(np.abs(y - ref))) <= 6e-8


def test_emu_float64_wide_mask_second_pass(emu_spec):
    """The case float32 cannot do: wide-open band mask, pass 2 on the distorted pass-1 signal (near-cancelling
    phasor sums with ~60 sources).  float64 reproduces the reference to float32 rounding."""
    g = np.load(os.path.join(HERE, "golden", "pipeline.npz"))
    xd = g["sb_wide_mask/post_dist"]
    kw = dict(key="A", scale="pentatonic", lo=0.0, hi=0.0)
    y64, _ = run_emu(emu_spec, xd, 48000, 2048, 4, 64, True, True, 1.0, 0.1, prec="f64", **kw)
    ref = oracle_pass(xd, 48000, 2048, True, True, 1.0, 0.1, **kw)
    assert float(np.max(np.abs(y64 - ref))) <= 2e-7


def test_emu_two_clip_groups_per_cta(emu_spec):
    """n_fft 2048 production variant: two independent 8-warp groups per CTA, one clip each, odd batch (the
    last CTA carries an empty group), clips cut into tiles."""
    n = 9000
    x = np.stack([synth.noise_clip(20 + i, n) for i in range(3)])
    y, tap = run_emu(emu_spec, x, 48000, 2048, 16, 16, True, True, 1.0, 0.1, epilogue=1, fold=2.0, bias=0.05)
    y, tap = y.reshape(3, n), tap.reshape(3, n)
    for i in range(3):
        ref_tap = oracle_pass(x[i], 48000, 2048, True, True, 1.0, 0.1)
        assert float(np.max(np.abs(tap[i] - ref_tap))) < 1e-5
        assert float(np.max(np.abs(y[i] - orc.apply_distortion(tap[i], "wavefold", fold_amount=2.0, bias=0.05)))) < 1e-5


@pytest.mark.parametrize("n_fft,nw,n,semitones", [(2048, 8, 5000, 3.0), (2048, 8, 4100, -5.0), (1024, 4, 2600, 7.0)])
def test_emu_formant_shift_pass(emu_spec, n_fft, nw, n, semitones):
    """One quantised pass with the formant shift (cepstral lifter through the warp FFT on the scratch buffer): kernel
    source (emulated) vs the oracle."""
    sr = 48000
    x = synth.bass_clip(21, n)
    y, _ = run_emu(emu_spec, x, sr, n_fft, nw, 64, True, True, 1.0, 0.1, formant_ratio=2.0 ** (semitones / 12.0))
    S, freqs = orc.stft(x, sr, n_fft)
    Sq = orc.spectral_quantize_stft(S, freqs, "D", "minor", 1.0, 0.1, True, formant_shift=semitones)
    ref = orc.istft(Sq, sr, n_fft, length=n)
    err = float(np.max(np.abs(y - ref)))
    assert err < 2e-5, err


def test_team_gather_tables_invariants(tmp_path):
    """Host side of the team kernel: qd_host::build_team_gather on random target tables for team widths 1 / 2 / 4 / 8 --
    32-aligned lists, every source exactly once and in order, each slot in one warp's list, off / tail per group of 32
    (tests/emu/team_gather_check.cpp, plain g++, no CUDA)."""
    exe = str(tmp_path / "team_gather_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(EMU_DIR, "team_gather_check.cpp")])
    assert subprocess.check_output([exe]).decode().strip() == "ok"


@pytest.mark.parametrize("n_fft,shape,n,tile,prec", [(512, 801, 1500, 7, "f32"), (512, 801, 24002, 64, "f32"), (1024, 801, 2100, 6, "f32"),
                                                     (1024, 801, 24000, 64, "f32"), (2048, 802, 6000, 64, "f64"), (2048, 802, 24000, 9, "f64"),
                                                     (4096, 704, 9000, 64, "f32"), (4096, 704, 9000, 3, "f32"),
                                                     (4096, 704, 60000, 17, "f32"),   # interior batches: bulk-copy staging
                                                     (4096, 404, 5003, 64, "f64"), (8192, 208, 17000, 64, "f64"), (8192, 404, 40000, 6, "f32"),
                                                     (8192, 208, 17001, 5, "f64")])
def test_emu_team_kernel(emu_spec, n_fft, shape, n, tile, prec):
    """qd_spec_team.cuh (a team of warps per frame, n_fft >= 4096; `shape` = 100 x frames per batch + warps per frame):
    pure STFT -> iSTFT and the quantised pass, source single-stepped on the CPU against the oracle -- butterfly groups
    dealt to the warps, per-warp gather lists, the row ranges of the paired walk with their border rows."""
    x = synth.noise_clip(5, n)
    y, tap = run_emu(emu_spec, x, 48000, n_fft, shape, tile, False, False, 1.0, 0.1, prec=prec)
    assert np.max(np.abs(y - oracle_pass(x, 48000, n_fft, False, False, 1.0, 0.1))) < 2e-6
    assert np.array_equal(y, tap)
    y, _ = run_emu(emu_spec, x, 48000, n_fft, shape, tile, True, True, 1.0, 0.1, prec=prec)
    assert np.max(np.abs(y - oracle_pass(x, 48000, n_fft, True, True, 1.0, 0.1))) < (1e-5 if prec == "f32" else 5e-7)


@pytest.mark.parametrize("n_fft,shape,prec", [(2048, 802, "f64"), (512, 801, "f32"), (1024, 801, "f32"), (4096, 704, "f32"), (4096, 404, "f64"), (8192, 208, "f64")])
def test_emu_team_kernel_wide_mask_and_epilogue(emu_spec, n_fft, shape, prec):
    """Targets that gather more than 32 sources (a slot's sources span several groups of its warp's list), no smoothing,
    and the wavefold epilogue."""
    x = synth.loud_clip(6, 9000)
    kw = dict(key="A", scale="pentatonic", lo=0.0, hi=0.0)
    y, tap = run_emu(emu_spec, x, 48000, n_fft, shape, 64, True, True, 0.9, 0.3, epilogue=1, fold=5.0, bias=0.1, prec=prec, **kw)
    ref_tap = oracle_pass(x, 48000, n_fft, True, True, 0.9, 0.3, **kw)
    assert np.max(np.abs(tap - ref_tap)) < (2e-5 if prec == "f32" else 5e-7)
    assert np.max(np.abs(y - orc.apply_distortion(tap, "wavefold", fold_amount=5.0, bias=0.1))) < 1e-5
    y, _ = run_emu(emu_spec, x, 48000, n_fft, shape, 64, True, False, 0.75, 0.0, prec=prec)
    assert np.max(np.abs(y - oracle_pass(x, 48000, n_fft, True, False, 0.75, 0.0))) < (1e-5 if prec == "f32" else 5e-7)


def test_emu_team_kernel_random_configurations(emu_spec):
    """Seeded random draws over the team kernel's instantiations: clip length (shorter than a hop ... a dozen frames, with
    and without a multiple of 4), tiling, quantizer on / off, smoothing, snap / smear, band edges, key / scale, sample rate
    and signal kind -- each against the oracle (a longer run of the same generator: 220 cases, worst error 1.0e-5 on a
    float32 wide-mask draw)."""
    rng = np.random.default_rng(2024)
    shapes = [(512, 801, "f32"), (1024, 801, "f32"), (2048, 802, "f64"), (4096, 704, "f32"), (4096, 404, "f64"), (8192, 208, "f64")]
    keys, scales = ["C", "D", "F#", "A", "Bb"], ["minor", "major", "pentatonic", "dorian", "harmonic_minor"]
    for _ in range(24):
        n_fft, shape, prec = shapes[rng.integers(len(shapes))]
        hop = n_fft // 4
        n = int(rng.choice([rng.integers(2, hop), rng.integers(hop, 3 * n_fft), rng.integers(3 * n_fft, 8 * n_fft),
                            4 * rng.integers(3 * n_fft // 4, 2 * n_fft)]))
        tile = int(rng.choice([64, rng.integers(1, 12)]))
        quant, smooth = bool(rng.integers(0, 4) > 0), bool(rng.integers(0, 2))
        snap, smear = float(rng.choice([1.0, 0.9, 0.5, 0.25])), float(rng.choice([0.0, 0.1, 0.3, 0.6]))
        lo, hi = [(110.0, 5000.0), (0.0, 0.0), (60.0, 12000.0), (300.0, 2000.0)][rng.integers(4)]
        key, scale = keys[rng.integers(len(keys))], scales[rng.integers(len(scales))]
        sr = int(rng.choice([48000, 44100]))
        kind, seed = int(rng.integers(3)), int(rng.integers(1000))
        x = synth.noise_clip(seed, n) if kind == 0 else synth.bass_clip(seed, n, sr) if kind == 1 else synth.loud_clip(seed, n, sr)
        y, tap = run_emu(emu_spec, x, sr, n_fft, shape, tile, quant, smooth, snap, smear, prec=prec, lo=lo, hi=hi, key=key, scale=scale)
        ref = oracle_pass(x, sr, n_fft, quant, smooth, snap, smear, lo=lo, hi=hi, key=key, scale=scale)
        what = (n_fft, shape, prec, n, tile, quant, smooth, snap, smear, lo, hi, key, scale, sr, kind, seed)
        assert float(np.max(np.abs(y - ref))) <= (3e-5 if prec == "f32" else 2e-6), what
        assert np.array_equal(y, tap), what


def test_emu_formant_float64_on_reference_scenario(emu_spec):
    """Both quantised passes of the reference's own formant test (qd_cases.REF_SCENARIOS["formant_six"], fixture from the
    live reference) through the float64 kernel source: each pass on the REFERENCE's input of that pass reproduces the
    reference's output to one float32 rounding, while the float32 kernels are 1e-3 off -- the case is ill-conditioned
    (tests/test_oracle_ref_scenarios.py), which is why precision="auto" renders the formant shift in float64."""
    sc = np.load(os.path.join(HERE, "golden", "scenarios.npz"))
    ratio = 2.0 ** (6.0 / 12.0)
    for src, dst in (("x", "pre_quant"), ("post_dist", "y")):
        x, ref = sc[f"formant_six/{src}"], sc[f"formant_six/{dst}"]
        _, tap = run_emu(emu_spec, x, 44100, 2048, 8, 64, True, True, 1.0, 0.1, prec="f64", formant_ratio=ratio)
        assert float(np.max(np.abs(tap - ref))) <= 1.5e-8, (src, dst)
    _, tap32 = run_emu(emu_spec, sc["formant_six/x"], 44100, 2048, 8, 64, True, True, 1.0, 0.1, formant_ratio=ratio)
    assert float(np.max(np.abs(tap32 - sc["formant_six/pre_quant"]))) > 1e-4


@pytest.fixture(scope="module")
def emu_yin():
    exe = os.path.join(tempfile.gettempdir(), "qd_emu_yin")
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_yin.cpp", "cuda_emu.h")] + [os.path.join(CSRC, f) for f in ("qd_yin.cuh", "qd_common.cuh")]
    return _build_emu(exe, srcs, os.path.join(EMU_DIR, "emu_yin.cpp"))


@pytest.mark.parametrize("n,frame,hop,max_tau,cols", [(1500, 512, 64, 100, 88), (1500, 512, 64, 100, 3), (1000, 1024, 128, 61, 88),
                                                      (2100, 512, 128, 201, 7), (700, 512, 64, 255, 20), (3000, 512, 64, 255, 2),
                                                      (900, 4096, 512, 2047, 42)])   # the last: a tile that needs two prefix rounds
def test_emu_yin_difference_function(emu_yin, n, frame, hop, max_tau, cols):
    """The sliding-sum walk (walker and helper warps, three sample groups x columns of 16 lags, captures inside the iteration,
    `cols` columns per CTA)
    against the definition d_f(tau) = sum_j (x[s+j] - x[s+j+tau])^2 over the zero-extended clip (dsp/autotune.py:141-151)."""
    t = np.arange(n)
    x = (0.5 * np.sin(2 * np.pi * t / 37.3) + 0.1 * np.random.default_rng(n).standard_normal(n)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        x.tofile(os.path.join(d, "x"))
        subprocess.check_call([emu_yin, str(n), str(frame), str(hop), str(max_tau), str(cols), os.path.join(d, "x"), os.path.join(d, "d")])
        got = np.fromfile(os.path.join(d, "d")).reshape(-1, (max_tau + 2) & ~1)[:, 1:max_tau + 1]
    xz = np.concatenate([x.astype(np.float64), np.zeros(frame + max_tau + hop)])
    ref = np.zeros_like(got)
    for f in range(got.shape[0]):
        for tau in range(1, max_tau + 1):
            dd = xz[f * hop:f * hop + frame - tau] - xz[f * hop + tau:f * hop + frame]
            ref[f, tau - 1] = np.dot(dd, dd)
    assert (got >= 0).all(), "every (frame, lag) entry is written"
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
