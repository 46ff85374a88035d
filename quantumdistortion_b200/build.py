"""Build libqd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m quantumdistortion_b200.build [--force] [-v]

The library is several translation units -- the API layer and one unit per family of spectral-pass kernels --
compiled in parallel and linked into one shared object; an object is rebuilt only when one of its own dependencies
changed.  The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libqd_b200.so")
HEADER = os.path.join("..", "..", "include", "qd_b200.h")
_KERNEL_DEPS = ["qd_spec.cuh", "qd_common.cuh", "qd_spec_launch.hpp", "qd_spec_launch.inl", "qd_err.hpp", HEADER]
# translation unit -> what it includes
UNITS = {
    "qd_api.cu": ["qd_spec.cuh", "qd_spec_launch.hpp", "qd_spec_team.cuh", "qd_spec_team_launch.hpp", "qd_peaks.cuh", "qd_autotune.cuh", "qd_yin.cuh", "qd_autotune_api.inc",
                  "qd_host_pipe.inc", "qd_time.cuh", "qd_common.cuh", "qd_host_tables.hpp", "qd_host_time.hpp",
                  "qd_err.hpp", HEADER],
    "qd_k_spec_f32.cu": _KERNEL_DEPS,
    "qd_k_spec_fx32.cu": _KERNEL_DEPS,
    "qd_k_spec_fx32b.cu": _KERNEL_DEPS,
    "qd_k_spec_f64.cu": _KERNEL_DEPS,
    "qd_k_spec_fx64.cu": _KERNEL_DEPS,
    "qd_k_spec_team.cu": ["qd_spec_team.cuh", "qd_spec_team_launch.hpp", "qd_spec.cuh", "qd_common.cuh", "qd_err.hpp", HEADER],
}

# no --split-compile: it changes the generated code from build to build (the headline kernel came out 21 % slower
# in one of them, measured); parallelism comes from the translation units instead
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libqd_b200.so cannot be built (there is no CPU fallback)")


def _obj(unit: str) -> str:
    return os.path.join(OBJ, os.path.splitext(unit)[0] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps)


def needs_build() -> bool:
    return not os.path.exists(LIB) or any(_stale(LIB, [u] + deps) for u, deps in UNITS.items())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    todo = [u for u, deps in UNITS.items() if force or _stale(_obj(u), [u] + deps)]

    def compile_unit(unit: str):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(unit), os.path.join(CSRC, unit)]
        return unit, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        for unit, res in pool.map(compile_unit, todo):
            if verbose:
                sys.stderr.write(res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {unit}:\n" + res.stdout + res.stderr)
    link = [nvcc, "-shared", "-cudart", "static", "-o", LIB] + [_obj(u) for u in UNITS]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
