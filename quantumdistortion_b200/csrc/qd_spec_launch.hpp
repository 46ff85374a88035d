// qd_spec_launch.hpp -- launcher of the spectral pass.  qd_api.cu sees only the DECLARATION; the definition
// (qd_spec_launch.inl) is explicitly instantiated in qd_k_spec_*.cu, one translation unit per kernel family, so that
// the families compile in parallel and a change to the API layer does not recompile a single kernel.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "qd_spec.cuh"

namespace qd_launch {

// T float | double; NC = n_fft / 2; NW warps per clip group; TS tables in shared memory; FX spectral-FX variant;
// NG clip groups per CTA; SA samples staged in the formant scratch (qd_spec.cuh, SpecSmem); EF float32-only epilogue
// (the caller checked that the wavefold is exact in float32, see epilogue_apply)
template <class T, int NC, int NW, bool TS, bool FX, int NG = 1, bool SA = false, bool EF = false>
int launch_spec_t(const qd::SpecArgsT<T> &a, int tiles, int64_t batch, cudaStream_t st);

}  // namespace qd_launch
