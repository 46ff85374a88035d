// qd_spec_team_launch.hpp -- launcher of the team kernel (qd_spec_team.cuh: CW warps per frame on swizzled buffers, the
// three-pass FFT plans, plain variant).  qd_api.cu sees only this declaration; qd_k_spec_team.cu holds the definition and the instantiations.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "qd_spec_team.cuh"

namespace qd_launch {

// frames per batch (NF) and warps per frame (CW) of the team kernel for this precision and n_fft / 2, or {0, 0} when the
// configuration keeps the one-warp-per-frame kernel.  Sized so that the CTA fills one SM: float32 n_fft 4096 has 7
// frame buffers of 17 KB (28 warps at 72 registers), float64 n_fft 4096 four of 34 KB, float64 n_fft 8192 two of 68 KB.
struct TeamShape { int nf, cw; };
// The short three-pass plans (n_fft 512 / 1024, float32) run the same kernel with one warp per frame (CW = 1) for its
// swizzled buffer layout: the padded layout of spec_pass_kernel spends 45 % / 33 % of its shared-memory wavefronts on
// bank-conflict replays there and is bound by them (LSU data pipe 84 %).
template <class T> inline TeamShape team_shape(int nc) {
    if (sizeof(T) == 4) return nc == 2048 ? TeamShape{7, 4} : nc == 4096 ? TeamShape{4, 4} : (nc == 256 || nc == 512) ? TeamShape{8, 1} : TeamShape{0, 0};
    // float64 at the default n_fft 2048 (16 x 8 x 8; what precision="auto" picks for a wide-open band mask): 8 frames x 2 warps
    return nc == 1024 ? TeamShape{8, 2} : nc == 2048 ? TeamShape{4, 4} : nc == 4096 ? TeamShape{2, 8} : TeamShape{0, 0};
}

// dynamic shared memory of that kernel (0: none)
template <class T> inline size_t team_smem_bytes(int nc, int n_slots) {
    if (sizeof(T) == 4)
        return nc == 2048 ? qd::SpecSmem<float, 2048, 7>::bytes(n_slots) : nc == 4096 ? qd::SpecSmem<float, 4096, 4>::bytes(n_slots)
             : nc == 512 ? qd::SpecSmem<float, 512, 8>::bytes(n_slots)
             : nc == 256 ? qd::SpecSmem<float, 256, 8>::bytes(n_slots) : 0;
    return nc == 1024 ? qd::SpecSmem<double, 1024, 8>::bytes(n_slots) : nc == 2048 ? qd::SpecSmem<double, 2048, 4>::bytes(n_slots)
         : nc == 4096 ? qd::SpecSmem<double, 4096, 2>::bytes(n_slots) : 0;
}

template <class T>
int launch_spec_team(int nc, const qd::SpecArgsT<T> &a, const qd::TeamGather &tg, int tiles, int64_t batch, cudaStream_t st);

}  // namespace qd_launch
