// qd_host_time.hpp -- host helpers for qd_time.cuh (plain C++; shared with the emulation tests).
#pragma once
#include <cmath>
#include <cstddef>

namespace qd_host {

// double-buffered |x| (padded) and group maxima plus the per-warp scan totals of limiter_mix_kernel
inline size_t limiter_smem_bytes(int lookahead) {
    const int span = 256 * 8 + lookahead + 8;
    const size_t abs_stride = ((size_t)(span + (span >> 5)) + 8 + 3) & ~(size_t)3;
    const size_t gmax_stride = ((size_t)(span / 8) + 2 + 3) & ~(size_t)3;
    return 2 * (abs_stride + gmax_stride) * sizeof(float) + 16 + 2 * (8 + 2) * sizeof(double);
}

// sos_low / sos_high: [2][6] rows (b0 b1 b2 a0 a1 a2), a0 == 1 (scipy layout).  Besides the coefficients this picks the
// tiling of crossover_kernel: `halo` = warm-up samples after which a zero-started state equals the true one to below
// 1e-15 (the zero-input response of a double pole pair decays like n r^n, r = largest pole radius: 44 / (1 - r)
// samples), `tile` = output samples per thread (>= 4096 and >= 2 halo, so the warm-up costs at most half the work).
// largest pole radius of `sections` second-order sections (rows b0 b1 b2 a0 a1 a2)
inline double sos_pole_radius(const double *sos, int sections) {
    double r = 0.0;
    for (int s = 0; s < sections; ++s) {
        const double a1 = sos[6 * s + 4], a2 = sos[6 * s + 5], disc = a1 * a1 - 4.0 * a2;
        const double rad = disc < 0.0 ? std::sqrt(a2 > 0.0 ? a2 : 0.0)
                                      : std::fmax(std::fabs((-a1 + std::sqrt(disc)) * 0.5), std::fabs((-a1 - std::sqrt(disc)) * 0.5));
        r = std::fmax(r, rad);
    }
    return r;
}

// warm-up length after which a zero-started state equals the true one to below 1e-15 (44 / (1 - r) samples, a multiple of
// 32), and the tile that keeps the warm-up at no more than half the work
inline void segment_tiling(double pole_radius, int *tile, int *halo) {
    double h = pole_radius < 1.0 ? 44.0 / (1.0 - pole_radius) : 1e9;
    if (h > (double)(1 << 22)) h = (double)(1 << 22);     // a (nearly) unstable design: bounded work, bounded accuracy
    *halo = (((int)std::ceil(h)) + 31) & ~31;
    int t = 4096;
    while (t < 2 * *halo) t *= 2;
    *tile = t;
}

template <class Args>
inline void fill_crossover(Args &a, const double *sos_low, const double *sos_high) {
    for (int s = 0; s < 4; ++s) {
        const double *src = (s < 2 ? sos_low : sos_high) + 6 * (s & 1);
        for (int i = 0; i < 6; ++i) a.co[s][i] = src[i];
    }
    segment_tiling(std::fmax(sos_pole_radius(sos_low, 2), sos_pole_radius(sos_high, 2)), &a.tile, &a.halo);
}

}  // namespace qd_host
