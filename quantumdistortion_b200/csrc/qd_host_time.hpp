// qd_host_time.hpp -- host helpers for qd_time.cuh (plain C++; shared with the emulation tests).
#pragma once
#include <cstddef>

namespace qd_host {

// double-buffered |x| (padded) and group maxima plus the per-warp scan totals of limiter_mix_kernel
inline size_t limiter_smem_bytes(int lookahead) {
    const int span = 256 * 8 + lookahead + 8;
    const size_t abs_stride = ((size_t)(span + (span >> 5)) + 8 + 3) & ~(size_t)3;
    const size_t gmax_stride = ((size_t)(span / 8) + 2 + 3) & ~(size_t)3;
    return 2 * (abs_stride + gmax_stride) * sizeof(float) + 16 + 2 * (8 + 2) * sizeof(double);
}

// A = [[-a1, 1], [-a2, 0]] is the DF2T state matrix of a section; apow[s][l] = A_s^(8 * 2^l)
// sos_low / sos_high: [2][6] rows (b0 b1 b2 a0 a1 a2), a0 == 1 (scipy layout)
template <class Args>
inline void fill_crossover(Args &a, const double *sos_low, const double *sos_high) {
    for (int s = 0; s < 4; ++s) {
        const double *src = (s < 2 ? sos_low : sos_high) + 6 * (s & 1);
        for (int i = 0; i < 6; ++i) a.co[s][i] = src[i];
        auto mul = [](const double *p, const double *q, double *o) {
            double r[4] = {p[0] * q[0] + p[1] * q[2], p[0] * q[1] + p[1] * q[3],
                           p[2] * q[0] + p[3] * q[2], p[2] * q[1] + p[3] * q[3]};
            for (int i = 0; i < 4; ++i) o[i] = r[i];
        };
        double p[4] = {-src[4], 1.0, -src[5], 0.0};
        for (int k = 0; k < 3; ++k) mul(p, p, p);  // A^8
        for (int l = 0; l < 6; ++l) {
            a.apow[s][l].a = p[0]; a.apow[s][l].b = p[1]; a.apow[s][l].c = p[2]; a.apow[s][l].d = p[3];
            mul(p, p, p);
        }
    }
}

}  // namespace qd_host
