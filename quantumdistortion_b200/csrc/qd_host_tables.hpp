// qd_host_tables.hpp -- host-side (plain C++) builders for the device tables of qd_spec.cuh.
// Pure arithmetic in double, rounded once to float; no CUDA dependency (also used by the
// g++ emulation tests).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qd_b200.h"

namespace qd_host {

struct F2 { float x, y; };
struct D2 { double x, y; };
template <class T> struct Pair;
template <> struct Pair<float>  { using type = F2; };
template <> struct Pair<double> { using type = D2; };

struct FftRadices { int r1, r2, r3; };
// must mirror qd::FftCfg<T, NC>
inline bool fft_radices(int nc, FftRadices *o, bool is_double = false) {
    switch (nc) {
        case 256:  *o = {8, 8, 4};   return true;
        case 512:  *o = {8, 8, 8};   return true;
        case 1024: if (is_double) *o = {16, 8, 8}; else *o = {32, 32, 1}; return true;
        case 2048: *o = {16, 16, 8}; return true;
        case 4096: *o = {16, 16, 16}; return true;
    }
    return false;
}

template <class T>
struct SpecTablesT {
    using P = typename Pair<T>::type;
    int nc = 0, hop = 0;
    std::vector<P> wtab, tw1, tw2, wsplit;
    std::vector<T> invw;  // [16][hop]
};
using SpecTables = SpecTablesT<float>;

// periodic Hann, hop = n_fft/4 (dsp/stft_utils.py:51-56, 137-141)
inline double hann_periodic(int n, int n_fft) { return 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)n_fft); }

template <class T>
inline bool build_spec_tables(int n_fft, SpecTablesT<T> *t) {
    using F2 = typename Pair<T>::type;
    constexpr bool is_double = sizeof(T) == 8;
    const int nc = n_fft / 2;
    FftRadices r;
    if (!fft_radices(nc, &r, is_double)) return false;
    t->nc = nc;
    t->hop = n_fft / 4;
    // Hann window by angle addition: the first-pass butterfly a0 (0 <= a0 < S = nc/R1) touches the sample pairs
    // 2(a0 + qS), 2(a0 + qS) + 1, q = 0..R1-1, whose window is 0.5 - 0.5 cos(A + 2 pi q / R1), A = 2 pi (2 a0 [+1]) / n_fft.
    // Entry 2 a0 = -0.5 cos A (even, odd sample), entry 2 a0 + 1 = 0.5 sin A; cos/sin(2 pi q / R1) are compile-time
    // constants of the kernel, so the window costs two FFMA2 instead of a shared-memory read per sample pair.
    {
        const int S = nc / r.r1;
        t->wtab.resize((size_t)2 * S);
        for (int a0 = 0; a0 < S; ++a0) {
            const double ae = 2.0 * M_PI * (double)(2 * a0) / (double)n_fft, ao = 2.0 * M_PI * (double)(2 * a0 + 1) / (double)n_fft;
            t->wtab[2 * a0] = F2{(T)(-0.5 * std::cos(ae)), (T)(-0.5 * std::cos(ao))};
            t->wtab[2 * a0 + 1] = F2{(T)(0.5 * std::sin(ae)), (T)(0.5 * std::sin(ao))};
        }
    }
    // access-ordered twiddles: entry [(i*R + k)*32 + lane] = exp(-2 pi i j k / M),
    // butterfly u = lane + 32 i of a pass with sub-size M, radix R, stride S = M/R, j = u % S
    auto fill = [&](std::vector<F2> &tw, int M, int R) {
        const int S = M / R, nb = nc / R / 32;
        tw.assign((size_t)nb * R * 32, F2{(T)1, (T)0});
        for (int i = 0; i < nb; ++i)
            for (int k = 0; k < R; ++k)
                for (int lane = 0; lane < 32; ++lane) {
                    const int j = (lane + 32 * i) % S;
                    const double ang = -2.0 * M_PI * (double)(((long long)j * k) % M) / (double)M;
                    tw[((size_t)i * R + k) * 32 + lane] = F2{(T)std::cos(ang), (T)std::sin(ang)};
                }
    };
    fill(t->tw1, nc, r.r1);
    if (r.r3 > 1) fill(t->tw2, nc / r.r1, r.r2);
    else t->tw2.assign(1, F2{(T)1, (T)0});
    t->wsplit.resize(nc / 2 + 1);
    for (int k = 0; k <= nc / 2; ++k) {
        const double ang = -2.0 * M_PI * (double)k / (double)n_fft;
        // V_k = -i/2 * exp(-2 pi i k / n_fft): the factor of the odd part in the real split (and, conjugated
        // and doubled, in the Hermitian merge) with the 1/(2i) already folded in
        t->wsplit[k] = F2{(T)(0.5 * std::sin(ang)), (T)(-0.5 * std::cos(ang))};
    }
    t->wsplit[0] = F2{(T)0, (T)-0.5};         // exact quarter turns
    t->wsplit[nc / 2] = F2{(T)-0.5, (T)0};
    // 1 / max(sum of w^2 over the frames covering a hop-block, 1e-10)  (dsp/stft_utils.py:174-186, 214)
    // frames are added in ascending t = descending slice index.
    const int hop = t->hop;
    t->invw.assign((size_t)16 * hop, (T)0);
    for (int a = 0; a < 4; ++a)
        for (int b = a; b < 4; ++b)
            for (int c = 0; c < hop; ++c) {
                double s = 0.0;
                for (int sl = b; sl >= a; --sl) {
                    const double w = hann_periodic(sl * hop + c, n_fft);
                    s += w * w;
                }
                // x 1/n_fft: the inverse FFT in the kernel is unnormalised
                t->invw[(size_t)(a * 4 + b) * hop + c] = (T)(1.0 / std::max(s, 1e-10) / (double)n_fft);
            }
    return true;
}

// host mirror of qd::spos<NC>: position of spectrum bin k inside a warp buffer
inline int host_spos(int nc, int k, bool is_double = false) {
    FftRadices r;
    if (!fft_radices(nc, &r, is_double)) return -1;
    if (k >= nc) return 32;  // QD_NYQ_SLOT
    const int k1 = k % r.r1, k2 = (k / r.r1) % r.r2, k3 = k / (r.r1 * r.r2);
    const int a = k1 * (nc / r.r1) + k2 * (nc / (r.r1 * r.r2)) + k3;
    return a + (a >> 5);
}

struct QuantTablesH {
    int n_bins = 0, n_slots = 0, rows = 0, row_limit = 0;
    std::vector<uint32_t> src_tab;  // tail<<31 | off<<26 | slot<<13 | source bin, grouped by slot
    std::vector<uint16_t> slot_begin, src_bin;
    std::vector<int32_t> slot_bin;
    std::vector<uint32_t> row_active;
    std::vector<uint16_t> slot_of_bin;  // [32*rows + 4], entry d+2 = slot with target bin d, else n_slots
    std::vector<float> slot_invk, slot_base;  // [n_slots + 1]
    float tap[5] = {0, 0, 0, 0, 0};
    float keep_active = 1.0f;
};

// Gather form of dsp/quantizer.py:424-515 derived from the reference's own integer tables.
inline bool build_quant_tables(const qd_tables &in, QuantTablesH *q, std::string *err, bool is_double = false) {
    const int n = in.n_bins;
    if (n < 3 || n > 8193 || !in.target_bins || !in.active_mask) { if (err) *err = "bad quantizer tables"; return false; }
    const int radius = in.smear_radius;
    if (radius < 0 || radius > 2) { if (err) *err = "smear_radius must be 0..2"; return false; }
    if (radius > 0 && !in.smear_w) { if (err) *err = "smear_w missing"; return false; }
    q->n_bins = n;
    q->rows = (n + 31) / 32;
    // sources: active bins with an in-range target (dsp/quantizer.py:426-431); E > 0 is a per-frame test
    std::vector<std::vector<uint16_t>> by_target(n);
    q->row_active.assign(q->rows, 0u);
    for (int i = 0; i < n; ++i) {
        const int t = in.target_bins[i];
        if (in.active_mask[i] && t >= 0 && t < n) {
            by_target[t].push_back((uint16_t)i);
            q->row_active[i >> 5] |= 1u << (i & 31);
        }
    }
    std::vector<int> slot_of(n, -1);
    q->slot_begin.clear(); q->src_bin.clear(); q->slot_bin.clear();
    for (int t = 0; t < n; ++t)
        if (!by_target[t].empty()) {
            slot_of[t] = (int)q->slot_bin.size();
            q->slot_bin.push_back(t);
            q->slot_begin.push_back((uint16_t)q->src_bin.size());
            for (uint16_t s : by_target[t]) q->src_bin.push_back(s);
        }
    q->slot_begin.push_back((uint16_t)q->src_bin.size());
    q->n_slots = (int)q->slot_bin.size();
    // entry i of the gather table; `off` = same-slot sources before it inside its group of 32,
    // `tail` = last source of its slot inside the group (see quantize_frame Q1)
    q->src_tab.clear();
    {
        std::vector<int> slot_of_src;
        for (int s = 0; s < q->n_slots; ++s)
            for (int i = q->slot_begin[s]; i < q->slot_begin[s + 1]; ++i) slot_of_src.push_back(s);
        const int ns = (int)slot_of_src.size();
        for (int i = 0; i < ns; ++i) {
            int off = 0;
            for (int j = i - 1; j >= (i / 32) * 32 && slot_of_src[j] == slot_of_src[i]; --j) ++off;
            const bool tail = (i % 32 == 31) || (i == ns - 1) || (slot_of_src[i + 1] != slot_of_src[i]);
            const uint32_t pos = (uint32_t)q->src_bin[i];
            (void)is_double;
            q->src_tab.push_back(((uint32_t)tail << 31) | ((uint32_t)off << 26) | ((uint32_t)slot_of_src[i] << 13) | pos);
        }
    }
    if (q->src_bin.empty()) q->src_bin.push_back(0);
    const double snap = in.snap, smear = in.smear;
    const bool do_smear = smear > 0.0 && radius > 0;  // dsp/quantizer.py:458
    q->keep_active = (float)(1.0 - snap);
    // bin d receives from the target at bin t = d+e-2 the tap at offset o = d - t = 2 - e  (kernel index o + radius)
    for (int e = 0; e < 5; ++e) {
        const int o = 2 - e;
        q->tap[e] = (do_smear && o >= -radius && o <= radius) ? (float)(snap * smear * in.smear_w[o + radius]) : 0.0f;
    }
    q->slot_of_bin.assign((size_t)32 * q->rows + 4, (uint16_t)q->n_slots);
    q->slot_invk.assign((size_t)q->n_slots + 1, 1.0f);
    q->slot_base.assign((size_t)q->n_slots + 1, 0.0f);
    q->row_limit = 0;
    for (int r = 0; r < q->rows; ++r)
        if (q->row_active[r]) q->row_limit = r + 1;
    for (int s = 0; s < q->n_slots; ++s) {
        const int t = q->slot_bin[s];
        q->slot_of_bin[(size_t)t + 2] = (uint16_t)s;
        // local kernel re-normalised over the part of [t-r, t+r] inside [0, n)  (dsp/quantizer.py:311-330)
        double ksum = 1.0;
        if (do_smear) {
            const int a = std::max(0, t - radius), b = std::min(n, t + radius + 1);
            const int k0 = std::max(0, radius - t);
            ksum = 0.0;
            for (int qq = k0; qq < k0 + (b - a); ++qq) ksum += in.smear_w[qq];
            if (!(ksum > 0.0)) ksum = 1.0;
        }
        q->slot_invk[s] = (float)(1.0 / ksum);
        q->slot_base[s] = (float)(snap * (1.0 - smear) * ksum);  // base energy lands on the target itself (:437, :446)
        const int last_row = std::min(q->rows - 1, (std::min(n - 1, t + (do_smear ? radius : 0))) >> 5);
        if (last_row + 1 > q->row_limit) q->row_limit = last_row + 1;
    }
    return true;
}

// Gather lists of the team kernel (qd_spec_team.cuh): the target slots, in ascending order, are cut into `cw` contiguous
// ranges of about equal source counts; warp w's list holds the sources of its slots in the order of src_tab, `off` and
// `tail` counted in groups of 32 from the start of ITS list, padded with null entries (0: no tail, nothing stored) to a
// multiple of 32.  begin[w] .. begin[w + 1] delimits warp w's list inside `tab`.
inline void build_team_gather(const QuantTablesH &q, int cw, std::vector<uint32_t> *tab, int *begin /* [cw + 1] */) {
    tab->clear();
    const int ns_total = q.n_slots > 0 ? (int)q.slot_begin[q.n_slots] : 0;
    int slot = 0;
    for (int w = 0; w < cw; ++w) {
        begin[w] = (int)tab->size();
        // slots whose first source lies below the w+1-th share of the sources
        const long long limit = ((long long)ns_total * (w + 1) + cw - 1) / cw;
        std::vector<int> slot_of_src;
        std::vector<uint16_t> bins;
        while (slot < q.n_slots && (w == cw - 1 || (long long)q.slot_begin[slot] < limit)) {
            for (int i = q.slot_begin[slot]; i < q.slot_begin[slot + 1]; ++i) {
                slot_of_src.push_back(slot);
                bins.push_back(q.src_bin[i]);
            }
            ++slot;
        }
        const int ns = (int)slot_of_src.size();
        for (int i = 0; i < ns; ++i) {
            int off = 0;
            for (int j = i - 1; j >= (i / 32) * 32 && slot_of_src[j] == slot_of_src[i]; --j) ++off;
            const bool tail = (i % 32 == 31) || (i == ns - 1) || (slot_of_src[i + 1] != slot_of_src[i]);
            tab->push_back(((uint32_t)tail << 31) | ((uint32_t)off << 26) | ((uint32_t)slot_of_src[i] << 13) | (uint32_t)bins[i]);
        }
        while (tab->size() % 32) tab->push_back(0u);
    }
    begin[cw] = (int)tab->size();
    if (tab->empty()) tab->push_back(0u);
}

}  // namespace qd_host
