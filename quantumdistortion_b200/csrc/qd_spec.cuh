// qd_spec.cuh -- one spectral pass of the STFT path as a single kernel:
//
//   frame -> Hann window -> real FFT -> [spectral FX] -> scale-snap quantizer (gather form,
//   5-tap smear, 3-tap smoothing) -> Hermitian inverse FFT -> Hann window -> overlap-add ->
//   1/sum(w^2) -> float32 -> [wavefold | soft-tube]
//
// Replaces, per pass, the reference's
//   stft_mono                               dsp/stft_utils.py:11-97
//   _apply_spectral_quantization_to_stft    dsp/pipeline.py:228-342
//   quantize_spectrum / _apply_smear_numba  dsp/quantizer.py:343-529, 253-340
//   istft_mono                              dsp/stft_utils.py:100-234
//   apply_distortion                        dsp/distortion.py:18-114 (fused epilogue)
//
// Execution model (B200): one WARP owns one frame.  The n_fft-point real FFT is an
// (n_fft/2)-point complex FFT held in the warp's private shared-memory buffer; each lane
// runs whole radix-R DFTs (R up to 32) in registers, so n_fft = 2048 needs two passes
// (32 x 32) and a single shared-memory exchange per direction, with only __syncwarp()
// between passes.  The NW warps of a CTA take NW consecutive frames; a CTA-wide step then
// overlap-adds them in frame order (deterministic, no atomics) and streams finished hops
// to HBM.  A CTA walks a tile of consecutive hops of one clip, carrying the 3-hop OLA
// tail in shared memory.  Nothing is a dense contraction, so tensor cores are not used.
#pragma once
#include "qd_common.cuh"

namespace qd {

// ---------------------------------------------------------------- FFT configuration
template <int NC> struct FftCfg;
template <> struct FftCfg<256>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 4;  };
template <> struct FftCfg<512>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 8;  };
template <> struct FftCfg<1024> { static constexpr int R1 = 32, R2 = 32, R3 = 1;  };
template <> struct FftCfg<2048> { static constexpr int R1 = 16, R2 = 16, R3 = 8;  };
template <> struct FftCfg<4096> { static constexpr int R1 = 16, R2 = 16, R3 = 16; };

// one pad slot per 32 complex values keeps the stride-32 accesses of the 32x32 plan
// conflict-free (stride 33); slot 32 is never produced by pidx() and holds the Nyquist bin.
QD_DEV int pidx(int a) { return a + (a >> 5); }
constexpr int QD_NYQ_SLOT = 32;
template <int NC> constexpr int buf_slots() { return NC + NC / 32; }

// position of spectrum bin k (0..NC) inside the warp buffer after the in-place DIF passes
template <int NC>
QD_DEV int spos(int k) {
    using C = FftCfg<NC>;
    if (k >= NC) return QD_NYQ_SLOT;
    const int k1 = k & (C::R1 - 1);
    const int k2 = (k / C::R1) & (C::R2 - 1);
    const int k3 = k / (C::R1 * C::R2);
    return pidx(k1 * (NC / C::R1) + k2 * (NC / (C::R1 * C::R2)) + k3);
}

// ---------------------------------------------------------------- device-side tables
struct AffEntry {          // one destination bin that can receive moved energy
    int16_t slot[5];       // slot of the target at bin d-2..d+2 (n_slots = "none", reads 0)
    int16_t pad;
    float   coef[5];       // snap*smear*k_t(d-t) (+ snap*(1-smear) for the own target)
};

struct QuantDev {
    int n_slots;                    // distinct target bins
    int n_aff;
    const uint16_t *slot_begin;     // [n_slots+1] CSR offsets into src_bin
    const uint16_t *src_bin;        // source bins, grouped by target, ascending
    const uint32_t *row_active;     // [rows] bit l: bin 32*row+l gives its energy away
    const uint32_t *row_aff;        // [rows] bit l: bin 32*row+l is in the affected list
    const uint16_t *row_aff_base;   // [rows] affected bins before this row
    const AffEntry *aff;            // [n_aff]
    float keep_active;              // 1 - snap
    int smoothing;                  // dsp/quantizer.py:523
};

struct SpecArgs {
    const float *x;        // [batch, n] input clips
    float *y;              // [batch, n] output (after the epilogue)
    float *tap;            // optional [batch, n]: iSTFT output before the epilogue
    int n;                 // samples per clip
    int n_frames;          // T = 1 + n / hop
    int tile_blocks;       // output hops per CTA tile
    int quant;             // run the quantizer (else pure STFT -> iSTFT)
    int epilogue;          // 0 none, 1 wavefold, 2 tube
    float fold, bias, tube_gain, tube_norm;
    const float2 *wtab;    // [NC] Hann window as pairs (w[2n], w[2n+1])
    const float2 *tw1;     // access-ordered twiddles of pass 1 / pass 2 (see host builder)
    const float2 *tw2;
    const float2 *wsplit;  // [NC/2+1] exp(-2 pi i k / n_fft)
    const float *invw;     // [16][hop]: 1/max(sum_{sl=a..b} w^2[sl*hop+c], 1e-10) at [(a*4+b)*hop + c]
    QuantDev q;
};

// ---------------------------------------------------------------- FFT passes (warp level)
template <int NC, int M, int R, bool TW>
QD_DEV void fwd_pass(float2 *buf, const float2 *tw, int lane) {
    constexpr int S = M / R;
    constexpr int NB = NC / R / 32;
    constexpr int LG = qd_log2(R);
#pragma unroll 1
    for (int i = 0; i < NB; ++i) {
        const int u = lane + 32 * i;
        const int a0 = (u / S) * M + (u % S);
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = buf[pidx(a0 + q * S)];
        dft_reg<R, -1>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = qd_bitrev(r, LG);
            float2 t = v[r];
            if (TW && k > 0) t = cmul(t, tw[(i * R + k) * 32 + lane]);
            buf[pidx(a0 + k * S)] = t;
        }
    }
    __syncwarp();
}

template <int NC, int M, int R, bool TW>
QD_DEV void inv_pass(float2 *buf, const float2 *tw, int lane) {
    constexpr int S = M / R;
    constexpr int NB = NC / R / 32;
    constexpr int LG = qd_log2(R);
#pragma unroll 1
    for (int i = 0; i < NB; ++i) {
        const int u = lane + 32 * i;
        const int a0 = (u / S) * M + (u % S);
        float2 v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float2 t = buf[pidx(a0 + k * S)];
            if (TW && k > 0) t = cmulc(t, tw[(i * R + k) * 32 + lane]);
            v[k] = t;
        }
        dft_reg<R, +1>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) buf[pidx(a0 + qd_bitrev(r, LG) * S)] = v[r];
    }
    __syncwarp();
}

// first forward pass: reads the frame from the staging buffer, applies the analysis window
template <int NC, int R>
QD_DEV void fwd_first(float2 *buf, const float2 *frame, const float2 *wtab, const float2 *tw, int lane) {
    constexpr int S = NC / R;
    constexpr int NB = NC / R / 32;
    constexpr int LG = qd_log2(R);
#pragma unroll 1
    for (int i = 0; i < NB; ++i) {
        const int a0 = lane + 32 * i;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const float2 s = frame[a0 + q * S];
            const float2 w = wtab[a0 + q * S];
            v[q] = make_float2(s.x * w.x, s.y * w.y);
        }
        dft_reg<R, -1>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = qd_bitrev(r, LG);
            float2 t = v[r];
            if (k > 0) t = cmul(t, tw[(i * R + k) * 32 + lane]);
            buf[pidx(a0 + k * S)] = t;
        }
    }
    __syncwarp();
}

// last inverse pass: synthesis window and 1/n_fft, leaves the time-domain frame in buf
template <int NC, int R>
QD_DEV void inv_last(float2 *buf, const float2 *wtab, const float2 *tw, int lane) {
    constexpr int S = NC / R;
    constexpr int NB = NC / R / 32;
    constexpr int LG = qd_log2(R);
    const float scale = 1.0f / (float)(2 * NC);
#pragma unroll 1
    for (int i = 0; i < NB; ++i) {
        const int a0 = lane + 32 * i;
        float2 v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float2 t = buf[pidx(a0 + k * S)];
            if (k > 0) t = cmulc(t, tw[(i * R + k) * 32 + lane]);
            v[k] = t;
        }
        dft_reg<R, +1>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = a0 + qd_bitrev(r, LG) * S;
            const float2 w = wtab[n];
            buf[pidx(n)] = make_float2(v[r].x * (w.x * scale), v[r].y * (w.y * scale));
        }
    }
    __syncwarp();
}

template <int NC>
QD_DEV void fft_forward(float2 *buf, const float2 *frame, const SpecArgs &a, const float2 *wtab,
                        const float2 *tw1, const float2 *tw2, int lane) {
    using C = FftCfg<NC>;
    fwd_first<NC, C::R1>(buf, frame, wtab, tw1, lane);
    if constexpr (C::R3 > 1) {
        fwd_pass<NC, NC / C::R1, C::R2, true>(buf, tw2, lane);
        fwd_pass<NC, C::R3, C::R3, false>(buf, nullptr, lane);
    } else {
        fwd_pass<NC, NC / C::R1, C::R2, false>(buf, nullptr, lane);
    }
    (void)a;
}

template <int NC>
QD_DEV void fft_inverse(float2 *buf, const float2 *wtab, const float2 *tw1, const float2 *tw2, int lane) {
    using C = FftCfg<NC>;
    if constexpr (C::R3 > 1) {
        inv_pass<NC, C::R3, C::R3, false>(buf, nullptr, lane);
        inv_pass<NC, NC / C::R1, C::R2, true>(buf, tw2, lane);
    } else {
        inv_pass<NC, NC / C::R1, C::R2, false>(buf, nullptr, lane);
    }
    inv_last<NC, C::R1>(buf, wtab, tw1, lane);
}

// ---------------------------------------------------------------- real <-> complex packing
// Z = FFT_NC(x[2n] + i x[2n+1])  ->  X[k], k = 0..NC   (in place, Nyquist in the pad slot)
//   E = (Z[k] + conj Z[NC-k]) / 2,  T = W_N^k (Z[k] - conj Z[NC-k]) / (2i)
//   X[k] = E + T,  X[NC-k] = conj(E - T)
template <int NC>
QD_DEV void real_split(float2 *buf, const float2 *wsplit, int lane) {
#pragma unroll 2
    for (int k = lane; k < NC / 2; k += 32) {
        if (k == 0) {
            const float2 z0 = buf[spos<NC>(0)];
            buf[spos<NC>(0)] = make_float2(z0.x + z0.y, 0.0f);
            buf[QD_NYQ_SLOT] = make_float2(z0.x - z0.y, 0.0f);
            const int pm = spos<NC>(NC / 2);
            buf[pm] = cconj(buf[pm]);
        } else {
            const int pa = spos<NC>(k), pb = spos<NC>(NC - k);
            const float2 za = buf[pa], zb = buf[pb];
            const float2 e = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
            const float2 o = make_float2(0.5f * (za.y + zb.y), -0.5f * (za.x - zb.x));  // (za - conj zb)/(2i)
            const float2 t = cmul(o, wsplit[k]);
            buf[pa] = cadd(e, t);
            buf[pb] = cconj(csub(e, t));
        }
    }
    __syncwarp();
}

// X'[k] (only Re of DC / Nyquist used, like pocketfft c2r) -> Z' with z = IFFT_NC(Z') * 1/(2 NC)
//   E2 = X'[k] + conj X'[NC-k],  T2 = X'[k] - conj X'[NC-k],  O2 = conj(W_N^k) T2
//   Z'[k] = E2 + i O2,  Z'[NC-k] = conj(E2 - i O2)
template <int NC>
QD_DEV void real_merge(float2 *buf, const float2 *wsplit, int lane) {
#pragma unroll 2
    for (int k = lane; k < NC / 2; k += 32) {
        if (k == 0) {
            const float a = buf[spos<NC>(0)].x, b = buf[QD_NYQ_SLOT].x;
            buf[spos<NC>(0)] = make_float2(a + b, a - b);
            const int pm = spos<NC>(NC / 2);
            const float2 xm = buf[pm];
            buf[pm] = make_float2(2.0f * xm.x, -2.0f * xm.y);
        } else {
            const int pa = spos<NC>(k), pb = spos<NC>(NC - k);
            const float2 xa = buf[pa], xb = buf[pb];
            const float2 e = make_float2(xa.x + xb.x, xa.y - xb.y);
            const float2 t = make_float2(xa.x - xb.x, xa.y + xb.y);
            const float2 o = cmulc(t, wsplit[k]);
            buf[pa] = make_float2(e.x - o.y, e.y + o.x);   // E2 + i O2
            buf[pb] = make_float2(e.x + o.y, o.x - e.y);   // conj(E2 - i O2)
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------- quantizer (one frame, one warp)
// Gather form of dsp/quantizer.py:424-527:
//   G_t = sum_{i -> t} |X_i|,  P_t = sum_{i -> t} X_i           (sources in ascending order)
//   tE_d = sum_e coef[d][e] G_{t(d,e)},  PS_d = sum_e coef[d][e] P_{t(d,e)}
//   new_d = |X_d| keep_d + tE_d ;  phasor_d = PS_d/|PS_d| if tE_d > 0 else X_d/|X_d|
//   out_d = smooth(new)_d * phasor_d      with [1/4,1/2,1/4], edges replicated
template <int NC>
QD_DEV void quantize_frame(float2 *buf, float *slotG, float2 *slotP, const QuantDev &q, int lane) {
    // Q1: per-target gathers
    for (int s = lane; s < q.n_slots; s += 32) {
        const int b = q.slot_begin[s], e = q.slot_begin[s + 1];
        float g = 0.0f;
        float2 p = make_float2(0.0f, 0.0f);
        for (int i = b; i < e; ++i) {
            const float2 xv = buf[spos<NC>(q.src_bin[i])];
            g += sqrtf(xv.x * xv.x + xv.y * xv.y);
            p = cadd(p, xv);
        }
        slotG[s] = g;
        slotP[s] = p;
    }
    if (lane == 0) {
        slotG[q.n_slots] = 0.0f;
        slotP[q.n_slots] = make_float2(0.0f, 0.0f);
    }
    __syncwarp();

    // Q3: rows of 32 bins, rolling window of three rows for the smoothing
    constexpr int NBINS = NC + 1;
    constexpr int ROWS = (NBINS + 31) / 32;
    float m_prev = 0.0f, m_cur = 0.0f, m_next = 0.0f;
    float2 u_cur = make_float2(1.0f, 0.0f), u_next = make_float2(1.0f, 0.0f);
#pragma unroll 1
    for (int row = 0; row <= ROWS; ++row) {
        // ---- compute row `row` into (m_next, u_next)
        m_next = 0.0f;
        u_next = make_float2(1.0f, 0.0f);
        const int d = 32 * row + lane;
        if (row < ROWS && d < NBINS) {
            const float2 xv = buf[spos<NC>(d)];
            const float m2 = xv.x * xv.x + xv.y * xv.y;
            const float m = sqrtf(m2);
            float2 u = make_float2(1.0f, 0.0f);  // np.angle(0) = 0
            if (m > 0.0f) {
                const float r = 1.0f / m;
                u = make_float2(xv.x * r, xv.y * r);
            }
            const uint32_t bit = 1u << lane;
            float nm = (q.row_active[row] & bit) ? m * q.keep_active : m;
            const uint32_t am = q.row_aff[row];
            if (am & bit) {
                const AffEntry &ae = q.aff[q.row_aff_base[row] + __popc(am & (bit - 1u))];
                float te = 0.0f;
                float2 ps = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int e = 0; e < 5; ++e) {
                    const int s = ae.slot[e];
                    const float c = ae.coef[e];
                    te += c * slotG[s];
                    const float2 pv = slotP[s];
                    ps.x += c * pv.x;
                    ps.y += c * pv.y;
                }
                nm += te;
                if (te > 0.0f) {
                    const float p2 = ps.x * ps.x + ps.y * ps.y;
                    if (p2 > 0.0f) {
                        const float r = rsqrtf(p2);
                        u = make_float2(ps.x * r, ps.y * r);
                    } else {
                        u = make_float2(1.0f, 0.0f);
                    }
                }
            }
            m_next = nm;
            u_next = u;
        }
        // ---- finish row-1 (needs its neighbours: lane-1/lane+1, across rows at the ends)
        if (row > 0) {
            const int dc = 32 * (row - 1) + lane;
            float out = m_cur;
            if (q.smoothing) {
                float left = __shfl_up_sync(QD_FULL, m_cur, 1);
                const float left_wrap = __shfl_sync(QD_FULL, m_prev, 31);
                float right = __shfl_down_sync(QD_FULL, m_cur, 1);
                const float right_wrap = __shfl_sync(QD_FULL, m_next, 0);
                if (lane == 0) left = left_wrap;
                if (lane == 31) right = right_wrap;
                if (dc == 0) left = m_cur;              // mode="nearest"
                if (dc == NBINS - 1) right = m_cur;
                out = 0.5f * m_cur + 0.25f * (left + right);
            }
            if (dc < NBINS) buf[spos<NC>(dc)] = make_float2(out * u_cur.x, out * u_cur.y);
        }
        m_prev = m_cur;
        m_cur = m_next;
        u_cur = u_next;
    }
    __syncwarp();
}

// ---------------------------------------------------------------- the kernel
template <int NC, int NW>
struct SpecSmem {
    static constexpr int HOP = NC / 2;                       // n_fft / 4 samples
    static constexpr int BUF = buf_slots<NC>();               // float2 per warp buffer
    static constexpr int STAGE = (NW + 3) * HOP;              // floats
    static constexpr int TAIL = 3 * HOP;                      // floats
    static constexpr size_t off_buf = 0;
    static constexpr size_t off_stage = off_buf + (size_t)NW * BUF * sizeof(float2);
    static constexpr size_t off_tail = off_stage + (size_t)STAGE * sizeof(float);
    static constexpr size_t off_flags = off_tail + (size_t)TAIL * sizeof(float);
    static constexpr size_t off_slot = off_flags + 64 * sizeof(int);
    // per-warp gather scratch: G[cap] floats then P[cap] float2, cap even so P stays 8-byte aligned
    static int slot_cap(int n_slots) { return (n_slots + 2) & ~1; }
    static size_t bytes(int n_slots) {
        return off_slot + (size_t)NW * (size_t)slot_cap(n_slots) * 3 * sizeof(float) + 16;
    }
};

// wavefold / tube on the float32 iSTFT sample (dsp/distortion.py:18-90)
QD_DEV float epilogue_apply(float v, int mode, float fold, float bias, float tg, float tn) {
    if (mode == 1) {
        // dsp/distortion.py:44-56: the negative fold is tested on the already folded value
        float y = (v + bias) * fold;
        if (y > 1.0f) y = 2.0f - y;
        if (y < -1.0f) y = -2.0f - y;
        return fminf(fmaxf(y, -1.0f), 1.0f);
    }
    if (mode == 2) return tanhf(tg * v) * tn;
    return v;
}

template <int NC, int NW>
__global__ void __launch_bounds__(32 * NW)
spec_pass_kernel(const SpecArgs a) {
    using L = SpecSmem<NC, NW>;
    constexpr int HOP = L::HOP;
    constexpr int NFFT = 2 * NC;
    QD_DYN_SMEM(smem);
    float2 *bufs = reinterpret_cast<float2 *>(smem + L::off_buf);
    float *stage = reinterpret_cast<float *>(smem + L::off_stage);
    float *tail = reinterpret_cast<float *>(smem + L::off_tail);
    int *flags = reinterpret_cast<int *>(smem + L::off_flags);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = 32 * NW;
    float2 *buf = bufs + (size_t)warp * L::BUF;
    const int slot_cap = (a.q.n_slots + 2) & ~1;
    float *slotG = reinterpret_cast<float *>(smem + L::off_slot) + (size_t)warp * slot_cap * 3;
    float2 *slotP = reinterpret_cast<float2 *>(slotG + slot_cap);

    const int clip = blockIdx.y;
    const float *x = a.x + (size_t)clip * a.n;
    float *y = a.y + (size_t)clip * a.n;
    float *tap = a.tap ? a.tap + (size_t)clip * a.n : nullptr;

    // output hop-blocks (of the zero-padded timeline) this CTA owns: [j0, j1); block 2 <-> sample 0
    const int j_end = 2 + (a.n + HOP - 1) / HOP;
    const int j0 = 2 + blockIdx.x * a.tile_blocks;
    const int j1 = min(j0 + a.tile_blocks, j_end);
    if (j0 >= j1) return;
    const int t_first = j0 - 3;  // first frame that touches block j0 (may be < 0: skipped)

    for (int i = tid; i < L::TAIL; i += nthreads) tail[i] = 0.0f;

    for (int tb = t_first; tb < j1; tb += NW) {
        // ---- stage the samples of frames tb .. tb+NW-1 (zero outside the clip)
        {
            const long long s0 = (long long)tb * HOP - NC;  // clip index of staging[0]
            for (int i = tid; i < L::STAGE; i += nthreads) {
                const long long s = s0 + i;
                stage[i] = (s >= 0 && s < a.n) ? x[s] : 0.0f;
            }
        }
        __syncthreads();
        // ---- one frame per warp
        const int t = tb + warp;
        const bool live = (t >= 0 && t < a.n_frames);
        if (lane == 0) flags[warp] = live ? 1 : 0;
        if (live) {
            const float2 *frame = reinterpret_cast<const float2 *>(stage + warp * HOP);
            fft_forward<NC>(buf, frame, a, a.wtab, a.tw1, a.tw2, lane);
            real_split<NC>(buf, a.wsplit, lane);
            if (a.quant) quantize_frame<NC>(buf, slotG, slotP, a.q, lane);
            real_merge<NC>(buf, a.wsplit, lane);
            fft_inverse<NC>(buf, a.wtab, a.tw1, a.tw2, lane);
        }
        __syncthreads();
        // ---- overlap-add in frame order; blocks tb .. tb+NW-1 are now complete
        for (int c = tid; c < HOP; c += nthreads) {
            float carry[3];
#pragma unroll
            for (int g = 0; g < 3; ++g) carry[g] = tail[g * HOP + c];
            // sample c of hop-slice `sl` of warp w's frame
            auto fr = [&](int w, int sl) -> float {
                const int s = sl * HOP + c;
                return reinterpret_cast<const float *>(bufs + (size_t)w * L::BUF)[2 * pidx(s >> 1) + (s & 1)];
            };
#pragma unroll 1
            for (int h = 0; h < NW + 3; ++h) {
                float v = (h < 3) ? carry[h] : 0.0f;
                const int w_lo = h - 3 > 0 ? h - 3 : 0;
                const int w_hi = h < NW - 1 ? h : NW - 1;
                for (int w = w_lo; w <= w_hi; ++w)
                    if (flags[w]) v += fr(w, h - w);
                if (h >= NW) {
                    tail[(h - NW) * HOP + c] = v;  // partial sums of the next three blocks
                    continue;
                }
                const int j = tb + h;
                if (j < j0 || j >= j1) continue;
                const long long nidx = (long long)(j - 2) * HOP + c;
                if (nidx >= a.n) continue;
                // frames covering block j: slices sl = j - t with t in [max(0,j-3), min(j,T-1)]
                const int sl_a = j - a.n_frames + 1 > 0 ? j - a.n_frames + 1 : 0;
                const int sl_b = j < 3 ? j : 3;
                const float inv = sl_a <= sl_b ? a.invw[(sl_a * 4 + sl_b) * HOP + c] : 0.0f;
                const float o = v * inv;
                if (tap) tap[nidx] = o;
                y[nidx] = epilogue_apply(o, a.epilogue, a.fold, a.bias, a.tube_gain, a.tube_norm);
            }
        }
        __syncthreads();
    }
    (void)NFFT;
}

}  // namespace qd
