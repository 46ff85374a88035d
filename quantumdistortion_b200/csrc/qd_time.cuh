// qd_time.cuh -- time-domain kernels of the STFT path: lookahead limiter + mix, Linkwitz-Riley
// crossover + low-band processing, elementwise distortion.
//
// Both recurrences are run as parallel scans in float64:
//   limiter   u_n = c * max(u_{n-1}, e_n)              dsp/limiter.py:62-77
//             associative on pairs (a, b): u -> max(a*u, b)
//   biquad    z_{n+1} = A z_n + B x_n (DF2T state)     scipy sosfilt, dsp/crossover.py:96-97
//             associative on pairs (A^k, v): z -> A^k z + v
// One CTA streams one clip in chunks; every thread owns KS consecutive samples, runs the
// recurrence over them from a zero state, the per-thread aggregates are combined with warp
// shuffles (+ one shared-memory hop across warps), and the thread re-runs its samples from
// its true incoming state.  The chunk-to-chunk state enters through thread 0.
#pragma once
#include "qd_common.cuh"

namespace qd {

constexpr int QD_TT = 256;          // threads per CTA
constexpr int QD_KS = 8;            // samples per thread per chunk
constexpr int QD_CHUNK = QD_TT * QD_KS;

QD_DEV int padi(int i) { return i + (i >> 5); }

struct LimiterArgs {
    const float *x;       // [batch, n] limiter input (x_post_quant)
    const float *dry;     // [batch, n] dry signal for the mix (tap_input)
    const float *low;     // optional [batch, n] processed low band (multiband recombine)
    const float *orig;    // optional [batch, n] original input (delta listen)
    float *y;             // [batch, n]
    long long n;
    int limiter_on;
    int lookahead;        // L >= 1
    double ceiling;
    double c;             // release coefficient
    float wet, dry_gain, trim;
    int apply_mix;        // 0: y = limited (stage-level entry point)
    int apply_trim;
};

// dsp/limiter.py:14-80 then dsp/pipeline.py:894-910, 1096, 1371-1375
__global__ void __launch_bounds__(QD_TT) limiter_mix_kernel(const LimiterArgs a) {
    QD_DYN_SMEM(smem);
    float *s_abs = reinterpret_cast<float *>(smem);  // padded |x| for [chunk, chunk + CHUNK + L)
    const int L = a.lookahead;
    const int span = QD_CHUNK + L + 8;
    float *s_gmax = s_abs + padi(span) + 8;          // max of each group of KS samples
    double *s_warp = reinterpret_cast<double *>(s_gmax + (span / QD_KS + 2));
    s_warp = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(s_warp) + 7) & ~(uintptr_t)7);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t base = (size_t)blockIdx.x * (size_t)a.n;
    const float *x = a.x + base;
    const double c = a.c;
    double c_ks = 1.0;
#pragma unroll
    for (int k = 0; k < QD_KS; ++k) c_ks *= c;
    double carry = 0.0;  // u at the end of the previous chunk

    for (long long n0 = 0; n0 < a.n; n0 += QD_CHUNK) {
        float xv[QD_KS];
        double u_in = 0.0;
        double e[QD_KS];
        if (a.limiter_on) {
            // |x| for the chunk and its lookahead (zero past the end: the reference window is clipped)
            for (int i = tid; i < span; i += QD_TT) {
                const long long s = n0 + i;
                s_abs[padi(i)] = s < a.n ? fabsf(x[s]) : 0.0f;
            }
            __syncthreads();
            const int ngroups = span / QD_KS;
            for (int g = tid; g < ngroups; g += QD_TT) {
                float m = 0.0f;
#pragma unroll
                for (int k = 0; k < QD_KS; ++k) m = fmaxf(m, s_abs[padi(g * QD_KS + k)]);
                s_gmax[g] = m;
            }
            __syncthreads();
            // forward-window maxima of this thread's KS samples
            float own[QD_KS];
#pragma unroll
            for (int k = 0; k < QD_KS; ++k) own[k] = s_abs[padi(tid * QD_KS + k)];
            float peak[QD_KS];
            if (L >= QD_KS) {
                const int q = L / QD_KS, rem = L % QD_KS;
                float mc = 0.0f;  // groups tid+1 .. tid+q-1 lie inside every window
                for (int g = tid + 1; g < tid + q; ++g) mc = fmaxf(mc, s_gmax[g]);
                float ga[QD_KS], gb[QD_KS];
#pragma unroll
                for (int k = 0; k < QD_KS; ++k) {
                    ga[k] = s_abs[padi((tid + q) * QD_KS + k)];
                    gb[k] = s_abs[padi((tid + q + 1) * QD_KS + k)];
                }
                float suf = 0.0f;
                float sufv[QD_KS];
#pragma unroll
                for (int k = QD_KS - 1; k >= 0; --k) { suf = fmaxf(suf, own[k]); sufv[k] = suf; }
                float pa[QD_KS + 1], pb[QD_KS + 1];  // prefix maxima of length j
                pa[0] = 0.0f; pb[0] = 0.0f;
#pragma unroll
                for (int k = 0; k < QD_KS; ++k) { pa[k + 1] = fmaxf(pa[k], ga[k]); pb[k + 1] = fmaxf(pb[k], gb[k]); }
#pragma unroll
                for (int r = 0; r < QD_KS; ++r) {
                    // window [8t+r, 8t+r+L): tail of own group, common groups, head of group t+q (+ t+q+1)
                    float m = fmaxf(sufv[r], mc);
                    const int t = r + rem;
                    if (t < QD_KS) m = fmaxf(m, pa[t]);
                    else m = fmaxf(m, fmaxf(pa[QD_KS], pb[t - QD_KS]));
                    peak[r] = m;
                }
            } else {
#pragma unroll
                for (int r = 0; r < QD_KS; ++r) {
                    float m = 0.0f;
                    for (int k = 0; k < L; ++k) m = fmaxf(m, s_abs[padi(tid * QD_KS + r + k)]);
                    peak[r] = m;
                }
            }
            // e_n = 1 - ceiling/peak where the limiter engages  (dsp/limiter.py:68-72)
            double b = 0.0;
#pragma unroll
            for (int k = 0; k < QD_KS; ++k) {
                const double p = (double)peak[k];
                e[k] = (p > a.ceiling && p > 1e-12) ? 1.0 - a.ceiling / p : 0.0;
            }
            // local aggregate from the chunk carry (thread 0) or zero
            double u = (tid == 0) ? carry : 0.0;
#pragma unroll
            for (int k = 0; k < QD_KS; ++k) u = c * fmax(u, e[k]);
            b = u;
            // inclusive scan over threads: b_t = max(c_ks^d * b_{t-d}, b_t)
            double apow = c_ks;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double o = __shfl_up_sync(QD_FULL, b, d);
                if (lane >= d) b = fmax(apow * o, b);
                apow *= apow;
            }
            if (lane == 31) s_warp[warp] = b;
            __syncthreads();
            // apow == c_ks^32 here: combine warp totals
            double wprev = 0.0;  // aggregate of all warps before this one
            {
                double acc = 0.0;
                for (int w = 0; w < warp; ++w) acc = fmax(apow * acc, s_warp[w]);
                wprev = acc;
            }
            // exclusive value for this thread: aggregate of threads < tid
            double excl = __shfl_up_sync(QD_FULL, b, 1);
            double lanepow = 1.0;  // c_ks^lane
            {
                double p = c_ks;
                int l = lane;
                while (l) { if (l & 1) lanepow *= p; p *= p; l >>= 1; }
            }
            if (lane == 0) excl = 0.0;
            u_in = fmax(lanepow * wprev, excl);
            if (tid == 0) u_in = carry;
            // chunk carry for the next iteration = inclusive value of the last thread
            double total = fmax(apow * wprev, s_warp[warp]);  // valid in the last warp
            __syncthreads();
            if (tid == QD_TT - 1) s_warp[0] = total;
            __syncthreads();
            carry = s_warp[0];
            __syncthreads();
        }
        // apply: y = float32(x * g), mix, trim, recombine, delta
        const long long s0 = n0 + (long long)tid * QD_KS;
        double u = u_in;
#pragma unroll
        for (int k = 0; k < QD_KS; ++k) {
            const long long s = s0 + k;
            if (s >= a.n) break;
            xv[k] = x[s];
            float w = xv[k];
            if (a.limiter_on) {
                u = c * fmax(u, e[k]);
                double g = 1.0 - u;
                g = fmin(fmax(g, 0.0), 1.0);
                w = (float)((double)xv[k] * g);
            }
            if (a.apply_mix) {
                w = __fadd_rn(__fmul_rn(a.wet, w), __fmul_rn(a.dry_gain, a.dry[base + s]));
                if (a.apply_trim) w = __fmul_rn(w, a.trim);
                if (a.low) w = __fadd_rn(a.low[base + s], w);
                if (a.orig) w = __fsub_rn(a.orig[base + s], w);
            }
            a.y[base + s] = w;
        }
    }
}

// ---------------------------------------------------------------- crossover
struct Mat2 { double a, b, c, d; };   // [[a b],[c d]]
struct CrossoverArgs {
    const float *x;       // [batch, n]
    float *low;           // [batch, n]
    float *high;          // [batch, n]
    long long n;
    double co[4][6];      // sections lp1, lp2, hp1, hp2 as (b0 b1 b2 1 a1 a2)
    Mat2 apow[4][6];      // per section A^(KS * 2^l), l = 0..5, A = [[-a1, 1], [-a2, 0]]
    int low_delay;        // samples the low band is delayed by (dsp/pipeline.py:389-396)
    int process_low;      // 1: saturate / blend / trim the low band (dsp/pipeline.py:1063-1073)
    float low_gain;
    double low_norm;
    float mono_a, mono_b;
    int apply_mono;
    float low_trim;
    int apply_low_trim;
};

QD_DEV void mat_apply(const Mat2 &m, double &z0, double &z1) {
    const double t0 = m.a * z0 + m.b * z1;
    const double t1 = m.c * z0 + m.d * z1;
    z0 = t0; z1 = t1;
}

// one DF2T section over the thread's KS samples, in place on v[], state (z0,z1) in/out
QD_DEV void biquad_run(const double *co, double (&v)[QD_KS], double &z0, double &z1) {
#pragma unroll
    for (int k = 0; k < QD_KS; ++k) {
        const double xin = v[k];
        const double o = co[0] * xin + z0;
        z0 = co[1] * xin - co[4] * o + z1;
        z1 = co[2] * xin - co[5] * o;
        v[k] = o;
    }
}

// state entering each thread for one section whose input is in[]; chunk state enters at thread 0
QD_DEV void biquad_scan(const double *co, const Mat2 *apow, const double (&in)[QD_KS], double cz0, double cz1,
                        double *s_w, int tid, double &zin0, double &zin1, double &zend0, double &zend1) {
    const int lane = tid & 31, warp = tid >> 5;
    double tmp[QD_KS];
#pragma unroll
    for (int k = 0; k < QD_KS; ++k) tmp[k] = in[k];
    double v0 = (tid == 0) ? cz0 : 0.0, v1 = (tid == 0) ? cz1 : 0.0;
    biquad_run(co, tmp, v0, v1);
    // inclusive scan: v_t = A^(KS*d) v_{t-d} + v_t
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        const int d = 1 << l;
        double o0 = __shfl_up_sync(QD_FULL, v0, d);
        double o1 = __shfl_up_sync(QD_FULL, v1, d);
        if (lane >= d) { mat_apply(apow[l], o0, o1); v0 += o0; v1 += o1; }
    }
    if (lane == 31) { s_w[2 * warp] = v0; s_w[2 * warp + 1] = v1; }
    __syncthreads();
    double p0 = 0.0, p1 = 0.0;  // aggregate of the warps before this one
    for (int w = 0; w < warp; ++w) { mat_apply(apow[5], p0, p1); p0 += s_w[2 * w]; p1 += s_w[2 * w + 1]; }
    // exclusive within the warp
    double e0 = __shfl_up_sync(QD_FULL, v0, 1), e1 = __shfl_up_sync(QD_FULL, v1, 1);
    if (lane == 0) { e0 = 0.0; e1 = 0.0; }
    // A^(KS*lane) applied to the warp-prefix
    double q0 = p0, q1 = p1;
#pragma unroll
    for (int l = 0; l < 5; ++l) if (lane & (1 << l)) mat_apply(apow[l], q0, q1);
    zin0 = q0 + e0; zin1 = q1 + e1;
    if (tid == 0) { zin0 = cz0; zin1 = cz1; }
    // end-of-chunk state = inclusive value of the last thread
    double t0 = p0, t1 = p1;
    mat_apply(apow[5], t0, t1);
    t0 += s_w[2 * warp]; t1 += s_w[2 * warp + 1];
    __syncthreads();
    if (tid == QD_TT - 1) { s_w[0] = t0; s_w[1] = t1; }
    __syncthreads();
    zend0 = s_w[0]; zend1 = s_w[1];
    __syncthreads();
}

QD_DEV float low_process(float v, const CrossoverArgs &a) {
    // dsp/saturation.py:44-54: float32 gain, float32 tanh(3x), float64 division by tanh(3)
    const float t = tanhf(__fmul_rn(3.0f, __fmul_rn(v, a.low_gain)));
    float r = (float)((double)t * a.low_norm);
    if (a.apply_mono) r = __fadd_rn(__fmul_rn(a.mono_a, r), __fmul_rn(a.mono_b, r));
    if (a.apply_low_trim) r = __fmul_rn(r, a.low_trim);
    return r;
}

__global__ void __launch_bounds__(QD_TT) crossover_kernel(const CrossoverArgs a) {
    __shared__ double s_w[2 * (QD_TT / 32) + 2];
    const int tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * (size_t)a.n;
    const float *x = a.x + base;
    double st[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // chunk states: lp1, lp2, hp1, hp2
    if (a.low_delay > 0)
        for (long long i = tid; i < a.low_delay && i < a.n; i += QD_TT) a.low[base + i] = a.process_low ? low_process(0.0f, a) : 0.0f;
    for (long long n0 = 0; n0 < a.n; n0 += QD_CHUNK) {
        const long long s0 = n0 + (long long)tid * QD_KS;
        double xin[QD_KS];
#pragma unroll
        for (int k = 0; k < QD_KS; ++k) xin[k] = (s0 + k < a.n) ? (double)x[s0 + k] : 0.0;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            double v[QD_KS];
#pragma unroll
            for (int k = 0; k < QD_KS; ++k) v[k] = xin[k];
#pragma unroll
            for (int sec = 0; sec < 2; ++sec) {
                double z0, z1, e0, e1;
                const double *co = a.co[2 * f + sec];
                biquad_scan(co, a.apow[2 * f + sec], v, st[2 * f + sec][0], st[2 * f + sec][1], s_w, tid, z0, z1, e0, e1);
                biquad_run(co, v, z0, z1);
                st[2 * f + sec][0] = e0;
                st[2 * f + sec][1] = e1;
            }
#pragma unroll
            for (int k = 0; k < QD_KS; ++k) {
                const long long s = s0 + k;
                if (s >= a.n) break;
                const float o = (float)v[k];
                if (f == 0) {
                    const long long dsts = s + a.low_delay;
                    if (dsts < a.n) a.low[base + dsts] = a.process_low ? low_process(o, a) : o;
                } else {
                    a.high[base + s] = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- elementwise
__global__ void distort_kernel(const float *__restrict__ x, float *__restrict__ y, long long count, int mode,
                               float fold, float bias, float tg, float tn) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float v = x[i];
        float r;
        if (mode == 0) {
            float t = (v + bias) * fold;
            if (t > 1.0f) t = 2.0f - t;
            if (t < -1.0f) t = -2.0f - t;  // sequential masks, dsp/distortion.py:44-53
            r = fminf(fmaxf(t, -1.0f), 1.0f);
        } else {
            r = tanhf(tg * v) * tn;
        }
        y[i] = r;
    }
}

__global__ void add_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ y,
                           long long count) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        y[i] = __fadd_rn(a[i], b[i]);
}

}  // namespace qd
