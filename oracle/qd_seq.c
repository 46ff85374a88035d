/*
 * oracle/qd_seq.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the two strictly sequential loops on the reference's
 * STFT hot path.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load this file.  The product path
 * (quantumdistortion_b200/) never links or calls it.
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off, see Makefile)
 *
 * -ffp-contract=off keeps every multiply/add a separately rounded IEEE-754
 * double operation, so the arithmetic below is bit-identical to the
 * reference's per-sample Python/NumPy float64 evaluation.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

/*
 * Lookahead peak limiter gain curve.
 * Follows quantum_distortion/dsp/limiter.py:62-77 step by step:
 *   peak_n = max |x[n : min(N, n+L)]|
 *   if peak_n > ceiling and peak_n > 1e-12:  g = min(g, ceiling / peak_n)
 *   g = 1 - (1 - g) * release_coeff ;  g = clip(g, 0, 1) ;  gain[n] = g
 * The forward-window maximum is kept in a monotonic deque (same value the
 * reference obtains with np.max over the slice; max is exact, so the order in
 * which it is evaluated does not matter).
 */
void qd_oracle_limiter_gain(const double *x, int64_t n, int64_t lookahead,
                            double ceiling_lin, double release_coeff,
                            double *gain)
{
    if (n <= 0) return;
    if (lookahead < 1) lookahead = 1;
    int64_t *dq = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    int64_t head = 0, tail = 0; /* dq[head..tail) holds indices, |x| decreasing */
    int64_t pushed = 0;         /* next index to push */
    double g = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t end = i + lookahead;
        if (end > n) end = n;
        while (pushed < end) {
            double a = fabs(x[pushed]);
            while (tail > head && fabs(x[dq[tail - 1]]) <= a) --tail;
            dq[tail++] = pushed++;
        }
        while (dq[head] < i) ++head;
        double peak = fabs(x[dq[head]]);
        if (peak > ceiling_lin && peak > 1e-12) {
            double desired = ceiling_lin / peak;
            if (desired < g) g = desired;
        }
        g = 1.0 - (1.0 - g) * release_coeff;
        if (g < 0.0) g = 0.0;
        if (g > 1.0) g = 1.0;
        gain[i] = g;
    }
    free(dq);
}

/*
 * Cascade of second-order sections, transposed direct form II, float64 state.
 * Restates the arithmetic scipy.signal.sosfilt performs for the call at
 * quantum_distortion/dsp/crossover.py:96-97 (scipy is a third-party
 * dependency of the reference, un-pinned in requirements.txt; this container
 * has scipy 1.18.1, against which tests/test_oracle.py checks this loop):
 *   for each sample, for each section s in order:
 *       y  = b0*x + z0
 *       z0 = b1*x - a1*y + z1
 *       z1 = b2*x - a2*y
 *       x  = y
 * sos is row-major [n_sections][6] = b0 b1 b2 a0 a1 a2 with a0 == 1.
 */
void qd_oracle_sosfilt(const double *sos, int n_sections, const double *x,
                       int64_t n, double *y)
{
    double z[16][2];
    if (n_sections > 16) n_sections = 16;
    for (int s = 0; s < n_sections; ++s) z[s][0] = z[s][1] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double v = x[i];
        for (int s = 0; s < n_sections; ++s) {
            const double *c = sos + 6 * s;
            double o = c[0] * v + z[s][0];
            z[s][0] = c[1] * v - c[4] * o + z[s][1];
            z[s][1] = c[2] * v - c[5] * o;
            v = o;
        }
        y[i] = v;
    }
}
