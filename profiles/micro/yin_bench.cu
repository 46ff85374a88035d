// times at_yin_diff_kernel<L> on synthetic clips: yin_bench [clips]
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../../quantumdistortion_b200/csrc/qd_yin.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

static int run(const float *det, double *diff, int clips, long long n, int max_tau, std::vector<double> *first) {
    constexpr int L = qd::AT_YL;
    qd::AtYinArgs a{};
    a.det = det; a.diff = diff; a.n = n; a.frame_size = 4096; a.hop = 512; a.max_tau = max_tau;
    a.frames = (int)((n + 511) / 512); a.stride = (max_tau + 2) & ~1;
    const int total_cols = (max_tau + L - 1) / L;
    const int tblocks = (total_cols + qd::AT_YC - 1) / qd::AT_YC;
    a.lag_threads = (total_cols + tblocks - 1) / tblocks;
    const int threads = qd::AT_YT;
    const size_t smem = qd::at_yin_smem_doubles(a.hop, (tblocks - 1) * a.lag_threads, a.lag_threads) * sizeof(double) + 128;
    CK(cudaFuncSetAttribute(qd::at_yin_diff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        qd::at_yin_diff_kernel<<<dim3(tblocks, clips), threads, smem>>>(a);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    first->resize((size_t)a.frames * a.stride);
    CK(cudaMemcpy(first->data(), diff, first->size() * sizeof(double), cudaMemcpyDeviceToHost));
    const double pairs = (double)clips * (a.frames + 8) * 512.0 * max_tau;
    printf("clips %d threads %d smem %zu KB: %.3f ms  (%.2f TFLOP/s of DFMA on the useful pairs)\n", clips, threads, smem / 1024, best,
           2.0 * pairs / best * 1e-9);
    return 0;
}

int main(int argc, char **argv) {
    const int clips = argc > 1 ? atoi(argv[1]) : 1024;
    const long long n = 480000;
    const int max_tau = 671;
    std::vector<float> h((size_t)n);
    for (long long i = 0; i < n; ++i) h[i] = 0.5f * sinf(0.0131f * i) + 0.2f * sinf(0.00217f * i * (1.0f + 1e-6f * i));
    float *det; double *diff;
    CK(cudaMalloc(&det, (size_t)clips * n * sizeof(float)));
    for (int c = 0; c < clips; ++c) CK(cudaMemcpy(det + (size_t)c * n, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    const size_t stride = (max_tau + 2) & ~1;
    CK(cudaMalloc(&diff, (size_t)clips * ((n + 511) / 512) * stride * sizeof(double)));
    std::vector<double> d;
    if (run(det, diff, clips, n, max_tau, &d)) return 1;
    // spot check against the definition on a few (frame, lag) entries
    double worst = 0.0;
    for (int f : {0, 1, 7, 8, 100, 500, 930, 937})
        for (int tau : {1, 2, 15, 16, 17, 100, 333, 670, 671}) {
            double ref = 0.0;
            for (int j = 0; j < 4096 - tau; ++j) {
                const long long i0 = (long long)f * 512 + j, i1 = i0 + tau;
                const double x0 = i0 < n ? (double)h[i0] : 0.0, x1 = i1 < n ? (double)h[i1] : 0.0;
                ref += (x0 - x1) * (x0 - x1);
            }
            worst = fmax(worst, fabs(d[(size_t)f * stride + tau] - ref) / fmax(ref, 1e-30));
        }
    printf("spot check against the definition: max relative error %.3e\n", worst);
    return 0;
}
