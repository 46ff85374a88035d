// DFMA rate of a register-only L x L outer-product update (the YIN walk's arithmetic without its loads)
#include <cstdio>
#include <cuda_runtime.h>
template <int L>
__global__ void __launch_bounds__(256) k_outer(double *out, int iters, const double *in) {
    double S[L], av[L], w[2 * L - 1];
#pragma unroll
    for (int u = 0; u < L; ++u) { S[u] = 0.0; av[u] = in[threadIdx.x + u]; }
#pragma unroll
    for (int v = 0; v < 2 * L - 1; ++v) w[v] = in[threadIdx.x + 64 + v];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < L; ++k)
#pragma unroll
            for (int u = 0; u < L; ++u) S[u] = fma(av[k], w[k + u], S[u]);
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < L; ++u) s += S[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int L> static void go(double *out, const double *in, int wps) {
    const int iters = L == 8 ? 8000 : 2000;
    const int ctas = 148 * wps / 4;   // CTAs of 128 threads
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_outer<L><<<ctas, 128>>>(out, iters, in); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_outer<L><<<ctas, 128>>>(out, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("L=%2d warps/SM %2d: %7.3f ms %6.2f TFLOP/s\n", L, wps, ms, 2.0 * L * L * iters * (double)ctas * 128 / ms * 1e-9);
}
int main() {
    double *out, *in; cudaMalloc(&out, 148 * 16 * 128 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    for (int wps : {4, 8, 12, 16, 24}) { go<8>(out, in, wps); go<16>(out, in, wps); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
