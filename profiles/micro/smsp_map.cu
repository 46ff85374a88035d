// which warps of a CTA share an SM sub-partition?  time DFMA streams on chosen warp subsets
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(double *out, int iters, unsigned mask, const double *in) {
    const int warp = threadIdx.x >> 5;
    if (!((mask >> warp) & 1)) return;
    double S[16], av[16], w[31];
#pragma unroll
    for (int u = 0; u < 16; ++u) { S[u] = 0.0; av[u] = in[threadIdx.x + u]; }
#pragma unroll
    for (int v = 0; v < 31; ++v) w[v] = in[threadIdx.x + 64 + v];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
#pragma unroll
            for (int u = 0; u < 16; ++u) S[u] = fma(av[kk], w[kk + u], S[u]);
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < 16; ++u) s += S[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *out, *in; cudaMalloc(&out, 148 * 384 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    const unsigned masks[] = {0x1, 0x3, 0x5, 0x11, 0x101, 0xf, 0x33, 0x55, 0xff, 0xfff, 0x0f0f >> 4, 0x303, 0x30};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (unsigned m : masks) {
        k<<<148, 384>>>(out, 2000, m, in); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<148, 384>>>(out, 2000, m, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("mask 0x%03x (%2d warps): %.3f ms\n", m, __builtin_popcount(m), ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
