// throughput of the float64 tensor-core shapes and of plain DFMA on one GPU
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double seed) {
    double c[NACC][4];
#pragma unroll
    for (int u = 0; u < NACC; ++u) for (int v = 0; v < 4; ++v) c[u][v] = 0.0;
    double a[8], b[4];
    for (int v = 0; v < 8; ++v) a[v] = seed + threadIdx.x * 1e-3 + v;
    for (int v = 0; v < 4; ++v) b[v] = seed - threadIdx.x * 1e-3 - v;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < NACC; ++u) {
            if constexpr (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a[u & 7]), "d"(b[u & 3]));
            } else if constexpr (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3]) : "d"(a[u & 7]), "d"(a[(u + 1) & 7]), "d"(b[u & 3]));
            } else if constexpr (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[u & 1]), "d"(b[2 + (u & 1)]));
            } else {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < NACC; ++u) for (int v = 0; v < 4; ++v) s += c[u][v];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double seed) {
    double c[NACC];
#pragma unroll
    for (int u = 0; u < NACC; ++u) c[u] = u;
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < NACC; ++u) c[u] = fma(c[u], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < NACC; ++u) s += c[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double *out; CK(cudaMalloc(&out, 148 * 8 * 256 * 8 * sizeof(double)));
    const int iters = 20000;
    const int macs[4] = {256, 512, 1024, 2048};
    const char *names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    for (int wps = 4; wps <= 16; wps *= 2) {   // warps per SM (one CTA per SM of 32*wps threads... use CTAs of 128)
        const int ctas = 148 * wps / 4;
        float ms[4];
        ms[0] = time_it([&] { k_dmma<0, 8><<<ctas, 128>>>(out, iters, 0.5); });
        ms[1] = time_it([&] { k_dmma<1, 8><<<ctas, 128>>>(out, iters, 0.5); });
        ms[2] = time_it([&] { k_dmma<2, 8><<<ctas, 128>>>(out, iters, 0.5); });
        ms[3] = time_it([&] { k_dmma<3, 8><<<ctas, 128>>>(out, iters, 0.5); });
        for (int s = 0; s < 4; ++s) {
            const double flop = 2.0 * macs[s] * 8.0 * iters * (double)ctas * 4;
            printf("warps/SM %2d  %-9s %8.3f ms  %7.2f TFLOP/s\n", wps, names[s], ms[s], flop / ms[s] * 1e-9);
        }
        const float md = time_it([&] { k_dfma<8><<<ctas, 128>>>(out, iters * 8, 0.5); });
        printf("warps/SM %2d  %-9s %8.3f ms  %7.2f TFLOP/s\n", wps, "dfma", md, 2.0 * 8 * iters * 8 * (double)ctas * 128 / md * 1e-9);
    }
    CK(cudaGetLastError());
    return 0;
}
