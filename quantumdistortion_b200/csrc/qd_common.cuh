// qd_common.cuh -- shared definitions for the sm_100a kernels of libqd_b200.so.
//
// The same headers compile two ways:
//   nvcc  -gencode arch=compute_100a,code=sm_100a   -> the product (qd_api.cu)
//   g++   -DQD_EMU (tests/emu/cuda_emu.h)            -> CPU single-stepping of the kernel
//          source for index-math tests in the GPU-less build container (test only).
#pragma once

#ifdef QD_EMU
#include "cuda_emu.h"
#define QD_SINCOSF(x, s, c) sincosf_emu((x), (s), (c))
#define QD_EXP10F(x) exp10f_emu(x)
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define QD_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define QD_SINCOSF(x, s, c) sincosf((x), (s), (c))
#define QD_EXP10F(x) exp10f(x)
#endif

#define QD_DEV __device__ __forceinline__
#define QD_FULL 0xffffffffu

namespace qd {

// ---------------------------------------------------------------- complex helpers
QD_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
QD_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
QD_DEV float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
QD_DEV float2 cmulc(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
QD_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
QD_DEV float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// cos(2*pi*k/32), k = 0..8 (quarter wave); everything else by symmetry
__host__ __device__ constexpr float qd_cos32_q(int k) {
    return k == 0 ? 1.0f
         : k == 1 ? 0.98078528040323043f
         : k == 2 ? 0.92387953251128674f
         : k == 3 ? 0.83146961230254524f
         : k == 4 ? 0.70710678118654752f
         : k == 5 ? 0.55557023301960218f
         : k == 6 ? 0.38268343236508978f
         : k == 7 ? 0.19509032201612825f
                  : 0.0f;
}
__host__ __device__ constexpr float qd_cos32(int k) {  // k in [0,32)
    return k <= 8 ? qd_cos32_q(k) : k <= 16 ? -qd_cos32_q(16 - k) : k <= 24 ? -qd_cos32_q(k - 16) : qd_cos32_q(32 - k);
}
__host__ __device__ constexpr float qd_sin32(int k) { return qd_cos32((k + 24) & 31); }  // sin(a) = cos(a - pi/2)

// d * exp(DIR * 2*pi*i * k/32);  k is a compile-time constant after unrolling.
template <int DIR>
QD_DEV float2 mul_w32(float2 d, int k) {
    k &= 31;
    if (k == 0) return d;
    if (k == 16) return make_float2(-d.x, -d.y);
    if (k == 8) return DIR > 0 ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
    if (k == 24) return DIR > 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
    const float c = qd_cos32(k);
    const float s = DIR > 0 ? qd_sin32(k) : -qd_sin32(k);
    return make_float2(d.x * c - d.y * s, d.x * s + d.y * c);
}

__host__ __device__ constexpr int qd_log2(int r) { return r <= 1 ? 0 : 1 + qd_log2(r >> 1); }
__host__ __device__ constexpr int qd_bitrev(int v, int bits) {
    int o = 0;
    for (int i = 0; i < bits; ++i) o |= ((v >> i) & 1) << (bits - 1 - i);
    return o;
}

// In-register radix-2 decimation-in-frequency DFT of R points (R = 2..32), fully unrolled.
// Input natural order; on return v[r] holds output index qd_bitrev(r, log2 R).
// DIR = -1: forward (exp(-2 pi i nk/R)); DIR = +1: inverse, unnormalised.
template <int R, int DIR>
QD_DEV void dft_reg(float2 (&v)[R]) {
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float2 a = v[base + i];
                const float2 b = v[base + i + half];
                v[base + i] = cadd(a, b);
                v[base + i + half] = mul_w32<DIR>(csub(a, b), i * (16 / half));
            }
        }
    }
}

}  // namespace qd
