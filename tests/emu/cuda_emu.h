// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE.
//
// A tiny host emulation of the CUDA execution model, so that the kernel source in
// quantumdistortion_b200/csrc/*.cuh can be compiled with g++ and single-stepped on the
// CPU-only build container (index math, table layout, branch shapes) before GPU time
// is spent.  One OS thread per CUDA thread, pthread barriers for __syncthreads /
// __syncwarp, warp shuffles through a per-warp exchange buffer.  It is slow and only
// ever runs a handful of blocks.  Nothing in the product path includes this file; the
// shipped library is built by nvcc from the same .cuh files with QD_EMU undefined.
#pragma once
#ifndef QD_EMU
#error "cuda_emu.h is only for -DQD_EMU host builds"
#endif

#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return double2{a, b}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static  /* blocks run one at a time, so one static instance per block */

namespace qd_emu {
struct Block {
    dim3 grid, block;
    pthread_barrier_t bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<uint64_t> xchg;  // [warps][32]
    int named_count[16] = {0};
    unsigned named_gen[16] = {0};
    pthread_mutex_t named_mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_cond_t named_cv = PTHREAD_COND_INITIALIZER;
    std::vector<unsigned char> smem;
};
extern thread_local Block *g_blk;
extern thread_local dim3 g_tid, g_bid;
inline int lane() { return (int)(g_tid.x & 31u); }
inline int warp() { return (int)(g_tid.x >> 5); }

template <class T>
inline T shfl_idx(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Block *b = g_blk;
    uint64_t *x = &b->xchg[(size_t)warp() * 32];
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    x[lane()] = raw;
    pthread_barrier_wait(&b->warp_bar[warp()]);
    uint64_t got = x[src & 31];
    pthread_barrier_wait(&b->warp_bar[warp()]);
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}

// bar.sync id, count / bar.arrive id, count: `count` threads take part in one generation of barrier `id`; a syncing thread
// waits for the generation to complete, an arriving thread only counts (producer / consumer hand-offs)
inline void named_barrier_op(int id, int count, bool wait) {
    Block *b = g_blk;
    pthread_mutex_lock(&b->named_mu);
    const unsigned gen = b->named_gen[id];
    if (++b->named_count[id] == count) {
        b->named_count[id] = 0;
        ++b->named_gen[id];
        pthread_cond_broadcast(&b->named_cv);
    } else if (wait) {
        while (b->named_gen[id] == gen) pthread_cond_wait(&b->named_cv, &b->named_mu);
    }
    pthread_mutex_unlock(&b->named_mu);
}
inline void named_barrier(int id, int count) { named_barrier_op(id, count, true); }
inline void named_arrive(int id, int count) { named_barrier_op(id, count, false); }

// Run `body` once per CUDA thread of every block of the grid (blocks sequentially).
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    unsigned nthreads = block.x * block.y * block.z;
    unsigned nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Block blk;
                blk.grid = grid;
                blk.block = block;
                blk.smem.assign(smem_bytes + 256, 0xCD);  // poison: reads of unwritten smem show up
                blk.xchg.assign((size_t)nwarps * 32, 0);
                blk.warp_bar.resize(nwarps);
                pthread_barrier_init(&blk.bar, nullptr, nthreads);
                for (unsigned w = 0; w < nwarps; ++w) {
                    unsigned cnt = std::min(32u, nthreads - w * 32);
                    pthread_barrier_init(&blk.warp_bar[w], nullptr, cnt);
                }
                std::vector<std::thread> ts;
                ts.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t)
                    ts.emplace_back([&, t] {
                        g_blk = &blk;
                        g_tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        g_bid = dim3(bx, by, bz);
                        body();
                    });
                for (auto &th : ts) th.join();
                pthread_barrier_destroy(&blk.bar);
                for (auto &wb : blk.warp_bar) pthread_barrier_destroy(&wb);
            }
}
}  // namespace qd_emu

#define threadIdx (qd_emu::g_tid)
#define blockIdx (qd_emu::g_bid)
#define blockDim (qd_emu::g_blk->block)
#define gridDim (qd_emu::g_blk->grid)

static inline void __syncthreads() { pthread_barrier_wait(&qd_emu::g_blk->bar); }
static inline int __syncthreads_or(int pred) {
    static int flag[2];
    static thread_local int phase = 0;
    // blocks run one at a time; two alternating slots so a fast thread cannot clobber a slot still being read
    int *f = &flag[phase & 1];
    phase++;
    pthread_barrier_wait(&qd_emu::g_blk->bar);
    if (pred) __atomic_store_n(f, 1, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&qd_emu::g_blk->bar);
    const int r = __atomic_load_n(f, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&qd_emu::g_blk->bar);
    if (qd_emu::g_tid.x == 0) *f = 0;
    return r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) {
    pthread_barrier_wait(&qd_emu::g_blk->warp_bar[qd_emu::warp()]);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return qd_emu::shfl_idx(v, src); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int l = qd_emu::lane();
    T o = qd_emu::shfl_idx(v, l - (int)d < 0 ? l : l - (int)d);
    return o;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    int l = qd_emu::lane();
    return qd_emu::shfl_idx(v, l + (int)d > 31 ? l : l + (int)d);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    return qd_emu::shfl_idx(v, qd_emu::lane() ^ m);
}

static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double fmax_emu(double a, double b) { return a > b ? a : b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdividef(float a, float b) { return a / b; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMax(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
using std::min;
using std::max;
static inline void sincospif(float x, float *s, float *c) {
    *s = (float)std::sin(M_PI * (double)x);
    *c = (float)std::cos(M_PI * (double)x);
}
static inline void sincosf_emu(float x, float *s, float *c) { *s = std::sin(x); *c = std::cos(x); }
static inline float exp10f_emu(float x) { return (float)std::pow(10.0, (double)x); }

// dynamic shared memory of the current block
#define QD_DYN_SMEM(name) unsigned char *name = (unsigned char *)(((uintptr_t)qd_emu::g_blk->smem.data() + 127) & ~(uintptr_t)127)
