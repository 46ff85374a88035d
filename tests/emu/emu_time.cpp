// tests/emu/emu_time.cpp -- TEST INFRASTRUCTURE: CPU single-stepping of qd_time.cuh (limiter, crossover).
//   emu_time limiter   <n> <L> <ceiling> <c> <x.f32> <y.f32>
//   emu_time crossover <n> <delay> <sos_lp(6 doubles).f64> <sos_hp.f64> <x.f32> <low.f32> <high.f32>
#include "cuda_emu.h"

namespace qd_emu {
thread_local Block *g_blk = nullptr;
thread_local dim3 g_tid, g_bid;
}  // namespace qd_emu

#include "../../quantumdistortion_b200/csrc/qd_time.cuh"
#include "../../quantumdistortion_b200/csrc/qd_host_time.hpp"

#include <fstream>
#include <iostream>
#include <string>

template <class T>
static std::vector<T> read_all(const char *path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { std::cerr << "cannot open " << path << "\n"; std::exit(2); }
    size_t bytes = (size_t)f.tellg();
    f.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}
template <class T>
static void write_all(const char *path, const std::vector<T> &v) {
    std::ofstream(path, std::ios::binary).write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char **argv) {
    std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "limiter" && argc == 8) {
        const long long n = std::atoll(argv[2]);
        auto x = read_all<float>(argv[6]);
        std::vector<float> y(n, -777.0f);
        qd::LimiterArgs a{};
        a.x = x.data(); a.y = y.data(); a.n = n;
        a.limiter_on = 1; a.lookahead = std::atoi(argv[3]); a.ceiling = std::atof(argv[4]); a.c = std::atof(argv[5]);
        a.apply_mix = 0;
        qd_emu::launch(dim3(1), dim3(qd::QD_TT), qd_host::limiter_smem_bytes(a.lookahead), [&] { qd::limiter_mix_kernel(a); });
        write_all(argv[7], y);
        return 0;
    }
    if (mode == "crossover" && (argc == 9 || argc == 11)) {
        const long long n = std::atoll(argv[2]);
        auto lp = read_all<double>(argv[4]);
        auto hp = read_all<double>(argv[5]);
        auto x = read_all<float>(argv[6]);
        std::vector<float> lo(n, -777.0f), hi(n, -777.0f);
        qd::CrossoverArgs a{};
        a.x = x.data(); a.low = lo.data(); a.high = hi.data(); a.n = n;
        a.low_delay = std::atoi(argv[3]);
        qd_host::fill_crossover(a, lp.data(), hp.data());
        if (argc > 9) { a.tile = std::atoi(argv[9]); a.halo = std::atoi(argv[10]); }   // small tiles: many tiles per clip
        const long long n_tiles = (n + a.tile - 1) / a.tile;
        const unsigned gx = (unsigned)((n_tiles + 32 * qd::QD_XO_WARPS - 1) / (32 * qd::QD_XO_WARPS));
        qd_emu::launch(dim3(gx, 1, 1), dim3(32 * qd::QD_XO_WARPS), 0, [&] { qd::crossover_kernel(a); });
        write_all(argv[7], lo);
        write_all(argv[8], hi);
        return 0;
    }
    std::cerr << "usage: see source\n";
    return 2;
}
