// tests/emu/emu_yin.cpp -- TEST INFRASTRUCTURE: runs at_yin_diff_kernel (quantumdistortion_b200/csrc/qd_yin.cuh) on the CPU
// through cuda_emu.h.  usage: emu_yin n frame_size hop max_tau cols x.f32 diff.f64
#include "cuda_emu.h"

namespace qd_emu {
thread_local Block *g_blk = nullptr;
thread_local dim3 g_tid, g_bid;
}  // namespace qd_emu

#include "../../quantumdistortion_b200/csrc/qd_yin.cuh"

#include <fstream>
#include <iostream>

int main(int argc, char **argv) {
    if (argc != 8) { std::cerr << "usage: see source\n"; return 2; }
    const long long n = std::atoll(argv[1]);
    qd::AtYinArgs a{};
    a.n = n; a.frame_size = std::atoi(argv[2]); a.hop = std::atoi(argv[3]); a.max_tau = std::atoi(argv[4]);
    constexpr int L = qd::AT_YL;
    const int total_cols = (a.max_tau + L - 1) / L;
    a.lag_threads = std::min(std::min(std::atoi(argv[5]), total_cols), (int)qd::AT_YC);
    const int tblocks = (total_cols + a.lag_threads - 1) / a.lag_threads;
    a.frames = (int)((n + a.hop - 1) / a.hop);
    a.stride = (a.max_tau + 2) & ~1;
    std::vector<float> x((size_t)n);
    std::ifstream(argv[6], std::ios::binary).read(reinterpret_cast<char *>(x.data()), n * sizeof(float));
    std::vector<double> diff((size_t)a.frames * a.stride, -1.0);
    a.det = x.data(); a.diff = diff.data();
    const int threads = qd::AT_YT;
    const size_t smem = qd::at_yin_smem_doubles(a.hop, (tblocks - 1) * a.lag_threads, a.lag_threads) * sizeof(double) + 128;
    qd_emu::launch(dim3(tblocks, 1, 1), dim3(threads), smem, [&] { qd::at_yin_diff_kernel(a); });
    std::ofstream(argv[7], std::ios::binary).write(reinterpret_cast<const char *>(diff.data()), diff.size() * sizeof(double));
    return 0;
}
