// tests/emu/emu_spec.cpp -- TEST INFRASTRUCTURE: runs quantumdistortion_b200/csrc/qd_spec.cuh on the CPU
// through cuda_emu.h (one OS thread per CUDA thread).  Used by tests/test_emu_kernels.py to
// check index math and table layout against the oracle without a GPU.
//
//   emu_spec <n_fft> <nw> <n> <tile_blocks> <quant> <smoothing> <snap> <smear> <epilogue> <fold> <bias>
//            <tube_gain> <tube_norm> <x.f32> <target_bins.i32> <mask.u8> <y_out.f32> <tap_out.f32>
//            [<fx_mode> <fx_a> <fx_b> <fx_c> <fx_table file or -> <fx_pass>]
#include "cuda_emu.h"

namespace qd_emu {
thread_local Block *g_blk = nullptr;
thread_local dim3 g_tid, g_bid;
}  // namespace qd_emu

#include "../../quantumdistortion_b200/csrc/qd_host_tables.hpp"
#include "../../quantumdistortion_b200/csrc/qd_spec.cuh"

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

template <int NC, int NW, bool TS = false, bool FX = false>
static void run(qd::SpecArgs a, int n_tiles, size_t smem) {
    qd_emu::launch(dim3(n_tiles, 1, 1), dim3(32 * NW, 1, 1), smem, [&] { qd::spec_pass_kernel<NC, NW, TS, FX>(a); });
}

int main(int argc, char **argv) {
    if (argc != 19 && argc != 25) { std::cerr << "usage: see source\n"; return 2; }
    int ai = 1;
    const int n_fft = std::atoi(argv[ai++]);
    const int nw = std::atoi(argv[ai++]);
    const int n = std::atoi(argv[ai++]);
    const int tile_blocks = std::atoi(argv[ai++]);
    const int quant = std::atoi(argv[ai++]);
    const int smoothing = std::atoi(argv[ai++]);
    const double snap = std::atof(argv[ai++]);
    const double smear = std::atof(argv[ai++]);
    const int epilogue = std::atoi(argv[ai++]);
    const float fold = (float)std::atof(argv[ai++]);
    const float bias = (float)std::atof(argv[ai++]);
    const float tg = (float)std::atof(argv[ai++]);
    const float tn = (float)std::atof(argv[ai++]);
    auto x = read_all<float>(argv[ai++]);
    auto tb = read_all<int32_t>(argv[ai++]);
    auto mask = read_all<uint8_t>(argv[ai++]);
    const char *y_path = argv[ai++];
    const char *tap_path = argv[ai++];
    if ((int)x.size() != n) { std::cerr << "x size\n"; return 2; }
    int fx_mode = 0, fx_pass = 0;
    double fx_a = 0, fx_b = 0, fx_c = 0;
    std::vector<unsigned char> fx_table;
    if (argc == 25) {
        fx_mode = std::atoi(argv[ai++]);
        fx_a = std::atof(argv[ai++]); fx_b = std::atof(argv[ai++]); fx_c = std::atof(argv[ai++]);
        const char *tp = argv[ai++];
        if (std::string(tp) != "-") fx_table = read_all<unsigned char>(tp);
        fx_pass = std::atoi(argv[ai++]);
    }

    qd_host::SpecTables st;
    if (!qd_host::build_spec_tables(n_fft, &st)) { std::cerr << "unsupported n_fft\n"; return 2; }
    // scipy-compatible 5-tap kernel (dsp/quantizer.py:460-465)
    double kw[5], ks = 0;
    for (int i = 0; i < 5; ++i) { kw[i] = std::exp(-0.5 * (i - 2) * (i - 2)); ks += kw[i]; }
    for (int i = 0; i < 5; ++i) kw[i] /= ks;
    qd_tables ht{};
    ht.n_bins = n_fft / 2 + 1;
    ht.target_bins = tb.data();
    ht.active_mask = mask.data();
    ht.snap = snap;
    ht.smear = smear;
    ht.smear_radius = 2;
    ht.smear_w = kw;
    qd_host::QuantTablesH qt;
    std::string err;
    if (!qd_host::build_quant_tables(ht, &qt, &err)) { std::cerr << err << "\n"; return 2; }

    std::vector<float> y(n, -777.0f), tap(n, -777.0f);
    qd::SpecArgs a{};
    a.x = x.data();
    a.y = y.data();
    a.tap = tap.data();
    a.n = n;
    a.n_frames = 1 + n / st.hop;
    a.tile_blocks = tile_blocks;
    a.quant = quant;
    a.epilogue = epilogue;
    a.fold = fold; a.bias = bias; a.tube_gain = tg; a.tube_norm = tn;
    a.wtab = reinterpret_cast<const float2 *>(st.wtab.data());
    a.tw1 = reinterpret_cast<const float2 *>(st.tw1.data());
    a.tw2 = reinterpret_cast<const float2 *>(st.tw2.data());
    a.wsplit = reinterpret_cast<const float2 *>(st.wsplit.data());
    a.invw = st.invw.data();
    a.q.n_slots = qt.n_slots;
    a.q.n_aff = qt.n_aff;
    a.q.n_src = (int)qt.src_tab.size();
    a.q.row_limit = qt.row_limit;
    a.q.src_tab = qt.src_tab.data();
    a.q.row_active = qt.row_active.data();
    a.q.row_aff = qt.row_aff.data();
    a.q.row_aff_base = qt.row_aff_base.data();
    a.q.aff = reinterpret_cast<const qd::AffEntry *>(qt.aff.data());
    a.q.keep_active = qt.keep_active;
    a.q.smoothing = smoothing;

    a.fx.mode = fx_mode;
    a.fx.a = (float)fx_a; a.fx.b = (float)fx_b; a.fx.c = (float)fx_c; a.fx.step = fx_a;
    a.fx.table = fx_table.empty() ? nullptr : fx_table.data();
    a.fx.table_frames = a.n_frames; a.fx.table_per_clip = 0; a.fx.pass = fx_pass; a.fx.clip_offset = 0;
    const int blocks_total = (n + st.hop - 1) / st.hop;
    const int n_tiles = (blocks_total + tile_blocks - 1) / tile_blocks;
    const int nc = n_fft / 2;
#define QD_FXCASE(NC_, NW_)                                                                          \
    if (fx_mode && nc == NC_ && nw == NW_) {                                                         \
        run<NC_, NW_, false, true>(a, n_tiles, qd::SpecSmem<NC_, NW_>::bytes(qt.n_slots, false, 0, 0, true)); \
        goto done;                                                                                   \
    }
    QD_FXCASE(1024, 8) QD_FXCASE(256, 4) QD_FXCASE(2048, 4)
#define QD_CASE(NC_, NW_)                                                        \
    if (nc == NC_ && nw == NW_) {                                                \
        run<NC_, NW_>(a, n_tiles, qd::SpecSmem<NC_, NW_>::bytes(qt.n_slots));    \
        goto done;                                                               \
    }
    QD_CASE(256, 4) QD_CASE(512, 4) QD_CASE(1024, 4) QD_CASE(1024, 8) QD_CASE(2048, 4) QD_CASE(4096, 4)
    if (nc == 1024 && nw == 16) {  // the shared-memory-table variant used for n_fft 2048 on the GPU
        run<1024, 16, true>(a, n_tiles, qd::SpecSmem<1024, 16>::bytes(qt.n_slots, true, a.q.n_src, a.q.n_aff));
        goto done;
    }
    std::cerr << "no instantiation for nc=" << nc << " nw=" << nw << "\n";
    return 2;
done:
    std::ofstream(y_path, std::ios::binary).write(reinterpret_cast<const char *>(y.data()), (std::streamsize)(n * sizeof(float)));
    std::ofstream(tap_path, std::ios::binary).write(reinterpret_cast<const char *>(tap.data()), (std::streamsize)(n * sizeof(float)));
    std::cout << "slots=" << qt.n_slots << " aff=" << qt.n_aff << " tiles=" << n_tiles << "\n";
    return 0;
}
