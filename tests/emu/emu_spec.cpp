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
#include "../../quantumdistortion_b200/csrc/qd_spec_team.cuh"

#include <cstdlib>
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

template <class T, int NC, int NW, bool TS = false, bool FX = false, int NG = 1>
static void run(qd::SpecArgsT<T> a, int n_tiles, size_t smem) {
    qd_emu::launch(dim3(n_tiles, (a.batch + NG - 1) / NG, 1), dim3(32 * NW * NG, 1, 1), smem,
                   [&] { qd::spec_pass_kernel<T, NC, NW, TS, FX, NG>(a); });
}

struct Cli {
    int n_fft, nw, n, tile_blocks, quant, smoothing, epilogue, fx_mode = 0, fx_pass = 0, batch = 1;
    double snap, smear, fx_a = 0, fx_b = 0, fx_c = 0;
    float fold, bias, tg, tn;
    std::vector<float> x;
    std::vector<int32_t> tb;
    std::vector<uint8_t> mask;
    std::vector<unsigned char> fx_table;
    std::vector<float> y, tap;
};

template <class T>
static int go(Cli &c) {
    using T2 = qd::V2<T>;
    constexpr bool is_double = sizeof(T) == 8;
    qd_host::SpecTablesT<T> st;
    if (!qd_host::build_spec_tables<T>(c.n_fft, &st)) { std::cerr << "unsupported n_fft\n"; return 2; }
    double kw[5], ks = 0;  // scipy-compatible 5-tap kernel (dsp/quantizer.py:460-465)
    for (int i = 0; i < 5; ++i) { kw[i] = std::exp(-0.5 * (i - 2) * (i - 2)); ks += kw[i]; }
    for (int i = 0; i < 5; ++i) kw[i] /= ks;
    qd_tables ht{};
    ht.n_bins = c.n_fft / 2 + 1;
    ht.target_bins = c.tb.data();
    ht.active_mask = c.mask.data();
    ht.snap = c.snap;
    ht.smear = c.smear;
    ht.smear_radius = 2;
    ht.smear_w = kw;
    qd_host::QuantTablesH qt;
    std::string err;
    if (!qd_host::build_quant_tables(ht, &qt, &err, is_double)) { std::cerr << err << "\n"; return 2; }
    const int n = c.n;
    c.y.assign((size_t)n * c.batch, -777.0f);
    c.tap.assign((size_t)n * c.batch, -777.0f);
    qd::SpecArgsT<T> a{};
    a.x = c.x.data();
    a.y = c.y.data();
    a.tap = c.tap.data();
    a.n = n;
    a.batch = c.batch;
    a.n_frames = 1 + n / st.hop;
    a.tile_blocks = c.tile_blocks;
    a.quant = c.quant;
    a.epilogue = c.epilogue;
    a.fold = c.fold; a.bias = c.bias; a.tube_gain = c.tg; a.tube_norm = c.tn;
    a.wtab = reinterpret_cast<const T2 *>(st.wtab.data());
    a.tw1 = reinterpret_cast<const T2 *>(st.tw1.data());
    a.tw2 = reinterpret_cast<const T2 *>(st.tw2.data());
    a.wsplit = reinterpret_cast<const T2 *>(st.wsplit.data());
    a.invw = st.invw.data();
    a.q.n_slots = qt.n_slots;
    a.q.n_src = (int)qt.src_tab.size();
    a.q.row_limit = qt.row_limit;
    a.q.src_tab = qt.src_tab.data();
    a.q.row_active = qt.row_active.data();
    a.q.slot_of_bin = qt.slot_of_bin.data();
    a.q.slot_invk = qt.slot_invk.data();
    a.q.slot_base = qt.slot_base.data();
    for (int e = 0; e < 5; ++e) a.q.tap[e] = qt.tap[e];
    a.q.keep_active = qt.keep_active;
    a.q.smoothing = c.smoothing;
    a.fx.mode = c.fx_mode;
    a.fx.a = (float)c.fx_a; a.fx.b = (float)c.fx_b; a.fx.c = (float)c.fx_c; a.fx.step = c.fx_a;
    a.fx.table = c.fx_table.empty() ? nullptr : c.fx_table.data();
    a.fx.table_frames = a.n_frames; a.fx.table_per_clip = 0; a.fx.pass = c.fx_pass; a.fx.clip_offset = 0;
    // formant shift (dsp/spectral_fx.py:116-195): QD_EMU_FORMANT_RATIO=<2^(st/12)> in the environment; same tables as
    // qd_plan_create builds
    std::vector<int16_t> fidx;
    std::vector<float> ffrac;
    bool formant = false;
    if (const char *fr = std::getenv("QD_EMU_FORMANT_RATIO")) {
        const double ratio = std::atof(fr);
        if (ratio > 0.0) {
            const int nb = c.n_fft / 2 + 1;
            fidx.resize(nb); ffrac.resize(nb);
            for (int k = 0; k < nb; ++k) {
                const double xk = (double)k / ratio;
                if (xk >= (double)(nb - 1)) { fidx[k] = (int16_t)(nb - 1); ffrac[k] = 0.0f; }
                else { const double fl = std::floor(xk); fidx[k] = (int16_t)fl; ffrac[k] = (float)(xk - fl); }
            }
            a.formant_idx = fidx.data();
            a.formant_frac = ffrac.data();
            a.formant_order = 30;
            formant = true;
        }
    }
    const int blocks_total = (n + st.hop - 1) / st.hop;
    const int n_tiles = (blocks_total + c.tile_blocks - 1) / c.tile_blocks;
    const int nc = c.n_fft / 2, nw = c.nw;
    const int fx_mode = c.fx_mode;
    // team kernel (qd_spec_team.cuh): <nw> = 100 * frames per batch + warps per frame
    if (c.nw >= 100) {
        const int nf = c.nw / 100, cw = c.nw % 100;
        std::vector<uint32_t> ttab;
        qd::TeamGather tg{};
        qd_host::build_team_gather(qt, cw, &ttab, tg.begin);
        tg.src_tab = ttab.data();
#define QD_TEAMCASE(NC_, NF_, CW_)                                                                                  \
        if (nc == NC_ && nf == NF_ && cw == CW_) {                                                                  \
            qd_emu::launch(dim3(n_tiles, a.batch, 1), dim3(32 * NF_ * CW_, 1, 1), qd::SpecSmem<T, NC_, NF_>::bytes(qt.n_slots), \
                           [&] { qd::spec_pass_team_kernel<T, NC_, NF_, CW_>(a, tg); });                               \
            return 0;                                                                                                \
        }
        if constexpr (is_double) { QD_TEAMCASE(1024, 8, 2) }   // 16 x 8 x 8; the float32 plan of this size is 32 x 32 (two passes)
        QD_TEAMCASE(256, 8, 1) QD_TEAMCASE(512, 8, 1) QD_TEAMCASE(2048, 7, 4) QD_TEAMCASE(2048, 4, 4) QD_TEAMCASE(2048, 2, 2) QD_TEAMCASE(4096, 2, 8) QD_TEAMCASE(4096, 2, 4) QD_TEAMCASE(4096, 4, 4)
        std::cerr << "no team instantiation for nc=" << nc << " nf=" << nf << " cw=" << cw << "\n";
        return 2;
    }
#define QD_FXCASE(NC_, NW_)                                                                                       \
    if ((fx_mode || formant) && nc == NC_ && nw == NW_) {                                                         \
        run<T, NC_, NW_, false, true>(a, n_tiles, qd::SpecSmem<T, NC_, NW_>::bytes(qt.n_slots, false, 0, 0, true, formant)); \
        return 0;                                                                                                 \
    }
    QD_FXCASE(1024, 8) QD_FXCASE(1024, 12) QD_FXCASE(256, 4) QD_FXCASE(512, 4) QD_FXCASE(2048, 4)
#define QD_CASE(NC_, NW_)                                                                 \
    if (nc == NC_ && nw == NW_) {                                                         \
        run<T, NC_, NW_>(a, n_tiles, qd::SpecSmem<T, NC_, NW_>::bytes(qt.n_slots));       \
        return 0;                                                                         \
    }
    QD_CASE(256, 4) QD_CASE(512, 4) QD_CASE(1024, 4) QD_CASE(1024, 8) QD_CASE(2048, 4) QD_CASE(4096, 4) QD_CASE(4096, 2)
    if constexpr (!is_double) {
        if (nc == 1024 && nw == 16) {  // the shared-memory-table variant used for n_fft 2048 on the GPU
            run<T, 1024, 8, true, false, 2>(a, n_tiles, qd::SpecSmem<T, 1024, 8, 2>::bytes(qt.n_slots, true, a.q.n_src, 0));
            return 0;
        }
    }
    std::cerr << "no instantiation for nc=" << nc << " nw=" << nw << "\n";
    return 2;
}

int main(int argc, char **argv) {
    // emu_spec <f32|f64> <n_fft> <nw> <n> <tile_blocks> <quant> <smoothing> <snap> <smear> <epilogue> <fold> <bias>
    //          <tube_gain> <tube_norm> <x.f32> <target_bins.i32> <mask.u8> <y_out.f32> <tap_out.f32>
    //          [<fx_mode> <fx_a> <fx_b> <fx_c> <fx_table file or -> <fx_pass>]
    if (argc != 20 && argc != 26) { std::cerr << "usage: see source\n"; return 2; }
    int ai = 1;
    const std::string prec = argv[ai++];
    Cli c;
    c.n_fft = std::atoi(argv[ai++]);
    c.nw = std::atoi(argv[ai++]);
    c.n = std::atoi(argv[ai++]);
    c.tile_blocks = std::atoi(argv[ai++]);
    c.quant = std::atoi(argv[ai++]);
    c.smoothing = std::atoi(argv[ai++]);
    c.snap = std::atof(argv[ai++]);
    c.smear = std::atof(argv[ai++]);
    c.epilogue = std::atoi(argv[ai++]);
    c.fold = (float)std::atof(argv[ai++]);
    c.bias = (float)std::atof(argv[ai++]);
    c.tg = (float)std::atof(argv[ai++]);
    c.tn = (float)std::atof(argv[ai++]);
    c.x = read_all<float>(argv[ai++]);
    c.tb = read_all<int32_t>(argv[ai++]);
    c.mask = read_all<uint8_t>(argv[ai++]);
    const char *y_path = argv[ai++];
    const char *tap_path = argv[ai++];
    if (c.n <= 0 || c.x.size() % (size_t)c.n != 0) { std::cerr << "x size\n"; return 2; }
    c.batch = (int)(c.x.size() / (size_t)c.n);   // several clips of n samples each
    if (argc == 26) {
        c.fx_mode = std::atoi(argv[ai++]);
        c.fx_a = std::atof(argv[ai++]); c.fx_b = std::atof(argv[ai++]); c.fx_c = std::atof(argv[ai++]);
        const char *tp = argv[ai++];
        if (std::string(tp) != "-") c.fx_table = read_all<unsigned char>(tp);
        c.fx_pass = std::atoi(argv[ai++]);
    }
    const int rc = prec == "f64" ? go<double>(c) : go<float>(c);
    if (rc) return rc;
    std::ofstream(y_path, std::ios::binary).write(reinterpret_cast<const char *>(c.y.data()), (std::streamsize)(c.y.size() * sizeof(float)));
    std::ofstream(tap_path, std::ios::binary).write(reinterpret_cast<const char *>(c.tap.data()), (std::streamsize)(c.tap.size() * sizeof(float)));
    return 0;
}
