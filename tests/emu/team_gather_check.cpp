// tests/emu/team_gather_check.cpp -- TEST INFRASTRUCTURE (host only): invariants of qd_host::build_team_gather, the per-warp
// gather lists of the team kernel (qd_spec_team.cuh).  For random target tables and every team width:
//   * the lists are 32-aligned and contiguous, padding entries are 0;
//   * every source of the single-warp list (src_tab) appears exactly once, in the same order, with the same slot and bin;
//   * a slot's sources all sit in ONE warp's list (slot sums are race free);
//   * `off` counts the same-slot entries before an entry inside its group of 32, `tail` marks the last one of its slot
//     inside the group -- so a scan over a group followed by tail stores adds every source exactly once.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>

#include "../../quantumdistortion_b200/csrc/qd_host_tables.hpp"

static int fail(const char *what, int cw, int seed) {
    std::fprintf(stderr, "team gather: %s (cw %d, seed %d)\n", what, cw, seed);
    return 1;
}

int main() {
    double kw[5] = {0.05449, 0.24420, 0.40262, 0.24420, 0.05449};
    for (int seed = 0; seed < 40; ++seed) {
        std::mt19937 rng(seed);
        const int n = (seed % 4 == 0) ? 4097 : (seed % 4 == 1) ? 257 : (seed % 4 == 2) ? 2049 : 1025;
        std::vector<int32_t> tb(n);
        std::vector<uint8_t> mask(n);
        // contiguous preimages like the reference's tables: runs of random length (1 .. 120) map to one bin of the run
        for (int i = 0; i < n;) {
            const int len = 1 + (int)(rng() % ((seed & 1) ? 120 : 12));
            const int hi = std::min(n, i + len);
            const int t = i + (int)(rng() % (unsigned)(hi - i));
            for (int j = i; j < hi; ++j) { tb[j] = t; mask[j] = (rng() % 5) != 0; }
            i = hi;
        }
        mask[0] = 0;   // the DC bin never gives energy away (dsp/pipeline.py:164-177), so no real entry encodes as 0
        qd_tables ht{};
        ht.n_bins = n; ht.target_bins = tb.data(); ht.active_mask = mask.data();
        ht.snap = 0.9; ht.smear = 0.2; ht.smear_radius = 2; ht.smear_w = kw;
        qd_host::QuantTablesH qt;
        std::string err;
        if (!qd_host::build_quant_tables(ht, &qt, &err)) return fail(err.c_str(), 0, seed);
        for (int cw : {1, 2, 4, 8}) {
            std::vector<uint32_t> tab;
            int begin[9] = {0};
            qd_host::build_team_gather(qt, cw, &tab, begin);
            if (begin[0] != 0) return fail("begin[0]", cw, seed);
            std::vector<uint32_t> flat;
            std::map<int, int> owner;   // slot -> warp
            for (int w = 0; w < cw; ++w) {
                if (begin[w] % 32 || begin[w + 1] % 32 || begin[w + 1] < begin[w]) return fail("alignment", cw, seed);
                for (int g0 = begin[w]; g0 < begin[w + 1]; g0 += 32) {
                    for (int l = 0; l < 32; ++l) {
                        const uint32_t e = tab[g0 + l];
                        if (e == 0u) {   // padding: only after the warp's last real entry
                            for (int i = g0 + l; i < begin[w + 1]; ++i) if (tab[i] != 0u) return fail("padding inside a list", cw, seed);
                            break;
                        }
                        const int slot = (int)((e >> 13) & 0x1fffu), off = (int)((e >> 26) & 31u), tail = (int)(e >> 31);
                        if (owner.count(slot) && owner[slot] != w) return fail("slot in two lists", cw, seed);
                        owner[slot] = w;
                        int want_off = 0;
                        for (int j = l - 1; j >= 0 && ((tab[g0 + j] >> 13) & 0x1fffu) == (uint32_t)slot && tab[g0 + j] != 0u; --j) ++want_off;
                        if (off != want_off) return fail("off", cw, seed);
                        const bool last = l == 31 || tab[g0 + l + 1] == 0u || ((tab[g0 + l + 1] >> 13) & 0x1fffu) != (uint32_t)slot;
                        if (tail != (int)last) return fail("tail", cw, seed);
                        flat.push_back(e & 0x03ffffffu);   // slot and bin
                    }
                }
            }
            if (begin[cw] != (int)tab.size() && !(tab.size() == 1 && begin[cw] == 0)) return fail("begin[cw]", cw, seed);
            if (flat.size() != qt.src_tab.size()) return fail("source count", cw, seed);
            for (size_t i = 0; i < flat.size(); ++i)
                if (flat[i] != (qt.src_tab[i] & 0x03ffffffu)) return fail("order / content", cw, seed);
            if ((int)owner.size() != qt.n_slots) return fail("slot count", cw, seed);
        }
    }
    std::puts("ok");
    return 0;
}
