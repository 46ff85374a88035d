// qd_spec_launch.inl -- definition of qd_launch::launch_spec_t; included by the qd_k_spec_*.cu units only.
#include <algorithm>
#include <atomic>

#include "qd_err.hpp"
#include "qd_spec_launch.hpp"

namespace qd_launch {

template <class T, int NC, int NW, bool TS, bool FX, int NG, bool SA, bool EF>
int launch_spec_t(const qd::SpecArgsT<T> &a, int tiles, int64_t batch, cudaStream_t st) {
    static std::atomic<uint64_t> attr_mask{0};  // devices this instantiation was opted in on
    auto kern = qd::spec_pass_kernel<T, NC, NW, TS, FX, NG, SA, EF>;
    if (int rc_ = qd_err::ensure_dyn_smem(kern, attr_mask, 227 * 1024)) return rc_;
    const size_t smem = qd::SpecSmem<T, NC, NW, NG, SA>::bytes(a.q.n_slots, TS, a.q.n_src, 0, FX, FX && a.formant_idx != nullptr);
    if (smem > 227 * 1024) return qd_err::fail(QD_ERR_UNSUPPORTED, "shared memory need of this kernel variant exceeds 227 KB");
    for (int64_t b0 = 0; b0 < batch; b0 += 65535 * NG) {  // gridDim.y limit
        const int64_t nb = std::min<int64_t>(65535 * NG, batch - b0);
        qd::SpecArgsT<T> c = a;
        c.batch = (int)nb;
        c.x = a.x + (size_t)b0 * a.n;
        c.y = a.y + (size_t)b0 * a.n;
        if (a.tap) c.tap = a.tap + (size_t)b0 * a.n;
        if (a.clip_peak) c.clip_peak = a.clip_peak + b0;
        if (a.frozen) c.frozen = a.frozen + (size_t)b0 * qd::buf_slots<NC>();
        c.fx.clip_offset = a.fx.clip_offset + (int)b0;
        kern<<<dim3((unsigned)tiles, (unsigned)((nb + NG - 1) / NG), 1), 32 * NW * NG, smem, st>>>(c);
    }
    QD_CUDA(cudaGetLastError());
    return QD_OK;
}

}  // namespace qd_launch

#define QD_INSTANTIATE_SPEC_EF(T, NC, NW, TS, FX, NG, SA, EF) \
    template int qd_launch::launch_spec_t<T, NC, NW, TS, FX, NG, SA, EF>(const qd::SpecArgsT<T> &, int, int64_t, cudaStream_t);
#define QD_INSTANTIATE_SPEC(T, NC, NW, TS, FX, NG, SA) QD_INSTANTIATE_SPEC_EF(T, NC, NW, TS, FX, NG, SA, false)
