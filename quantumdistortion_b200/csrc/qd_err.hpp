// qd_err.hpp -- error plumbing shared by the translation units of libqd_b200.so: the thread-local message behind
// qd_last_error(), the status helpers, and the per-device opt-in for large dynamic shared memory.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>

#include "../../include/qd_b200.h"

namespace qd_err {

inline thread_local std::string g_err;   // one instance per thread across the whole library (C++17 inline variable)

inline int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

}  // namespace qd_err

#define QD_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return qd_err::fail(QD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

namespace qd_err {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember the devices a kernel was opted in on
// (a process may render on several GPUs one after the other)
template <class K>
int ensure_dyn_smem(K kern, std::atomic<uint64_t> &mask, int bytes) {
    int dev = 0;
    QD_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(mask.load(std::memory_order_acquire) & bit)) {
        QD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        mask.fetch_or(bit, std::memory_order_release);
    }
    return QD_OK;
}

}  // namespace qd_err
