// Host-side plumbing shared by the C-ABI translation units: error reporting, device checks,
// launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dynax_b200.h"

namespace dynax {

char *last_error_buf();  // thread-local, defined in rc_api.cu
constexpr int kErrBufLen = 512;

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), kErrBufLen, fmt, ap);
    va_end(ap);
    return code;
}

#define DYNAX_CUDA_TRY(expr)                                                                     \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return ::dynax::fail(DYNAX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                 __FILE__, __LINE__);                                            \
    } while (0)

// NVTX range around every C-ABI entry point: shows up as "dynax_..." in Nsight Systems / ncu
// --nvtx; header-only NVTX3, a no-op unless a tool is attached.
struct ApiRange {
    explicit ApiRange(const char *name) { nvtxRangePushA(name); }
    ~ApiRange() { nvtxRangePop(); }
};
#define DYNAX_API_RANGE() ::dynax::ApiRange dynax_api_range_(__func__)

// There is no CPU fallback: refuse to run unless the current device is Blackwell (cc 10.x).
int require_sm100();

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// (n, m, p) of the dense-LTI kernels that are instantiated (lti_api.cu, lti_chunked.cu)
#define DYNAX_LTI_DIMS(X) \
    X(1, 1, 1) X(2, 1, 1) X(2, 2, 1) X(2, 2, 2) X(3, 1, 1) X(3, 3, 1) X(3, 3, 3) X(3, 5, 1) X(4, 1, 1) X(4, 2, 2) X(4, 3, 4) X(4, 4, 4)

struct DeviceInfo {
    int sm_count;
    int smem_optin;
};
int device_info(DeviceInfo *out);

// ---- launch registry (debug hook; tests/test_gpu_variants.py) -------------------------------------------------
// Every call site that launches a rollout kernel carries DYNAX_NOTE_LAUNCH(): the enclosing (template) function
// registers its signature -- template arguments included -- when the library is loaded, and counts its launches.
// dynax_debug_kernel_count/_tag/_launches enumerate them, so a test can prove that every kernel variant the
// library instantiates is selected (and checked against the oracle) by some test case.
int registry_add(const char *tag);   // defined in rc_api.cu
void registry_note(int id);
template <class Here>
struct LaunchTag {
    static const int id;
};
template <class Here>
const int LaunchTag<Here>::id = registry_add(Here::tag());
#define DYNAX_NOTE_LAUNCH()                                                      \
    do {                                                                         \
        struct Here_ {                                                           \
            static const char *tag() { return __PRETTY_FUNCTION__; }             \
        };                                                                       \
        ::dynax::registry_note(::dynax::LaunchTag<Here_>::id);                   \
    } while (0)

template <class Kernel>
inline int set_smem(Kernel k, size_t bytes) {
    DYNAX_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DYNAX_OK;
}

}  // namespace dynax
