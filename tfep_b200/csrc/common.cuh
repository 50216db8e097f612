// Shared host/device helpers for the tfep_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/tfep_b200.h"
#include "hd_math.cuh"

namespace tfepb {

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local last error, C ABI returns int status)
// ---------------------------------------------------------------------------------------------
char* last_error_buffer();
int fail(int code, const char* fmt, ...);

#define TFEPB_CHECK_ARG(cond, ...)                                     \
    do {                                                               \
        if (!(cond)) return ::tfepb::fail(-1, __VA_ARGS__);            \
    } while (0)

#define TFEPB_CUDA(call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return ::tfepb::fail((int)e__, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return 0;
}

// Refuse to run anywhere but Blackwell datacenter parts: there is no fallback path.
int require_sm100();
int sm_count();
// SMs the current context may use (green contexts / SM partitions give fewer than the device has)
int sm_available();
// min(wanted, CTAs of `kernel` that can be co-resident in the current context)
int coresident_grid(const void* kernel, int threads, size_t smem, int wanted, int* grid);
// cooperative launch of a kernel taking ONE by-value parameter struct: co-residency of the grid is enforced by the driver
int launch_cooperative(const void* kernel, int grid, int threads, size_t smem, void* param, cudaStream_t stream);
// raise the dynamic shared memory limit of `kernel` on the current device to at least `bytes` (cached)
int ensure_dynamic_smem(const void* kernel, size_t bytes);

inline cudaStream_t as_stream(tfepb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// NVTX range around every entry point of the C ABI (header-only NVTX 3: free unless a profiler is attached), so that
// nsys / ncu timelines show which call of the boundary enqueued which kernels.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define TFEPB_NVTX() ::tfepb::NvtxRange nvtx_range__(__func__)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}

}  // namespace tfepb
