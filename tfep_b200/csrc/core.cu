// Status reporting and device checks shared by every entry point of the C ABI.
#include "common.cuh"

#include <cuda.h>
#include <string.h>

namespace tfepb {

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

namespace {
struct DeviceCache {
    int device = -1, sms = 0, major = 0, minor = 0;
};
int query(DeviceCache* out) {
    static thread_local DeviceCache cache;
    int dev = -1;
    TFEPB_CUDA(cudaGetDevice(&dev));
    if (cache.device != dev) {
        cudaDeviceProp prop;
        TFEPB_CUDA(cudaGetDeviceProperties(&prop, dev));
        cache.device = dev;
        cache.sms = prop.multiProcessorCount;
        cache.major = prop.major;
        cache.minor = prop.minor;
    }
    *out = cache;
    return 0;
}
}  // namespace

int require_sm100() {
    DeviceCache c;
    if (int rc = query(&c)) return rc;
    if (c.major != 10)
        return fail(-2, "tfep_b200 is built for sm_100a only; device has compute capability %d.%d "
                        "(there is no fallback path)", c.major, c.minor);
    return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: remember, per (device, kernel), the
// largest size configured so far (a process may drive several GPUs from one thread).
int ensure_dynamic_smem(const void* kernel, size_t bytes) {
    struct Entry { int device; const void* kernel; size_t bytes; };
    static thread_local Entry table[64];
    static thread_local int used = 0;
    int dev = -1;
    TFEPB_CUDA(cudaGetDevice(&dev));
    Entry* e = nullptr;
    for (int i = 0; i < used; ++i)
        if (table[i].device == dev && table[i].kernel == kernel) e = &table[i];
    if (e != nullptr && e->bytes >= bytes) return 0;
    TFEPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (e == nullptr && used < 64) e = &table[used++];
    if (e != nullptr) *e = Entry{dev, kernel, bytes};
    return 0;
}

// SMs the CURRENT context may use: the device's count unless the context is a green context / SM partition
// (cuCtxGetDevResource, CUDA >= 12.4).  Persistent kernels whose CTAs wait for each other size their grid with this.
int sm_available() {
    int n = sm_count();
    // driver entry points are fetched through the runtime (the library must load on machines without a driver)
    typedef CUresult (*GetCurrent)(CUcontext*);
    typedef CUresult (*GetDevResource)(CUcontext, CUdevResource*, CUdevResourceType);
    static thread_local GetCurrent get_current = nullptr;
    static thread_local GetDevResource get_resource = nullptr;
    static thread_local bool looked_up = false;
    if (!looked_up) {
        looked_up = true;
        void* f0 = nullptr; void* f1 = nullptr;
        cudaDriverEntryPointQueryResult q0, q1;
        if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &f0, cudaEnableDefault, &q0) == cudaSuccess && q0 == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuCtxGetDevResource", &f1, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess) {
            get_current = reinterpret_cast<GetCurrent>(f0);
            get_resource = reinterpret_cast<GetDevResource>(f1);
        }
        cudaGetLastError();
    }
    CUcontext ctx = nullptr;
    if (get_current != nullptr && get_resource != nullptr && get_current(&ctx) == CUDA_SUCCESS && ctx != nullptr) {
        CUdevResource res;
        memset(&res, 0, sizeof(res));
        if (get_resource(ctx, &res, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS && res.sm.smCount > 0 && (int)res.sm.smCount < n)
            n = (int)res.sm.smCount;
    }
    return n;
}

// Grid of a persistent kernel whose CTAs spin on flags published by other CTAs of the same launch: at most the number
// of CTAs that can be CO-RESIDENT (occupancy query x SMs of the context), and launched cooperatively, so that the driver
// refuses the launch (an error return) instead of letting it dead-lock when fewer CTAs fit than assumed (MPS limits, ...).
int coresident_grid(const void* kernel, int threads, size_t smem, int wanted, int* grid) {
    int per_sm = 0;
    TFEPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) return fail(-3, "persistent kernel does not fit on an SM (%d threads, %zu bytes of shared memory)", threads, smem);
    const int cap = sm_available() * per_sm;
    *grid = wanted < cap ? wanted : cap;
    return 0;
}

int launch_cooperative(const void* kernel, int grid, int threads, size_t smem, void* param, cudaStream_t stream) {
    void* args[] = {param};
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TFEPB_CUDA(cudaLaunchKernelExC(&cfg, kernel, args));
    return 0;
}

int sm_count() {
    DeviceCache c;
    if (query(&c) != 0) return 148;
    return c.sms;
}

}  // namespace tfepb

extern "C" int tfepb_abi_version(void) { return TFEPB_ABI_VERSION; }

extern "C" const char* tfepb_last_error(void) { return tfepb::last_error_buffer(); }

extern "C" int tfepb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
    int dev = -1;
    TFEPB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    TFEPB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}
