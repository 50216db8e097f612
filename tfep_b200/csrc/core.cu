// Status reporting and device checks shared by every entry point of the C ABI.
#include "common.cuh"

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
