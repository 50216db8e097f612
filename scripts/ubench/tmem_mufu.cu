// Micro-benchmark (development aid): do tcgen05.ld (TMEM -> registers), MUFU and FMA work overlap on one SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_mufu tmem_mufu.cu && ./tmem_mufu
// Every CTA allocates all 512 TMEM columns and runs 16 warps (4 per sub-partition = 4 per TMEM lane quadrant).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld16(uint32_t addr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ void ld32(uint32_t addr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
}
__device__ __forceinline__ void ld16x256(uint32_t addr, uint32_t* r) {     // 16 lanes x 256 bit, x4: 16 registers
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ void waitld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float v) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

// mode bits: 1 = TMEM loads (x16), 2 = 16 MUFU per iteration, 4 = 64 FFMA per iteration, 8 = loads only in warps 0-7 and
// math only in warps 8-15, 16 = use .x32 loads (2 KB -> 4 KB per instruction), 32 = 16x256b shape
template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const bool do_ld = (MODE & 1) && (!(MODE & 8) || warp < 8);
    const bool do_math = !(MODE & 8) || warp >= 8;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.001f * (threadIdx.x + i);
    float f0 = 1.0001f, f1 = 0.5f;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (do_ld) {
            const uint32_t a = lane_addr + ((it * 32) & 255);
            if (MODE & 32) ld16x256(a, r); else if (MODE & 16) ld32(a, r); else ld16(a, r);
        }
        if ((MODE & 2) && do_math) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = ex2(acc[i]);
        }
        if ((MODE & 4) && do_math) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], f0, f1);
        }
        if (do_ld) {
            waitld();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i] & 1u);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

template <int MODE>
void run(const char* what, int iters) {
    long long* d; float* sink;
    cudaMalloc(&d, 148 * sizeof(long long)); cudaMalloc(&sink, 4);
    bench<MODE><<<148, 512>>>(iters, d, sink);
    bench<MODE><<<148, 512>>>(iters, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / iters;
    const int ldw = (MODE & 1) ? ((MODE & 8) ? 8 : 16) : 0;
    const double bytes = ldw * 32.0 * ((MODE & 16) ? 32 : 16) * 4;
    printf("%-58s %8.1f cycles/iter", what, cyc);
    if (ldw) printf("  TMEM read %6.1f B/clk/SM", bytes / cyc);
    if (MODE & 2) printf("  MUFU %5.2f /clk/SM", ((MODE & 8) ? 8 : 16) * 32.0 * 16 / cyc);
    if (MODE & 4) printf("  FFMA %6.1f /clk/SM", ((MODE & 8) ? 8 : 16) * 32.0 * 64 / cyc);
    printf("  (%s)\n", cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}

int main() {
    const int it = 2000;
    run<1>("16 warps: LDTM.x16 + wait", it);
    run<17>("16 warps: LDTM.x32 + wait", it);
    run<33>("16 warps: LDTM 16x256b.x4 + wait", it);
    run<2>("16 warps: 16 MUFU", it);
    run<4>("16 warps: 64 FFMA", it);
    run<3>("16 warps: LDTM.x16 in flight over 16 MUFU, then wait", it);
    run<5>("16 warps: LDTM.x16 in flight over 64 FFMA, then wait", it);
    run<7>("16 warps: LDTM.x16 over 16 MUFU + 64 FFMA", it);
    run<6>("16 warps: 16 MUFU + 64 FFMA", it);
    run<9>("8 warps LDTM.x16 only", it);
    run<11>("8 warps LDTM.x16 | 8 warps 16 MUFU", it);
    run<13>("8 warps LDTM.x16 | 8 warps 64 FFMA", it);
    run<10>("8 warps 16 MUFU only", it);
    run<12>("8 warps 64 FFMA only", it);
    return 0;
}
