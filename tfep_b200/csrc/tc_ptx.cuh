// Inline-PTX wrappers shared by the tensor-core kernels (sm_100a): mbarriers with a bounded wait, bulk
// (TMA-engine) copies, tensor-memory allocation / loads / stores, tcgen05.mma with the A operand in tensor
// memory, and the fast-math intrinsics of the log2-domain epilogues.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace tfepb {
namespace tc {

constexpr int TILE_M = 128;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
constexpr uint64_t WATCHDOG_CYCLES = 4000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Up to 256 polls in a four-instruction loop (try_wait itself suspends the warp for a hardware-defined time): the
// bookkeeping of the watchdog runs once per burst, not once per poll (ncu, r02: the polling loops of the waiting warps
// were a fifth of all issued instructions of the epilogue-bound tc_gemm kernels).
__device__ __forceinline__ bool mbar_poll_burst(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "POLL_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 q, n, 256;\n\t"
        "@q bra POLL_%=;\n\t"
        "DONE_%=:\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap (error return), never in a hung GPU.
// (Kernels that re-balance registers between warpgroups with setmaxnreg define TFEPB_INLINE_SLOW_WAIT: ptxas
// cannot allocate registers across an ABI call inside a setmaxnreg region.)
#ifdef TFEPB_INLINE_SLOW_WAIT
static __device__ __forceinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int* error, int tag) {
    uint32_t bursts = 0;
    long long t0 = 0;
    const uint32_t addr = smem_u32(bar);
    while (!mbar_poll_burst(addr, parity)) {
        if ((++bursts & 0xfu) != 0) continue;
        if (t0 == 0) t0 = clock64();
        if ((uint64_t)(clock64() - t0) > WATCHDOG_CYCLES) {
            if (error) atomicExch(error, tag);
            __threadfence_system();
            __trap();
        }
    }
}
#else
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, int* error, int tag) {
    uint32_t bursts = 0;
    long long t0 = 0;
    const uint32_t addr = smem_u32(bar);
    while (!mbar_poll_burst(addr, parity)) {
        if ((++bursts & 0xfu) != 0) continue;           // look at the clock every 4096 failed polls only
        if (t0 == 0) t0 = clock64();
        if ((uint64_t)(clock64() - t0) > WATCHDOG_CYCLES) {
            if (error) atomicExch(error, tag);
            __threadfence_system();
            __trap();
        }
    }
}
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    mbar_wait_slow(bar, parity, error, tag);
}
// Variants on a 32-bit shared-memory address held in a register (hot loops: no generic -> shared conversion per use).
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity, int* error, int tag) {
    if (mbar_try_wait_s(bar, parity)) return;
    uint32_t bursts = 0;
    long long t0 = 0;
    while (!mbar_poll_burst(bar, parity)) {
        if ((++bursts & 0xfu) != 0) continue;
        if (t0 == 0) t0 = clock64();
        if ((uint64_t)(clock64() - t0) > WATCHDOG_CYCLES) {
            if (error) atomicExch(error, tag);
            __threadfence_system();
            __trap();
        }
    }
}
// A value the compiler must keep in a register instead of re-deriving it from threadIdx / kernel parameters.
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void lds_v4(uint32_t addr, uint32_t (&v)[4]) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The same copy delivered to the same shared-memory offset (and signalling the mbarrier at the same offset) of every CTA of
// the cluster named in `mask`.
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// TMEM -> registers (this thread's lane, consecutive columns).  The values are only defined after the
// matching tmem_wait*(), which takes the registers as read-write operands so that no use can be
// scheduled above it.
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(addr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
}
__device__ __forceinline__ void tmem_ld1(uint32_t addr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(addr));
}
__device__ __forceinline__ void tmem_wait1(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]) :: "memory");
}
__device__ __forceinline__ void tmem_wait8(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :: "memory");
}
// registers -> TMEM: 8 consecutive columns of this thread's lane
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t addr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "r"(r) : "memory");
}
__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] . B[smem]^T : A = 128 rows x 16 bf16 (8 TMEM columns, two k-values per column),
// B = N rows x 16 bf16, K-major no-swizzle core matrices in shared memory.
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// kind::f16, A = B = bf16, D = fp32, both operands K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
// One lane of a converged warp; everything the elected lane consumes is computed warp-uniformly outside,
// so the compiler keeps descriptors in uniform registers instead of emitting a per-lane election loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred) :: "memory");
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ... arriving on the mbarrier at the same offset in every CTA of `mask` (a stage both CTAs of a pair read may only be
// refilled when both have consumed it)
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float lg2(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rcp(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// Packed fp32 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2, two IEEE round-to-nearest operations per issue slot on an aligned
// register pair; the results are those of the scalar instructions bit for bit).  The epilogues are bound by issue slots, not
// by the FMA pipe, so every pair of identical operations on two independent values is worth packing.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// ELU in the log2 domain: t = log2(e) h  ->  log2(e) ELU(h) = t > 0 ? t : log2(e) (2^t - 1)
// Branch-free: g(t) = log2(e) (2^min(t, 0) - 1) is 0 for t >= 0 and >= t for t <= 0, so the result is max(t, g).
__device__ __forceinline__ float elu_l2(float t, float l2e) { return fmaxf(t, fmaf(ex2(fminf(t, 0.f)), l2e, -LOG2E)); }
// two of them with ONE packed fma (l2e2 = (l2e, l2e), ml2e2 = (-log2(e), -log2(e))); same values as elu_l2
__device__ __forceinline__ void elu_l2x2(float t0, float t1, f32x2 l2e2, f32x2 ml2e2, float& r0, float& r1) {
    float g0, g1;
    unpack2(fma2(pack2(ex2(fminf(t0, 0.f)), ex2(fminf(t1, 0.f))), l2e2, ml2e2), g0, g1);
    r0 = fmaxf(t0, g0);
    r1 = fmaxf(t1, g1);
}
// softplus of a log2-domain argument z = log2(e) * v: log(1 + e^v) = ln2 * lg2(1 + 2^z)
__device__ __forceinline__ float softplus_l2(float z) { return z > 28.85f ? z * LN2 : LN2 * lg2(1.f + ex2(z)); }


}  // namespace tc
}  // namespace tfepb
