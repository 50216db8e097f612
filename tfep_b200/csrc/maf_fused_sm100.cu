// Fused MAF layer forward for sm_100a: MADE conditioner on tcgen05 tensor cores + neural-spline transformer
// and per-sample log|det J| in the GEMM epilogue.
//
//   y, logdet = T_spline(x ; MADE(x))            (reference: nn/flows/autoregressive.py:144-177,
//                                                  nn/conditioners/made.py:294-329, nn/transformers/spline.py)
//
// One persistent CTA per SM walks over tiles of 128 samples.  For one tile:
//   x tile (fp32)  --bulk copy-->  smem  --convert-->  A0 (bf16, UMMA K-major core-matrix layout)
//   GEMM1  D1 = A0 W1^T   (tcgen05.mma, accumulator in TMEM)   -> epilogue: +b1, ELU, bf16 -> A1 (smem)
//   GEMM2  D2 = A1 W2^T                                         -> epilogue: +b2, ELU, bf16 -> A2 (smem)
//   GEMM3  chunk c: 8 features x 25 spline parameters = 200 (+8 pad) columns, double-buffered in TMEM;
//          epilogue: softmax / softplus / bin search / rational-quadratic map + log-det straight out of TMEM,
//          so the (batch, 1650) parameter tensor never exists in memory.
//   y tile is written in place over the x tile in smem and leaves with one bulk store.
//
// Hidden units are degree-sorted and the output layer is packed feature-major (tfep_b200/_pack.py), so the
// autoregressive masks are block lower-triangular: the host-built schedule only lists the weight blocks
// that contain non-zeros -- masked blocks are neither fetched nor multiplied.  Weight blocks are stored in
// global memory as exact images of their shared-memory layout and streamed by the bulk-copy (TMA) engine
// through an mbarrier ring.
//
// Biases ride in the GEMMs: every A operand carries two constant-one columns and the weight blocks hold
// bf16(b) and bf16(b - bf16(b)) in the matching columns, so the epilogues never load a bias.  The rows of
// the output layer that feed softmax / softplus are pre-multiplied by log2(e), so the epilogue uses the
// hardware ex2 / lg2 directly.
//
// Warp roles: warp 0 = TMEM allocator + bulk-copy producer, warp 1 = MMA issuer, warps 2..17 = epilogue
// (four warpgroups of 128 threads: thread <-> sample row, TMEM lane quadrant = warp % 4; the warpgroups
// split the columns of a hidden layer / the feature slots of a chunk).
#include "common.cuh"

#include <cuda_bf16.h>

namespace tfepb {
namespace fused {

constexpr int TILE_M = 128;
constexpr int EPI_WGS = 4;                      // epilogue warpgroups
constexpr int EPI_THREADS = EPI_WGS * 128;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 256 * 48 * 2;      // one weight block: <= 256 rows x 48 k x bf16
constexpr int SLAB_BYTES = TILE_M * 16;        // 8 k-values of 128 rows
constexpr int FEATS_PER_CHUNK = 8;
constexpr int NPAR = 25;                        // circular spline, K = 8: 8 widths, 8 heights, 8 slopes, shift
constexpr int PSTRIDE = 32;                     // accumulator columns per feature slot (25 used)
constexpr int CHUNK_N = FEATS_PER_CHUNK * PSTRIDE;   // 256
constexpr int ACC1_COL = 256;                   // TMEM column of the second GEMM3 accumulator buffer
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr uint64_t WATCHDOG_CYCLES = 4000000000ull;

struct Op {                 // one weight block = one ring stage
    uint32_t w_off;         // byte offset into the packed weights (multiple of 16)
    uint32_t w_bytes;
    uint16_t n;             // MMA N of this block (rows of the weight block)
    uint16_t tmem_col;      // destination accumulator column
    uint16_t ksteps;        // K = 16 steps in this block
    uint16_t a_slab0;       // first A slab (8 k-values) this block multiplies
    uint32_t flags;
};
enum : uint32_t {
    OP_FIRST = 1u,          // first block of its accumulator: overwrite instead of accumulate
    OP_COMMIT = 2u,         // last block of its accumulator group: commit to acc_full[acc]
    OP_ACC1 = 4u,           // accumulator id 1 (else 0)
    OP_WAIT_A = 16u,        // wait for the A operand (start of a GEMM phase)
    OP_WAIT_EMPTY = 32u,    // wait until the epilogue drained accumulator `acc` (GEMM3 chunks)
};

struct __align__(16) FeatConst {   // per sorted feature
    int col;                // column in x / y, -1 = padding
    float x0, L, invL, Rw, Rh, y0, pad;
};

struct Params {
    const float* x; float* y; float* logdet;
    int batch, D;               // D = row length of x / y
    int K1;                     // D padded to a multiple of 16
    int HP;                     // hidden width padded to a multiple of 16
    int n_chunks;               // GEMM3 chunks
    int n_ops;
    const Op* ops;
    const uint8_t* weights;     // packed bf16 weight blocks
    const FeatConst* feats;     // n_chunks * FEATS_PER_CHUNK
    float min_bin, min_slope, slope_offset2;   // slope_offset2 = log2(e) * log(exp(1 - min_slope) - 1)
    int* error;                 // device int: set on watchdog timeout
    float* debug_params;        // optional (batch, n_chunks * CHUNK_N): conditioner outputs as seen by the epilogue
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if ((uint64_t)(clock64() - t0) > WATCHDOG_CYCLES) {
            if (error) atomicExch(error, tag);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
// TMEM -> registers: 8 consecutive columns of this thread's lane.  The values are only defined after
// tmem_wait(), which takes the registers as read-write operands so that no use can be scheduled above it.
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
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
__device__ __forceinline__ void tmem_wait8(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :: "memory");
}

// K-major, no-swizzle operand descriptor: core matrix = 8 rows x 16 bytes stored contiguously (128 B);
// SBO = distance between 8-row groups, LBO = distance between the two 8-element K halves of one MMA.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;         // descriptor version of sm_100
    return d;                        // base offset 0, layout type 0 = no swizzle
}
// kind::f16, A = B = bf16, D = fp32, both operands K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ float fast_elu(float v) { return v > 0.f ? v : ex2(v * LOG2E) - 1.f; }
// softplus of a log2-domain argument z = log2(e) * v: log(1 + e^v) = ln2 * lg2(1 + 2^z)
__device__ __forceinline__ float softplus_l2(float z) { return z > 28.85f ? z * LN2 : LN2 * lg2(1.f + ex2(z)); }

// ------------------------------------------------------------------------------------------------
// shared memory plan
// ------------------------------------------------------------------------------------------------
constexpr int MAX_OPS = 128;
struct Smem {
    Op ops[MAX_OPS];
    uint64_t w_full[STAGES], w_empty[STAGES];
    uint64_t x_full, x_empty, a_ready;
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

// Spline epilogue for ONE feature of one sample.  r[0..24] = conditioner outputs of the feature (bias
// included by the GEMM); widths r[0..7], heights r[8..15] and slopes r[16..23] arrive pre-multiplied by
// log2(e), the shift r[24] does not.  Circular spline with K = 8 (reference nn/transformers/spline.py:
// 184-241, 319-417, 424-501): after the wrap the far tails cannot be reached, and x exactly on the first
// knot gives the same value through the first bin.
__device__ __forceinline__ float spline8_circular(const uint32_t (&r)[32], float x, const FeatConst& fc, float min_bin,
                                                  float min_slope, float slope_offset2, float& y) {
    float p[NPAR];
#pragma unroll
    for (int i = 0; i < NPAR; ++i) p[i] = __uint_as_float(r[i]);
    // wrap: (x - x0 + shift) mod L, result in [0, L)
    float t = x - fc.x0 + p[24];
    t = t - fc.L * floorf(t * fc.invL);
    t = (t < 0.f) ? t + fc.L : t;
    t = (t >= fc.L) ? t - fc.L : t;
    // softmax numerators (log2 domain)
    float mw = fmaxf(fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3])), fmaxf(fmaxf(p[4], p[5]), fmaxf(p[6], p[7])));
    float mh = fmaxf(fmaxf(fmaxf(p[8], p[9]), fmaxf(p[10], p[11])), fmaxf(fmaxf(p[12], p[13]), fmaxf(p[14], p[15])));
    float ew[8], eh[8], sw = 0.f, sh = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ew[k] = ex2(p[k] - mw); sw += ew[k];
        eh[k] = ex2(p[8 + k] - mh); sh += eh[k];
    }
    const float rw = fc.Rw * rcp(sw), rh = fc.Rh * rcp(sh);
    // walk the knots (relative to x0 / y0): last bin whose left knot is below t
    float left = 0.f, bottom = 0.f;
    float w_sel = fmaf(ew[0], rw, min_bin), h_sel = fmaf(eh[0], rh, min_bin), xk = 0.f, yk = 0.f;
    float raw0 = p[16], raw1 = p[17];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        left += fmaf(ew[k], rw, min_bin);
        bottom += fmaf(eh[k], rh, min_bin);
        const bool adv = t > left;
        w_sel = adv ? fmaf(ew[k + 1], rw, min_bin) : w_sel;
        h_sel = adv ? fmaf(eh[k + 1], rh, min_bin) : h_sel;
        xk = adv ? left : xk;
        yk = adv ? bottom : yk;
        raw0 = adv ? p[16 + k + 1] : raw0;
        raw1 = adv ? p[16 + ((k + 2) & 7)] : raw1;       // knot 8 is tied to knot 0 (circular)
    }
    const float dk = softplus_l2(raw0 + slope_offset2) + min_slope;
    const float dk1 = softplus_l2(raw1 + slope_offset2) + min_slope;
    const float iw = rcp(w_sel);
    const float e = (t - xk) * iw;
    const float s = h_sel * iw;
    const float ome = 1.f - e, u = e * ome, e2 = e * e;
    const float q = dk1 + dk - 2.f * s;
    const float den = fmaf(q, u, s);
    const float iden = rcp(den);
    y = fc.y0 + yk + h_sel * fmaf(s, e2, dk * u) * iden;
    const float nn = fmaf(dk1, e2, fmaf(2.f * s, u, dk * ome * ome));
    const float rr = s * iden;
    return LN2 * lg2(nn * rr * rr);
}

__global__ void __launch_bounds__(THREADS, 1) maf_spline_fwd_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int a_slabs = max(p.K1, p.HP) / 8;
    uint8_t* sA = smem_raw;                                         // A operand: a_slabs x 2048 B
    uint8_t* sW = sA + (size_t)a_slabs * SLAB_BYTES;                // weight ring
    float* sX = reinterpret_cast<float*>(sW + (size_t)STAGES * STAGE_BYTES);   // x / y tile, row-major [128][D]
    const int x_tile_bytes = TILE_M * p.D * 4;
    FeatConst* sFeat = reinterpret_cast<FeatConst*>(reinterpret_cast<uint8_t*>(sX) + ((x_tile_bytes + 127) & ~127));
    const int n_feat = p.n_chunks * FEATS_PER_CHUNK;
    float* sLd = reinterpret_cast<float*>(sFeat + n_feat);          // [EPI_WGS - 1][128] log-det partials
    Smem* sm = reinterpret_cast<Smem*>(sLd + (EPI_WGS - 1) * TILE_M);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (p.batch + TILE_M - 1) / TILE_M;

    // ---- one-time setup ----
    for (int i = tid; i < n_feat; i += THREADS) sFeat[i] = p.feats[i];
    for (int i = tid; i < p.n_ops; i += THREADS) sm->ops[i] = p.ops[i];
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm->w_full[s], 1); mbar_init(&sm->w_empty[s], 1); }
        mbar_init(&sm->x_full, 1);
        mbar_init(&sm->x_empty, 1);
        mbar_init(&sm->a_ready, EPI_THREADS);
        for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], EPI_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    if (warp == 0) {
        // =========================== producer: x tiles + weight blocks ===========================
        if (lane == 0) {
            uint32_t stage = 0, wphase = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
                const int rows = min(TILE_M, p.batch - tile * TILE_M);
                mbar_wait(&sm->x_empty, (tcount & 1) ^ 1, p.error, 1);
                if (rows == TILE_M) {
                    mbar_expect_tx(&sm->x_full, (uint32_t)x_tile_bytes);
                    bulk_g2s(sX, p.x + (size_t)tile * TILE_M * p.D, (uint32_t)x_tile_bytes, &sm->x_full);
                } else {
                    mbar_arrive(&sm->x_full);       // ragged last tile: the epilogue warps copy it themselves
                }
                for (int i = 0; i < p.n_ops; ++i) {
                    const uint32_t bytes = sm->ops[i].w_bytes;
                    mbar_wait(&sm->w_empty[stage], wphase ^ 1, p.error, 2);
                    mbar_expect_tx(&sm->w_full[stage], bytes);
                    bulk_g2s(sW + (size_t)stage * STAGE_BYTES, p.weights + sm->ops[i].w_off, bytes, &sm->w_full[stage]);
                    if (++stage == STAGES) { stage = 0; wphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            uint32_t stage = 0, wphase = 0, a_cnt = 0, empty_cnt[2] = {0, 0};
            const uint32_t a_base = smem_u32(sA);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int i = 0; i < p.n_ops; ++i) {
                    const Op op = sm->ops[i];
                    const uint32_t acc = (op.flags & OP_ACC1) ? 1u : 0u;
                    if (op.flags & OP_WAIT_A) {
                        mbar_wait(&sm->a_ready, a_cnt & 1, p.error, 3);
                        ++a_cnt;
                        tc_fence_after();
                    }
                    if (op.flags & OP_WAIT_EMPTY) {
                        mbar_wait(&sm->acc_empty[acc], (empty_cnt[acc] & 1) ^ 1, p.error, 4);
                        ++empty_cnt[acc];
                        tc_fence_after();
                    }
                    mbar_wait(&sm->w_full[stage], wphase, p.error, 5);
                    tc_fence_after();
                    const uint32_t b_base = smem_u32(sW + (size_t)stage * STAGE_BYTES);
                    const uint32_t idesc = make_idesc(op.n);
                    const uint32_t b_lbo = (uint32_t)op.n * 16u;
                    for (uint32_t ks = 0; ks < op.ksteps; ++ks) {
                        const uint64_t da = make_desc(a_base + (op.a_slab0 + 2 * ks) * SLAB_BYTES, SLAB_BYTES, 128);
                        const uint64_t db = make_desc(b_base + ks * 2 * b_lbo, b_lbo, 128);
                        umma(tmem + op.tmem_col, da, db, idesc, ((op.flags & OP_FIRST) && ks == 0) ? 0u : 1u);
                    }
                    umma_commit(&sm->w_empty[stage]);                   // frees the ring stage when the MMAs retire
                    if (op.flags & OP_COMMIT) umma_commit(&sm->acc_full[acc]);
                    if (++stage == STAGES) { stage = 0; wphase ^= 1; }
                }
            }
        }
    } else {
        // =========================== epilogue warps ===========================
        const int et = tid - 64;                  // 0..511
        const int wg = et >> 7;                   // warpgroup 0..3
        const int row = (warp & 3) * 32 + lane;   // sample row of the tile = TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t full_cnt[2] = {0, 0}, tcount = 0;
        const int groups = p.HP / 8;              // 8-column groups of a hidden layer
        const int g_lo = wg * groups / EPI_WGS, g_hi = (wg + 1) * groups / EPI_WGS;
        float* xrow = sX + row * p.D;

        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
            const int rows = min(TILE_M, p.batch - tile * TILE_M);
            // ---- x tile -> A0 (two constant-one columns after the D inputs carry the biases) ----
            mbar_wait(&sm->x_full, tcount & 1, p.error, 6);
            if (rows < TILE_M) {
                const float* src = p.x + (size_t)tile * TILE_M * p.D;
                for (int i = et; i < TILE_M * p.D; i += EPI_THREADS) sX[i] = i < rows * p.D ? src[i] : 0.f;
                asm volatile("bar.sync 1, 512;" ::: "memory");
            }
            {
                const int slabs1 = p.K1 / 8;
                for (int j = wg; j < slabs1; j += EPI_WGS) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = j * 8 + i;
                        v[i] = k < p.D ? xrow[k] : (k < p.D + 2 ? 1.f : 0.f);
                    }
                    uint4 q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    *reinterpret_cast<uint4*>(sA + (size_t)j * SLAB_BYTES + row * 16) = q;
                }
                fence_async_smem();
                mbar_arrive(&sm->a_ready);
            }
            // ---- two hidden layers: ELU, bf16 -> A (bias already in the accumulator) ----
            for (int layer = 0; layer < 2; ++layer) {
                mbar_wait(&sm->acc_full[0], full_cnt[0] & 1, p.error, 7);
                ++full_cnt[0];
                tc_fence_after();
                for (int g = g_lo; g < g_hi; ++g) {
                    uint32_t r[8];
                    tmem_ld8(lane_addr + g * 8, r);
                    tmem_wait8(r);
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = fast_elu(__uint_as_float(r[i]));
                    uint4 q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    *reinterpret_cast<uint4*>(sA + (size_t)g * SLAB_BYTES + row * 16) = q;
                }
                fence_async_smem();
                tc_fence_before();
                mbar_arrive(&sm->a_ready);
            }
            // ---- output layer chunks: spline transformer straight out of TMEM ----
            float ld = 0.f;
            for (int c = 0; c < p.n_chunks; ++c) {
                const int b = c & 1;
                mbar_wait(&sm->acc_full[b], full_cnt[b] & 1, p.error, 8);
                ++full_cnt[b];
                tc_fence_after();
                const uint32_t col0 = lane_addr + (b ? ACC1_COL : 0);
#pragma unroll 1
                for (int jj = 0; jj < FEATS_PER_CHUNK / EPI_WGS; ++jj) {
                    const int slot = wg * (FEATS_PER_CHUNK / EPI_WGS) + jj;
                    uint32_t r[32];
                    tmem_ld32(col0 + slot * PSTRIDE, r);
                    tmem_wait8(r); tmem_wait8(r + 8); tmem_wait8(r + 16); tmem_wait8(r + 24);
                    const FeatConst fc = sFeat[c * FEATS_PER_CHUNK + slot];
                    if (fc.col >= 0) {
                        if (p.debug_params != nullptr && row < rows) {
                            float* dbg = p.debug_params + ((size_t)tile * TILE_M + row) * p.n_chunks * CHUNK_N + c * CHUNK_N +
                                         slot * PSTRIDE;
#pragma unroll
                            for (int i = 0; i < NPAR; ++i) dbg[i] = __uint_as_float(r[i]);
                        }
                        float yv;
                        ld += spline8_circular(r, xrow[fc.col], fc, p.min_bin, p.min_slope, p.slope_offset2, yv);
                        xrow[fc.col] = yv;
                    }
                }
                tc_fence_before();
                mbar_arrive(&sm->acc_empty[b]);
            }
            // ---- log-det: combine the warpgroups, store; y tile leaves with one bulk store ----
            if (wg > 0) sLd[(wg - 1) * TILE_M + row] = ld;
            fence_async_smem();                      // y tile writes -> visible to the bulk-copy engine
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (wg == 0) {
                if (row < rows) {
#pragma unroll
                    for (int g = 0; g < EPI_WGS - 1; ++g) ld += sLd[g * TILE_M + row];
                    p.logdet[(size_t)tile * TILE_M + row] = ld;
                }
                if (rows == TILE_M && et == 0) {
                    bulk_s2g(p.y + (size_t)tile * TILE_M * p.D, sX, (uint32_t)x_tile_bytes);
                    bulk_wait_read();
                    mbar_arrive(&sm->x_empty);
                }
            }
            if (rows < TILE_M) {
                float* dst = p.y + (size_t)tile * TILE_M * p.D;
                for (int i = et; i < rows * p.D; i += EPI_THREADS) dst[i] = sX[i];
                asm volatile("bar.sync 1, 512;" ::: "memory");
                if (et == 0) mbar_arrive(&sm->x_empty);
            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

size_t smem_bytes(const Params& p) {
    const size_t a_slabs = (size_t)(p.K1 > p.HP ? p.K1 : p.HP) / 8;
    size_t s = a_slabs * SLAB_BYTES + (size_t)STAGES * STAGE_BYTES;
    s += ((size_t)TILE_M * p.D * 4 + 127) & ~(size_t)127;
    s += (size_t)p.n_chunks * FEATS_PER_CHUNK * sizeof(FeatConst);
    s += (EPI_WGS - 1) * TILE_M * 4 + sizeof(Smem);
    return s + 256;
}

}  // namespace fused
}  // namespace tfepb

using namespace tfepb;

static_assert(sizeof(fused::Op) == sizeof(tfepb_fused_op), "schedule entry layout mismatch");
static_assert(sizeof(fused::FeatConst) == sizeof(tfepb_fused_feature), "feature table layout mismatch");

extern "C" int tfepb_maf_spline_forward_bf16(const tfepb_fused_args* a, tfepb_stream_t stream) {
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->x && a->y && a->logdet && a->ops && a->weights && a->feats, "null buffer");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->n_features > 0, "bad sizes");
    TFEPB_CHECK_ARG(a->k1 % 16 == 0 && a->k1 >= a->n_features + 2, "k1 must hold n_features + 2 bias columns, rounded up to 16");
    TFEPB_CHECK_ARG(a->hidden_padded % 16 == 0 && a->hidden_padded > 0 && a->hidden_padded <= 512, "bad hidden width");
    TFEPB_CHECK_ARG(a->n_chunks > 0 && a->n_ops > 0 && a->n_ops <= fused::MAX_OPS, "bad schedule length");
    TFEPB_CHECK_ARG((a->n_features * 4 * fused::TILE_M) % 16 == 0, "tile of x must be a multiple of 16 bytes");
    TFEPB_CHECK_ARG(((uintptr_t)a->x % 16 == 0) && ((uintptr_t)a->y % 16 == 0) && ((uintptr_t)a->weights % 16 == 0),
                    "x, y and the packed weights must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    fused::Params p{};
    p.x = (const float*)a->x; p.y = (float*)a->y; p.logdet = (float*)a->logdet;
    p.batch = a->batch; p.D = a->n_features; p.K1 = a->k1; p.HP = a->hidden_padded;
    p.n_chunks = a->n_chunks; p.n_ops = a->n_ops;
    p.ops = (const fused::Op*)a->ops; p.weights = (const uint8_t*)a->weights;
    p.feats = (const fused::FeatConst*)a->feats;
    p.min_bin = a->min_bin_size; p.min_slope = a->min_slope; p.slope_offset2 = a->slope_offset * fused::LOG2E;
    p.error = a->error_flag;
    p.debug_params = a->debug_params;
    const size_t smem = fused::smem_bytes(p);
    TFEPB_CHECK_ARG(smem <= 227 * 1024, "shared memory plan of %zu bytes exceeds 227 KB", smem);
    static thread_local size_t configured = 0;
    if (configured < smem) {
        TFEPB_CUDA(cudaFuncSetAttribute(fused::maf_spline_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int n_tiles = (a->batch + fused::TILE_M - 1) / fused::TILE_M;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    fused::maf_spline_fwd_kernel<<<grid, fused::THREADS, smem, as_stream(stream)>>>(p);
    return check_launch("maf_spline_fwd_kernel");
}
