// General masked linear layers on the tcgen05 tensor cores (bf16 operands, fp32 accumulation), for any MADE
// shape and for training: the three products of nn/masked.py:266-302
//     forward         Y  = act(X W^T + b)
//     backward input  dX = (dY W) * ELU'(H_prev)
//     backward weight dW += dY^T X
// The first two are C[M x N] = A[M x K] . B[N x K]^T with both operands K-major (tc_gemm_kernel); the weight gradient reduces
// over the batch and reads the SAME row images MN-major (tc_wgrad_kernel, further down), so every activation and every
// cotangent exists as ONE bf16 image.  Operand IMAGES (tfepb_tc_pack, or the epilogue of the producing product): bf16 blocks
// that are exact images of the shared-memory operand layout, so a block moves with one bulk copy.
//
// Image of an operand with R rows and K reduction elements, blocked (BR rows x 64 k), BR = 128 (A) or 256 (B):
//     block (rb, kb) at byte offset (rb * ceil(K / 64) + kb) * BR * 128;
//     inside a block: for each slab s of 8 k-values (8 slabs): BR rows x 16 bytes (K-major core matrices, no swizzle).
//
// Kernel: persistent CTAs walk the (m-tile of 128 rows, n-tile of 256 columns) output tiles, n fastest (the A rows
// of an m-tile stay in L2 across its n-tiles).  Warp 0 streams (A block, B block) stages of 48 KB through a 3-deep
// mbarrier ring with bulk copies; warp 1 issues tcgen05.mma (M = 128, N = 256, K = 16, both operands from shared
// memory) into one of TWO 256-column accumulators in tensor memory; warps 2-17 (four TMEM lane quadrants x four column
// groups; a warp's two 32-column sub-tiles are interleaved over the tile) drain the other accumulator: bias, ELU or ELU'
// multiplier (read from the fp32 activations or from their bf16 image), fp32 store through a per-warp transposition buffer,
// and optionally the bf16 image of the result = the A operand of the next layer (written coalesced: consecutive rows are
// consecutive 16-byte chunks of a slab), the image of the transposed result, the column sums.  What does not depend on the
// accumulator (bias, staged operands of a fused transformer) travels by cp.async during the accumulator wait; per-tile
// tables (k-block ranges, column table) live in shared memory.  A per-n-tile range of k-blocks skips the all-zero part of a
// degree-sorted (staircase) masked weight; split-K with fp32 atomics serves K-major weight gradients.  Opt-in cluster mode:
// two CTAs share the B operand by multicast (see the kernel).
//
// SPLIT PRECISION (n_split = 2 or 3; the fp32-class conditioner on the tensor cores): every operand is
// stored as n_split bf16 images x = x_0 + x_1 (+ x_2) with x_0 = bf16(x), x_1 = bf16(x - x_0), x_2 = bf16(x - x_0 - x_1)
// (16 / 24 significant bits), and a k-step issues the products A_i B_j with i + j < n_split -- 3 or 6 MMAs, the dropped
// cross terms are below 2^-16 / 2^-24 of the product -- into the same fp32 accumulator.  A ring stage then holds half a
// k-block (32 k) of all images of both operands; the epilogue splits the activations it hands to the next layer the
// same way.
//
// FUSED TRANSFORMER (tfepb_tc_tx; the output layer of a MAF whose transformer is affine / SOS (2 polynomials) / Moebius on
// 3-vectors; the 8-bin neural spline takes a 32-column sub-tile per feature, see tx_spline_feat): the output columns are laid out in 16-column chunks that hold the parameters of 8 / 3 / 5 whole units (the
// caller pads the packed weight rows accordingly), so the 16 accumulator values an epilogue thread reads are complete
// parameter sets of its sample.  Forward: the thread applies the transformer to its row of x, writes y and adds the
// log-det -- the (batch x parameters) matrix never exists in memory.  Backward: the same product is recomputed and the
// epilogue replaces the parameters by their cotangents (tx_math.cuh VJPs), one chunk at a time, before writing them as
// the bf16 row image (+ column sums by warp shuffles), i.e. it emits grad_parameters directly as the operand of the two
// products below it.  These variants are compiled without the code paths they never take (instruction-cache footprint).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tx_math.cuh"

namespace tfepb {
namespace tcg {

using namespace tc;

constexpr int BM = 128, BN = 256, KB = 64;
constexpr int A_BLOCK = BM * KB * 2, B_BLOCK = BN * KB * 2;   // 16 + 32 KB: one k-block of one image
constexpr int MAX_STAGES = 3;
constexpr int EPI_WARPS = 16;
constexpr int XP_LD = 33;                                  // transposition buffer: [32 columns][33] floats per epilogue warp
constexpr int XP_FLOATS = 32 * XP_LD;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int KR_SMEM = 1024;                              // k-block ranges of up to 512 n-tiles kept in shared memory
constexpr int TX_COLS_SMEM = 2048;                         // fused transformer: x-column table kept in shared memory up to this size

struct Params {
    const uint8_t* a_img; const uint8_t* b_img;
    int64_t a_split_stride, b_split_stride, out_split_stride;   // bytes between the images of the split terms
    int M, N, K, k_blocks;              // k_blocks = ceil(K / 64)
    float* C; int64_t ldc;              // fp32 output (row-major) or null
    const float* bias;                  // (N,) or null
    int act;                            // TFEPB_ACT_*
    const float* aux; int64_t ldaux;    // multiply by ELU'(aux) = (h > 0 ? 1 : h + 1), or null
    const uint8_t* aux_img; int aux_k_blocks;   // the same operand h as its bf16 image (128-row blocks, k = column), or null
    uint8_t* out_img; int out_k_blocks; // bf16 image (128-row blocks, k = column index) of the result, or null
    uint8_t* out_img_t;                 // bf16 image of the transposed result (rows = columns of C, k = rows of C), or null
    int t_rows, t_k_blocks, t_rows_padded;   // its block_rows (128 / 256), ceil(M / 64), rows rounded up to block_rows
    float* colsum;                      // (N,) += column sums of the result, or null
    const int* kranges;                 // per n-tile: [first, end) k-block, or null
    const int* row_ranges;              // per n-tile: rows [begin, end) of C that can be non-zero; other tiles are skipped, or null
    int atomic;                         // atomicAdd into C (split-K)
    int c_add;                          // C += result (no split-K)
    int mn_major, a_kblocks, b_kblocks; // operands are read MN-major out of row images (weight gradient), their k-block counts
    int cluster;                        // launched as clusters of two CTAs that share the B operand (multicast)
    const int* tile_list; int n_list;   // weight gradient: the (tm, tn) tiles the mask leaves non-zero, or null
    int k_chunk_blocks;                 // split-K: k-blocks per blockIdx.y slice, 0 = no split
    int tiles_m, tiles_n;
    int* error;
    // fused transformer (TX != 0)
    int tx_units, tx_unit_sphere;
    float tx_max_radius;
    const int* tx_cols;
    const float* tx_x; int64_t tx_ldx;
    float* tx_y; int64_t tx_ldy;
    float* tx_logdet;
    const float* tx_gy; int64_t tx_ldgy;
    const float* tx_gl;
    float* tx_gx; int64_t tx_ldgx;
    // neural spline (TFEPB_TCTX_SPLINE8): per unit 8 floats = x0, xf, y0, yf, min_bin_size, min_slope, flags (int bits: 0
    // circular, 1 identity boundary slopes, 2 / 3 learnable lower / upper bound; spline.py:166-182), unused
    const float* sp_table;
};

// Neural spline transformer, 8 bins: one unit = the 32 columns of a sub-tile (<= 27 parameters + padding).  The unit's
// parameters go to the thread's own slots of the transposition buffer (the spline math indexes them by the bin found at run
// time), x / grad_y are read and y / grad_x written directly (one element per row and unit).
__device__ __forceinline__ SplineFeat<float> tx_spline_feat(const Params& p, int u) {
    SplineFeat<float> c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.sp_table) + 2 * u), b = __ldg(reinterpret_cast<const float4*>(p.sp_table) + 2 * u + 1);
    const int flags = __float_as_int(b.z);
    c.K = 8;
    c.circular = flags & 1; c.idslopes = (flags >> 1) & 1; c.learn_lo = (flags >> 2) & 1; c.learn_hi = (flags >> 3) & 1;
    c.x0 = a.x; c.xf = a.y; c.y0 = a.z; c.yf = a.w;
    c.min_bin = b.x; c.min_slope = b.y;
    return c;
}

// TX template argument: 0 = plain product, else 2 * kind + backward with kind = TFEPB_TCTX_*
__host__ __device__ constexpr int tx_kind(int TX) { return TX >> 1; }
__host__ __device__ constexpr bool tx_bwd(int TX) { return (TX & 1) != 0; }

template <int TX> struct TxGeo {
    static constexpr int KIND = tx_kind(TX);
    static constexpr int UPC = KIND == TFEPB_TCTX_AFFINE ? 8 : KIND == TFEPB_TCTX_SOS2 ? 3 : 5;     // units per 16-column chunk
    static constexpr int PPU = KIND == TFEPB_TCTX_AFFINE ? 2 : KIND == TFEPB_TCTX_SOS2 ? 5 : 3;     // parameter columns per unit
    static constexpr int XPU = KIND == TFEPB_TCTX_MOEBIUS3 ? 3 : 1;                                // x columns per unit
    static constexpr int XPC = UPC * XPU;                                                          // x columns per chunk
    static constexpr bool SPLINE = KIND == TFEPB_TCTX_SPLINE8;      // one unit per 32-column sub-tile, no staging (see below)
    // chunks per staging group of the transposition buffer (32 columns): forward x only, backward x and grad_y
    static constexpr int GF = 4 * XPC <= 32 ? 4 : 2;
    static constexpr int GB = 4 * XPC <= 32 ? 2 : 1;
};

// The epilogue thread owns one ROW (sample) but row-major x / y want a warp instruction to cover a row segment: the
// columns cols[c0 .. c0 + XC) of the warp's 32 rows travel through buf[column][row] (pitch XP_LD).  A warp instruction
// covers 32 / CW rows x CW columns (CW = XC rounded up to a power of two).
// In: through the asynchronous copy unit (global -> shared without a register round trip; invalid rows / columns are
// zero-filled): the copies are in flight while the warp waits for the accumulator or reads tensor memory; the caller
// waits (cp_async_wait_all) and synchronises the warp before the first use.  Out: plain loads / stores.
template <int XC>
__device__ __forceinline__ void tx_stage_in_async(const float* __restrict__ src, int64_t ld, int gm0, int M, const int* __restrict__ cols,
                                                  int c0, int cn, float* buf, int lane) {
    constexpr int CW = XC <= 4 ? 4 : XC <= 8 ? 8 : XC <= 16 ? 16 : 32;
    constexpr int RPI = 32 / CW;
    const int c = lane % CW, r0 = lane / CW;
    if (c < XC) {
        const int col = c0 + c < cn ? cols[c0 + c] : -1;
        const uint32_t dst = smem_u32(buf + c * XP_LD + r0);
        const float* s = src + (int64_t)(gm0 + r0) * ld + (col >= 0 ? col : 0);
        const int64_t step = (int64_t)RPI * ld;
        if (col >= 0 && gm0 + 32 <= M) {
#pragma unroll
            for (int i = 0; i < CW; ++i, s += step)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(i * RPI * 4)), "l"(s) : "memory");
        } else {
#pragma unroll 1
            for (int i = 0; i < CW; ++i) {
                const bool ok = col >= 0 && gm0 + r0 + i * RPI < M;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (uint32_t)(i * RPI * 4)),
                             "l"(ok ? s + i * step : src), "r"(ok ? 4 : 0) : "memory");
            }
        }
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_newest() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int XC>
__device__ __forceinline__ void tx_stage_out(float* __restrict__ dst, int64_t ld, int gm0, int M, const int* __restrict__ cols,
                                             int c0, int cn, const float* buf, int lane) {
    constexpr int CW = XC <= 4 ? 4 : XC <= 8 ? 8 : XC <= 16 ? 16 : 32;
    constexpr int RPI = 32 / CW;
    const int c = lane % CW, r0 = lane / CW;
    if (c < XC && c0 + c < cn) {
        const int col = cols[c0 + c];
        const float* s = buf + c * XP_LD + r0;
        float* d = dst + (int64_t)(gm0 + r0) * ld + col;
        const int64_t step = (int64_t)RPI * ld;
        if (gm0 + 32 <= M) {
#pragma unroll
            for (int i = 0; i < CW; ++i, d += step) *d = s[i * RPI];
        } else {
#pragma unroll
            for (int i = 0; i < CW; ++i, d += step)
                if (gm0 + r0 + i * RPI < M) *d = s[i * RPI];
        }
    }
}

// Sums over the 32 rows (lanes) of the 16 columns a chunk holds, without shared memory: recursive halving -- every step a lane
// keeps one half of its columns, sends the other half to the partner that keeps it, and adds what it receives (16 shuffles);
// lane l ends with the sum of column (l >> 1) & 15 (both lanes of a pair hold it).
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (h4 ? v[8 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, h4 ? v[i] : v[8 + i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (h3 ? a[4 + i] : a[i]) + __shfl_xor_sync(0xffffffffu, h3 ? a[i] : a[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (h2 ? b[2 + i] : b[i]) + __shfl_xor_sync(0xffffffffu, h2 ? b[i] : b[2 + i], 4);
    const float d = (h1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c[0] : c[1], 2);
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

// One 16-column chunk of one row: v holds the parameters (bias added) of the chunk's units; xs / gs: the staged x (and, for
// the backward direction, grad_y) columns of the chunk, [column][row].  Forward: y replaces x in its slot, the log-det is
// accumulated.  Backward: grad_x replaces grad_y in its slot and v becomes the parameter cotangents (zeros in pad columns,
// for units beyond n_units and for rows beyond M).
template <int TX>
__device__ __forceinline__ void tx_chunk(const Params& p, float (&v)[16], float* xs, float* gs, int lane, float gl, bool row_ok,
                                         int chunk, float& ld_acc) {
    using G = TxGeo<TX>;
    constexpr int KIND = G::KIND, UPC = G::UPC, PPU = G::PPU;
    constexpr bool BWD = tx_bwd(TX);
    if constexpr (KIND == TFEPB_TCTX_MOEBIUS3) {
        // The Moebius bodies are long (three square roots, divisions, a logarithm; the VJP twice that): five unrolled
        // copies per chunk overflow the instruction cache (ncu, r02: 22 % of the stall samples were instruction fetch).
        // One rolled loop instead; the unit's three parameters are selected out of / merged into the register array.
#pragma unroll 1
        for (int j = 0; j < UPC; ++j) {
            const int u = chunk * UPC + j;
            const bool on = row_ok && u < p.tx_units;
            float w3[3], r3[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                w3[i] = v[i];
#pragma unroll
                for (int jj = 1; jj < UPC; ++jj) w3[i] = j == jj ? v[3 * jj + i] : w3[i];
            }
            if (on) {
                float x3[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) x3[i] = xs[(3 * j + i) * XP_LD + lane];
                if constexpr (!BWD) {
                    float y3[3];
                    ld_acc += moebius_eval<float>(x3, 1, w3, 1, 1.f, 3, p.tx_max_radius, p.tx_unit_sphere != 0, y3, 1);
#pragma unroll
                    for (int i = 0; i < 3; ++i) xs[(3 * j + i) * XP_LD + lane] = y3[i];
                } else {
                    float gy3[3], gx3[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) gy3[i] = gs[(3 * j + i) * XP_LD + lane];
                    moebius_vjp<float>(x3, 1, w3, 1, 3, p.tx_max_radius, p.tx_unit_sphere != 0, gy3, 1, gl, gx3, 1, r3, 1);
#pragma unroll
                    for (int i = 0; i < 3; ++i) gs[(3 * j + i) * XP_LD + lane] = gx3[i];
                }
            }
            if constexpr (BWD) {
#pragma unroll
                for (int jj = 0; jj < UPC; ++jj) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) v[3 * jj + i] = j == jj ? r3[i] : v[3 * jj + i];
                }
            }
        }
        if constexpr (BWD) v[15] = 0.f;
        return;
    }
#pragma unroll
    for (int j = 0; j < UPC; ++j) {
        const int u = chunk * UPC + j;
        const bool on = row_ok && u < p.tx_units;
        if (on) {
            const float x = xs[j * XP_LD + lane];
            float par[PPU];
#pragma unroll
            for (int i = 0; i < PPU; ++i) par[i] = v[PPU * j + i];
            if constexpr (!BWD) {
                float y, ld;
                if constexpr (KIND == TFEPB_TCTX_SOS2) sos_eval<float>(ParIn<float>{par, 1}, 2, x, y, ld);
                else affine_eval<float, false>(ParIn<float>{par, 1}, x, y, ld);
                xs[j * XP_LD + lane] = y;
                ld_acc += ld;
            } else {
                const float gy = gs[j * XP_LD + lane];
                float gx, gpar[PPU];
                if constexpr (KIND == TFEPB_TCTX_SOS2) {
                    sos_vjp<float>(ParIn<float>{par, 1}, 2, x, gy, gx, ParOut<float>{gpar, 1});     // no log-det cotangent (sos.py:233)
                } else {
                    affine_vjp<float>(ParIn<float>{par, 1}, x, gy, gl, gx, ParOut<float>{gpar, 1});
                }
                gs[j * XP_LD + lane] = gx;
#pragma unroll
                for (int i = 0; i < PPU; ++i) v[PPU * j + i] = gpar[i];
            }
        } else if constexpr (BWD) {
#pragma unroll
            for (int i = 0; i < PPU; ++i) v[PPU * j + i] = 0.f;
        }
    }
    if constexpr (BWD && UPC * PPU < 16) {
#pragma unroll
        for (int i = UPC * PPU; i < 16; ++i) v[i] = 0.f;
    }
}

struct Smem {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

// Stage geometry per split count: one term -> a whole k-block (64 k) per stage, three stages; two terms -> half a
// k-block of both terms (48 KB again), three stages; three terms -> 72 KB, two stages.
template <int NSPLIT> struct Geo {
    static constexpr int KS = NSPLIT == 1 ? KB : KB / 2;               // k extent of a stage
    static constexpr int A_PART = BM * KS * 2, B_PART = BN * KS * 2;   // one term of one operand
    static constexpr int STAGE = NSPLIT * (A_PART + B_PART);
    static constexpr int N_STAGES = NSPLIT == 3 ? 2 : 3;
    static constexpr int PARTS = KB / KS;                              // stages per k-block
};

template <int NSPLIT, int TX = 0>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const __grid_constant__ Params p) {
    using G = Geo<NSPLIT>;
    constexpr int STAGES = G::N_STAGES, STAGE_BYTES = G::STAGE;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* ring = smem_raw;
    float* xpose = reinterpret_cast<float*>(ring + (size_t)STAGES * STAGE_BYTES);
    float* bias_s = xpose + EPI_WARPS * XP_FLOATS;                 // per epilogue warp: the bias of its 64 columns
    int* cols_s = reinterpret_cast<int*>(bias_s + EPI_WARPS * 64);  // fused transformer: the x-column table (else empty)
    int* kr_s = cols_s + (TX != 0 ? TX_COLS_SMEM : 0);             // per n-tile [first, end) k-block (looked up by all three roles per tile)
    Smem* sm = reinterpret_cast<Smem*>(kr_s + KR_SMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // The column of every unit is looked up before each staging copy: from global memory that lookup was 15-25 % of the
    // stall samples of the fused epilogues (address -> copy dependency); the table is tiny, keep it in shared memory.
    const int* kranges = p.kranges;
    if (p.kranges != nullptr && 2 * p.tiles_n <= KR_SMEM) {
        for (int i = threadIdx.x; i < 2 * p.tiles_n; i += THREADS) kr_s[i] = p.kranges[i];
        kranges = kr_s;
    }
    [[maybe_unused]] const int* tx_cols = p.tx_cols;
    if constexpr (TX != 0) {
        const int n_cols = p.tx_units * TxGeo<TX>::XPU;
        if (n_cols <= TX_COLS_SMEM) {
            for (int i = tid; i < n_cols; i += THREADS) cols_s[i] = p.tx_cols[i];
            tx_cols = cols_s;                 // (visible after the __syncthreads below)
        }
    }
    const int n_tiles = p.tiles_m * p.tiles_n;

    // CLUSTER MODE (p.cluster, opt-in): every operand byte is fetched from L2 per CTA (ncu, r02: 11.1 of the 11.8 TB/s the L2
    // can send on the fused SOS forward product).  Here two CTAs form a cluster that works on two m-tiles of the SAME n-tile
    // at a time: each fetches its own A block and HALF of the B block (four of the eight k slabs, 16 KB contiguous),
    // multicast into both CTAs -- 32 KB instead of 48 KB read from L2 per CTA and k-block.  A stage may only be refilled when
    // both CTAs have consumed it: the MMA commit arrives on the `empty` barrier of both.  Measured: L2 -> SM traffic
    // 11.1 -> 7.6 TB/s, same run time (0.50 ms) -- the products are bound by the dependent-instruction latency of the
    // row-per-thread epilogue on 16 warps, not by the operand feed; kept as an option for hosts that share the L2.
    const bool cl = p.cluster != 0;
    const uint32_t cl_rank = cl ? cluster_ctarank() : 0u;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], cl ? 2 : 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    if (cl) cluster_sync_all();               // the peer's barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;
    // Tile walk.  Plain: tile t = blockIdx.x, += gridDim.x, (tm, tn) = (t / tiles_n, t % tiles_n).  Cluster mode: the
    // cluster walks (m-tile PAIR, n-tile) items and this CTA takes m-tile 2 * pair + rank; with an odd number of m-tiles
    // the last pair has a GHOST half that loads, multiplies and hands over like any other (its peer needs the B half and
    // the barrier arrivals) but stores nothing.
    const int t_first = cl ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int t_step = cl ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int t_count = cl ? ((p.tiles_m + 1) >> 1) * p.tiles_n : n_tiles;
    auto decode = [&](int t, int& tm, int& tn) -> bool {
        const int q = t / p.tiles_n;
        tn = t - q * p.tiles_n;
        tm = cl ? 2 * q + (int)cl_rank : q;
        return tm >= p.tiles_m;               // ghost
    };

    // k-block range of a tile (staircase range of its n-tile, intersected with this CTA's split-K slice)
    auto skipped = [&](int tm, int tn) -> bool {
        if (p.row_ranges == nullptr) return false;
        const int rb = p.row_ranges[2 * tn], re = p.row_ranges[2 * tn + 1];
        return tm * BM + BM <= rb || tm * BM >= re;
    };
    auto krange = [&](int tn, int& k0, int& k1) {
        k0 = 0; k1 = p.k_blocks;
        if (kranges != nullptr) { k0 = max(0, kranges[2 * tn]); k1 = min(p.k_blocks, kranges[2 * tn + 1]); }
        if (p.k_chunk_blocks > 0) {
            k0 = max(k0, (int)blockIdx.y * p.k_chunk_blocks);
            k1 = min(k1, ((int)blockIdx.y + 1) * p.k_chunk_blocks);
        }
        if (k1 < k0) k1 = k0;
    };

    if (warp == 0) {
        // =========================== producer ===========================
        uint32_t stage = 0, phase = 0;
        for (int t = t_first; t < t_count; t += t_step) {
            int tm, tn;
            if (decode(t, tm, tn)) tm = p.tiles_m - 1;                  // ghost: any valid A block
            if (skipped(tm, tn)) continue;
            int k0, k1;
            krange(tn, k0, k1);
            for (int kp = k0 * G::PARTS; kp < k1 * G::PARTS; ++kp) {
                const int kb = kp / G::PARTS, part = kp % G::PARTS;     // a stage = part `part` of k-block kb, all terms
                mbar_wait(&sm->empty[stage], phase ^ 1, p.error, 1);
                if (cl) {
                    if (elect_one()) {
                        uint8_t* dst = ring + (size_t)stage * STAGE_BYTES;
                        mbar_expect_tx(&sm->full[stage], STAGE_BYTES);      // own A + both halves of B
                        bulk_g2s(dst, p.a_img + ((size_t)tm * p.k_blocks + kb) * A_BLOCK, A_BLOCK, &sm->full[stage]);
                        const size_t half = (size_t)cl_rank * (B_BLOCK / 2);
                        bulk_g2s_multicast(dst + A_BLOCK + half, p.b_img + ((size_t)tn * p.k_blocks + kb) * B_BLOCK + half, B_BLOCK / 2,
                                           &sm->full[stage], (uint16_t)3);
                    }
                } else if (elect_one()) {
                    uint8_t* dst = ring + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&sm->full[stage], STAGE_BYTES);
                    const uint8_t* a_src = p.a_img + ((size_t)tm * p.k_blocks + kb) * A_BLOCK + (size_t)part * G::A_PART;
                    const uint8_t* b_src = p.b_img + ((size_t)tn * p.k_blocks + kb) * B_BLOCK + (size_t)part * G::B_PART;
#pragma unroll
                    for (int i = 0; i < NSPLIT; ++i) {
                        bulk_g2s(dst + i * G::A_PART, a_src + i * p.a_split_stride, G::A_PART, &sm->full[stage]);
                        bulk_g2s(dst + NSPLIT * G::A_PART + i * G::B_PART, b_src + i * p.b_split_stride, G::B_PART, &sm->full[stage]);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        uint32_t stage = 0, phase = 0, tcount = 0;
        // start-address field of the shared-memory descriptors: 14 bits of (address >> 4).  In a cluster launch the shared
        // window address of a CTA carries its rank in the upper bits; unmasked they spill into the leading-offset field
        // (first cluster version: every tile of the rank-1 CTAs came out as deterministic garbage).
        const uint32_t ring16 = (smem_u32(ring) >> 4) & 0x3fffu;
        // kind::f16, A = B = bf16, D = fp32, both K-major, M = 128, N = 256
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int t = t_first; t < t_count; t += t_step) {
            int tm, tn;
            if (decode(t, tm, tn)) tm = p.tiles_m - 1;
            if (skipped(tm, tn)) continue;
            int k0, k1;
            krange(tn, k0, k1);
            const uint32_t buf = tcount & 1;
            mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) & 1) ^ 1, p.error, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem + buf * BN;
            uint32_t accumulate = 0u;
            for (int kp = k0 * G::PARTS; kp < k1 * G::PARTS; ++kp) {
                mbar_wait(&sm->full[stage], phase, p.error, 3);
                tc_fence_after();
                const uint32_t a16 = ring16 + stage * (STAGE_BYTES >> 4);
                const uint32_t b16 = a16 + ((NSPLIT * G::A_PART) >> 4);
                if (elect_one()) {
                    // terms A_i B_j with i + j < NSPLIT, the smallest first (they meet an accumulator of their own size)
#pragma unroll
                    for (int sum = NSPLIT - 1; sum >= 0; --sum) {
#pragma unroll
                        for (int i = 0; i <= sum; ++i) {
                            const int j = sum - i;
                            const uint32_t ai = a16 + i * (G::A_PART >> 4), bj = b16 + j * (G::B_PART >> 4);
#pragma unroll
                            for (uint32_t ks = 0; ks < G::KS / 16; ++ks) {
                                // per K = 16 step two 8-k slabs: A slab = 128 rows x 16 B = 2 KB, B slab = 256 rows x 16 B = 4 KB
                                const uint64_t da = ((uint64_t)DESC_HI << 32) | (ai + ks * 256u + ((2048u >> 4) << 16));
                                const uint64_t db = ((uint64_t)DESC_HI << 32) | (bj + ks * 512u + ((4096u >> 4) << 16));
                                umma_ss(d_tmem, da, db, idesc, accumulate);
                                accumulate = 1u;
                            }
                        }
                    }
                    if (cl) umma_commit_multicast(&sm->empty[stage], (uint16_t)3);
                    else umma_commit(&sm->empty[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&sm->acc_full[buf]);
            __syncwarp();
            ++tcount;
        }
    } else {
        // =========================== epilogue: 16 warps, four column groups x four lane quadrants ===========================
        const int ew = warp - 2;
        const int cgroup = ew >> 2;                    // columns [cgroup * 64, cgroup * 64 + 64) of the tile
        const int row = (warp & 3) * 32 + lane;        // TMEM lane = row of the tile
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t tcount = 0;
        for (int t = t_first; t < t_count; t += t_step) {
            int tm, tn;
            const bool ghost = decode(t, tm, tn);
            if (!ghost && skipped(tm, tn)) continue;
            int k0, k1;
            krange(tn, k0, k1);
            const uint32_t buf = tcount & 1;
            if (ghost) {
                // cluster mode, odd number of m-tiles: this half of the last pair only keeps the hand-shakes going
                mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1, p.error, 4);
                tc_fence_before();
                mbar_arrive(&sm->acc_empty[buf]);
                ++tcount;
                continue;
            }
            // Every warp owns 32 rows (its TMEM lane quadrant) x 64 columns = four chunks of 16 columns, processed as two
            // sub-tiles of 32 columns that go through a per-warp transposition buffer: a thread holds one ROW of the
            // accumulator, but global memory wants a warp instruction to cover one row segment (128 contiguous bytes), both
            // for the ELU' operand coming in and for the fp32 result going out.
            float* xp = xpose + ew * XP_FLOATS;
            float* bias_w = bias_s + ew * 64;
            const int gm0 = tm * BM + (warp & 3) * 32;         // first row of this warp
            const int gm = gm0 + lane;
            const bool row_ok = gm < p.M;
            const bool empty = k1 <= k0;                       // nothing was accumulated: the tile is all zeros
            // The warp's two sub-tiles are INTERLEAVED over the tile (sub-tile s of the warp = 32-column block s * 4 + cgroup
            // of the tile): a last tile with few valid columns still spreads over all four column groups.
            const int gn_sub0 = tn * BN + cgroup * 32, gn_sub1 = gn_sub0 + 128;
            auto gcol = [&](int q) { return (q < 2 ? gn_sub0 : gn_sub1) + (q & 1) * 16; };     // first column of chunk q (0..3)
            // What does not depend on the accumulator is fetched BEFORE waiting for it: the bias of the warp's columns and,
            // for a fused transformer, the first staging group of x (and grad_y).
            using TG = TxGeo<TX>;
            // forward: where the buffer holds the x columns of TWO tiles (8 XPC <= 32) the copies run one tile ahead
            [[maybe_unused]] constexpr bool FWD_AHEAD = TX != 0 && !tx_bwd(TX) && TG::GF == 4 && 8 * TG::XPC <= 32;
            const bool with_bias = !p.atomic && p.bias != nullptr;
            if (with_bias) {
                // asynchronous copies (zero-filled beyond N): in flight during the accumulator wait, no register round trip
                const bool ok0 = gn_sub0 + lane < p.N, ok1 = gn_sub1 + lane < p.N;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(bias_w + lane)),
                             "l"(ok0 ? p.bias + gn_sub0 + lane : p.bias), "r"(ok0 ? 4 : 0) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(bias_w + lane + 32)),
                             "l"(ok1 ? p.bias + gn_sub1 + lane : p.bias), "r"(ok1 ? 4 : 0) : "memory");
                if constexpr (FWD_AHEAD) cp_async_commit();   // (a group of its own: the wait below leaves only the NEXT tile's pending)
            }
            [[maybe_unused]] const int tx_cn = p.tx_units * TG::XPU;            // x columns in the table
            [[maybe_unused]] float gl = 0.f;
            [[maybe_unused]] constexpr bool xp_free = true;   // the VJP variants never use the transposition buffer for anything else
            [[maybe_unused]] float* xs_tile = xp + ((FWD_AHEAD && (tcount & 1)) ? 4 * TG::XPC * XP_LD : 0);
            if constexpr (TX != 0 && TG::SPLINE) {
                if (tx_bwd(TX) && p.tx_gl != nullptr && row_ok) gl = __ldg(p.tx_gl + gm);
            } else if constexpr (TX != 0 && !tx_bwd(TX)) {
                // x of the first sub-tile's two chunks, and of the second one's where the buffer holds both (GF == 4)
                auto stage_tile = [&](int tt, float* dst) {
                    int tm2, tn2;
                    if (decode(tt, tm2, tn2)) return;
                    const int g0 = tm2 * BM + (warp & 3) * 32, s0 = tn2 * BN + cgroup * 32, s1 = s0 + 128;
                    tx_stage_in_async<2 * TG::XPC>(p.tx_x, p.tx_ldx, g0, p.M, tx_cols, (s0 >> 4) * TG::XPC, tx_cn, dst, lane);
                    if constexpr (TG::GF == 4) {
                        if (s1 < p.N)
                            tx_stage_in_async<2 * TG::XPC>(p.tx_x, p.tx_ldx, g0, p.M, tx_cols, (s1 >> 4) * TG::XPC, tx_cn,
                                                           dst + 2 * TG::XPC * XP_LD, lane);
                    }
                };
                if (!FWD_AHEAD || tcount == 0) {
                    stage_tile(t, xs_tile);
                    cp_async_commit();
                }
                if constexpr (FWD_AHEAD) {
                    // the NEXT tile of this CTA into the other half of the buffer (its previous user, the tile before this
                    // one, is finished); one commit group per tile, the wait below leaves this newest group pending
                    const int t2 = t + t_step;
                    if (t2 < t_count) stage_tile(t2, xp + ((tcount & 1) ? 0 : 4 * TG::XPC * XP_LD));
                    cp_async_commit();
                }
            } else if constexpr (TX != 0) {
                tx_stage_in_async<TG::GB * TG::XPC>(p.tx_x, p.tx_ldx, gm0, p.M, tx_cols, (gn_sub0 >> 4) * TG::XPC, tx_cn, xp, lane);
                tx_stage_in_async<TG::GB * TG::XPC>(p.tx_gy, p.tx_ldgy, gm0, p.M, tx_cols, (gn_sub0 >> 4) * TG::XPC, tx_cn,
                                               xp + TG::GB * TG::XPC * XP_LD, lane);
                if constexpr (TG::GB == 2 && 8 * TG::XPC <= 32) {
                    // the second sub-tile's operands too, when the buffer holds both and nothing else needs it (the column sums
                    // of the VJP variants are reduced with shuffles)
                    if (xp_free && gn_sub1 < p.N) {
                        tx_stage_in_async<2 * TG::XPC>(p.tx_x, p.tx_ldx, gm0, p.M, tx_cols, (gn_sub1 >> 4) * TG::XPC, tx_cn,
                                                 xp + 4 * TG::XPC * XP_LD, lane);
                        tx_stage_in_async<2 * TG::XPC>(p.tx_gy, p.tx_ldgy, gm0, p.M, tx_cols, (gn_sub1 >> 4) * TG::XPC, tx_cn,
                                                 xp + 6 * TG::XPC * XP_LD, lane);
                    }
                }
                if (p.tx_gl != nullptr && row_ok) gl = __ldg(p.tx_gl + gm);
            }
            mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1, p.error, 4);
            tc_fence_after();
            if constexpr (FWD_AHEAD) cp_async_wait_but_newest();
            else cp_async_wait_all();                          // bias / staged operands travelled during the wait
            __syncwarp();
            // accumulator chunks q, q + 1 (16 columns each) of this thread's row, bias added
            auto load_pair = [&](int q, float (&va)[16], float (&vb)[16]) {
                uint32_t ra[16], rb[16];
                // ELU' operand from the image of h: this row's 32 columns are four 16-byte pieces of consecutive slabs
                // (consecutive rows = consecutive pieces: coalesced as is); requested together with the accumulator
                uint4 h4[4];
                const bool with_aux_img = TX == 0 && p.aux_img != nullptr && !p.atomic && (gcol(q) >> 6) < p.aux_k_blocks;
                if (with_aux_img) {
                    const int gn0 = gcol(q);
                    const uint8_t* blk = p.aux_img + ((size_t)tm * p.aux_k_blocks + (gn0 >> 6)) * A_BLOCK + (size_t)row * 16 +
                                         (size_t)((gn0 & 63) >> 3) * 2048;
#pragma unroll
                    for (int j = 0; j < 4; ++j) h4[j] = __ldg(reinterpret_cast<const uint4*>(blk + (size_t)j * 2048));
                }
                if (!empty) {
                    const uint32_t tcol = buf * BN + (uint32_t)(gcol(q) - tn * BN);
                    tmem_ld16(lane_addr + tcol, ra);
                    tmem_ld16(lane_addr + tcol + 16, rb);
                    tmem_wait8(ra); tmem_wait8(ra + 8); tmem_wait8(rb); tmem_wait8(rb + 8);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) ra[i] = rb[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) { va[i] = __uint_as_float(ra[i]); vb[i] = __uint_as_float(rb[i]); }
                if (with_bias) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a4 = *reinterpret_cast<const float4*>(bias_w + q * 16 + 4 * i);
                        const float4 b4 = *reinterpret_cast<const float4*>(bias_w + q * 16 + 16 + 4 * i);
                        va[4 * i] += a4.x; va[4 * i + 1] += a4.y; va[4 * i + 2] += a4.z; va[4 * i + 3] += a4.w;
                        vb[4 * i] += b4.x; vb[4 * i + 1] += b4.y; vb[4 * i + 2] += b4.z; vb[4 * i + 3] += b4.w;
                    }
                }
                if (with_aux_img) {
                    const uint32_t hw[16] = {h4[0].x, h4[0].y, h4[0].z, h4[0].w, h4[1].x, h4[1].y, h4[1].z, h4[1].w,
                                             h4[2].x, h4[2].y, h4[2].z, h4[2].w, h4[3].x, h4[3].y, h4[3].z, h4[3].w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        va[2 * i] *= fminf(__uint_as_float(hw[i] << 16), 0.f) + 1.f;
                        va[2 * i + 1] *= fminf(__uint_as_float(hw[i] & 0xffff0000u), 0.f) + 1.f;
                        vb[2 * i] *= fminf(__uint_as_float(hw[8 + i] << 16), 0.f) + 1.f;
                        vb[2 * i + 1] *= fminf(__uint_as_float(hw[8 + i] & 0xffff0000u), 0.f) + 1.f;
                    }
                }
            };
            // one chunk (fused-transformer variants with long per-unit bodies go chunk by chunk through ONE copy of the code)
            [[maybe_unused]] auto load_one = [&](int q, float (&v)[16]) {
                uint32_t r[16];
                if (!empty) {
                    tmem_ld16(lane_addr + buf * BN + (uint32_t)(gcol(q) - tn * BN), r);
                    tmem_wait8(r); tmem_wait8(r + 8);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                if (with_bias) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a4 = *reinterpret_cast<const float4*>(bias_w + q * 16 + 4 * i);
                        v[4 * i] += a4.x; v[4 * i + 1] += a4.y; v[4 * i + 2] += a4.z; v[4 * i + 3] += a4.w;
                    }
                }
            };
            // activation / ELU' / padding, then the chunk goes to the transposition buffer and to the image of the result
            auto emit_chunk = [&](int q, float (&v)[16]) {
                const int c16 = q & 1;
                const int gn0 = gcol(q);
                // (the fused-transformer variants have no activation, ELU' operand, fp32 output or transposed image: compiled
                // out there, their instruction footprint decides how much of the epilogue stays in the instruction cache)
                if (TX == 0 && !p.atomic) {
                    if (p.act == TFEPB_ACT_ELU) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            // four instructions per element (FMUL, MUFU.EX2, FSETP, predicated FADD); ex2 of a large positive
                            // argument is +inf and unused
                            const float e = ex2(v[i] * LOG2E);
                            if (v[i] <= 0.f) v[i] = e - 1.f;
                        }
                    }
                    if (p.aux != nullptr) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] *= fminf(xp[(c16 * 16 + i) * XP_LD + lane], 0.f) + 1.f;
                    }
                    if (p.c_add) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += xp[(c16 * 16 + i) * XP_LD + lane];
                    }
                    if (gn0 + 15 >= p.N) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (gn0 + i >= p.N) v[i] = 0.f;
                    }
                }
                if (TX == 0 && (p.C != nullptr || p.out_img_t != nullptr || p.colsum != nullptr)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) xp[(c16 * 16 + i) * XP_LD + lane] = v[i];
                }
                if (p.out_img != nullptr) {
                    if (!row_ok) {
                        // rows beyond M carry the bias only; a weight-gradient product reads the image MN-major and reduces
                        // over its rows, so they must be zero
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = 0.f;
                    }
                    // columns are the reduction index of the next product: k-block = gn / 64, slab = (gn % 64) / 8;
                    // consecutive rows are consecutive 16-byte chunks of a slab (coalesced as is)
                    const int kblk = gn0 >> 6;
                    if (kblk < p.out_k_blocks) {
                        uint8_t* blk = p.out_img + ((size_t)tm * p.out_k_blocks + kblk) * A_BLOCK + (size_t)row * 16;
                        const int slab = (gn0 & 63) >> 3;
                        float res[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) res[i] = v[i];
#pragma unroll
                        for (int t = 0; t < NSPLIT; ++t) {
                            // term t of the split: bf16 of what the previous terms left over
                            uint32_t q8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                q8[i] = pack_bf16(res[2 * i], res[2 * i + 1]);
                                if (t + 1 < NSPLIT) {
                                    res[2 * i] -= __uint_as_float(q8[i] << 16);
                                    res[2 * i + 1] -= __uint_as_float(q8[i] & 0xffff0000u);
                                }
                            }
                            uint8_t* dst = blk + (size_t)t * p.out_split_stride;
                            *reinterpret_cast<uint4*>(dst + (size_t)slab * 2048) = make_uint4(q8[0], q8[1], q8[2], q8[3]);
                            *reinterpret_cast<uint4*>(dst + (size_t)(slab + 1) * 2048) = make_uint4(q8[4], q8[5], q8[6], q8[7]);
                        }
                    }
                }
            };
            if constexpr (TX != 0 && !tx_bwd(TX)) {
                // ---- fused transformer, forward: x comes in and y goes out through the (otherwise idle) transposition buffer,
                // GF chunks per staging group; y and the log-det are the only outputs ----
                float ld_acc = 0.f;
#pragma unroll 1
                for (int sub = 0; sub < 2; ++sub) {
                    const int gns = sub == 0 ? gn_sub0 : gn_sub1;
                    if (gns >= p.N) break;                                                     // warp-uniform
                    if constexpr (TG::SPLINE) {
                        float va[16], vb[16];
                        load_pair(2 * sub, va, vb);
                        const int u = gns >> 5;
#pragma unroll
                        for (int i = 0; i < 16; ++i) { xp[i * XP_LD + lane] = va[i]; xp[(16 + i) * XP_LD + lane] = vb[i]; }
                        if (row_ok && u < p.tx_units) {
                            const int col = tx_cols[u];
                            float y, ld;
                            int bin;
                            spline_eval<float, 8, false>(tx_spline_feat(p, u), ParIn<float>{xp + lane, XP_LD},
                                                         __ldg(p.tx_x + (int64_t)gm * p.tx_ldx + col), y, ld, bin);
                            p.tx_y[(int64_t)gm * p.tx_ldy + col] = y;
                            ld_acc += ld;
                        }
                        continue;
                    }
                    const int c0 = (gns >> 4) * TG::XPC;
                    float* xs = xs_tile + (TG::GF == 4 ? sub * 2 * TG::XPC * XP_LD : 0);
                    const bool staged = TG::GF != 4 && sub > 0;
                    if (staged) tx_stage_in_async<2 * TG::XPC>(p.tx_x, p.tx_ldx, gm0, p.M, tx_cols, c0, tx_cn, xs, lane);
                    if constexpr (TG::KIND == TFEPB_TCTX_MOEBIUS3) {
#pragma unroll 1
                        for (int c16 = 0; c16 < 2; ++c16) {
                            float v[16];
                            load_one(2 * sub + c16, v);
                            if (staged && c16 == 0) { cp_async_wait_all(); __syncwarp(); }
                            tx_chunk<TX>(p, v, xs + c16 * TG::XPC * XP_LD, nullptr, lane, 0.f, row_ok, (gns >> 4) + c16, ld_acc);
                        }
                    } else {
                        float va[16], vb[16];
                        load_pair(2 * sub, va, vb);
                        if (staged) { cp_async_wait_all(); __syncwarp(); }
                        tx_chunk<TX>(p, va, xs, nullptr, lane, 0.f, row_ok, gns >> 4, ld_acc);
                        tx_chunk<TX>(p, vb, xs + TG::XPC * XP_LD, nullptr, lane, 0.f, row_ok, (gns >> 4) + 1, ld_acc);
                    }
                    __syncwarp();
                    tx_stage_out<2 * TG::XPC>(p.tx_y, p.tx_ldy, gm0, p.M, tx_cols, c0, tx_cn, xs, lane);
                    __syncwarp();
                }
                if (row_ok && p.tx_logdet != nullptr) atomicAdd(p.tx_logdet + gm, ld_acc);
            } else {
#pragma unroll 1
                for (int sub = 0; sub < 2; ++sub) {
                    const int gns = sub == 0 ? gn_sub0 : gn_sub1;  // first column of the sub-tile
                    if (gns >= p.N && p.out_img == nullptr && (p.out_img_t == nullptr || gns >= p.t_rows_padded)) continue;   // warp-uniform
                    if (TX == 0 && p.aux != nullptr && !p.atomic) {
                        // coalesced load of the 32 x 32 ELU' operand: lane = column, transposed into the buffer
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const int grow = gm0 + r;
                            xp[lane * XP_LD + r] = (grow < p.M && gns + lane < p.N) ? __ldg(p.aux + (int64_t)grow * p.ldaux + gns + lane) : 0.f;
                        }
                        __syncwarp();
                    }
                    if (TX == 0 && p.c_add) {
                        // C += result: the old values come in through the transposition buffer (asynchronous copies, all 32 of the
                        // lane in flight while the accumulator is read; waited for right before the chunks are emitted)
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const int grow = gm0 + r;
                            const bool ok = grow < p.M && gns + lane < p.N;
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(xp + lane * XP_LD + r)),
                                         "l"(ok ? p.C + (int64_t)grow * p.ldc + gns + lane : p.C), "r"(ok ? 4 : 0) : "memory");
                        }
                    }
                    if constexpr (TX != 0 && TG::SPLINE) {
                        // ---- neural spline, backward: the sub-tile is one unit; parameters -> cotangents in place in the
                        // thread's slots of the buffer (spline_vjp reads every parameter before it overwrites its range) ----
                        float va[16], vb[16];
                        load_pair(2 * sub, va, vb);
                        const int u = gns >> 5;
#pragma unroll
                        for (int i = 0; i < 16; ++i) { xp[i * XP_LD + lane] = va[i]; xp[(16 + i) * XP_LD + lane] = vb[i]; }
                        const bool on = row_ok && u < p.tx_units;
                        int n_par = 0;
                        if (on) {
                            const SplineFeat<float> feat = tx_spline_feat(p, u);
                            n_par = feat.n_params();
                            const int col = tx_cols[u];
                            float gx;
                            spline_vjp<float, 8>(feat, ParIn<float>{xp + lane, XP_LD}, __ldg(p.tx_x + (int64_t)gm * p.tx_ldx + col),
                                                 __ldg(p.tx_gy + (int64_t)gm * p.tx_ldgy + col), gl, gx, ParOut<float>{xp + lane, XP_LD});
                            p.tx_gx[(int64_t)gm * p.tx_ldgx + col] = gx;
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            va[i] = i < n_par ? xp[i * XP_LD + lane] : 0.f;
                            vb[i] = 16 + i < n_par ? xp[(16 + i) * XP_LD + lane] : 0.f;
                        }
                        if (p.colsum != nullptr) {
                            const float sa = warp_colsum16(va, lane), sb = warp_colsum16(vb, lane);
                            const int n = gns + ((lane >> 1) & 15);
                            if ((lane & 1) == 0) {
                                if (n < p.N) atomicAdd(p.colsum + n, sa);
                                if (n + 16 < p.N) atomicAdd(p.colsum + n + 16, sb);
                            }
                        }
                        emit_chunk(2 * sub, va);
                        emit_chunk(2 * sub + 1, vb);
                        __syncwarp();
                        continue;
                    }
                    if constexpr (TX != 0) {
                        // ---- fused transformer, backward, one chunk at a time through ONE copy of the code (the size of the
                        // epilogue loop decides how much of it the instruction cache holds): accumulator -> parameter cotangents
                        // (x and grad_y wait in the staging buffer, grad_x leaves through the slots of grad_y) -> bias gradient
                        // (shuffles) -> image.  Staging: GB == 1 one chunk at a time; GB == 2 a sub-tile at a time, both
                        // sub-tiles before the accumulator wait where the buffer holds them (8 XPC <= 32) ----
                        constexpr int XPC = TG::XPC;
                        constexpr bool BOTH = TG::GB == 2 && 8 * XPC <= 32;
#pragma unroll 1
                        for (int c16 = 0; c16 < 2; ++c16) {
                            const int q = 2 * sub + c16;
                            float unused = 0.f;
                            float v[16];
                            float *xs, *gs;
                            if constexpr (TG::GB == 1) {
                                xs = xp; gs = xp + XPC * XP_LD;
                                if (q > 0) {
                                    const int c0 = ((gns >> 4) + c16) * XPC;
                                    tx_stage_in_async<XPC>(p.tx_x, p.tx_ldx, gm0, p.M, tx_cols, c0, tx_cn, xs, lane);
                                    tx_stage_in_async<XPC>(p.tx_gy, p.tx_ldgy, gm0, p.M, tx_cols, c0, tx_cn, gs, lane);
                                }
                            } else {
                                float* base = xp + ((BOTH && sub > 0) ? 4 * XPC * XP_LD : 0);
                                if (!BOTH && sub > 0 && c16 == 0) {
                                    const int c0 = (gns >> 4) * XPC;
                                    tx_stage_in_async<2 * XPC>(p.tx_x, p.tx_ldx, gm0, p.M, tx_cols, c0, tx_cn, base, lane);
                                    tx_stage_in_async<2 * XPC>(p.tx_gy, p.tx_ldgy, gm0, p.M, tx_cols, c0, tx_cn, base + 2 * XPC * XP_LD, lane);
                                }
                                xs = base + c16 * XPC * XP_LD;
                                gs = xs + 2 * XPC * XP_LD;
                            }
                            load_one(q, v);
                            cp_async_wait_all();                   // (the copies overlapped the tensor-memory load)
                            __syncwarp();
                            tx_chunk<TX>(p, v, xs, gs, lane, gl, row_ok, (gns >> 4) + c16, unused);
                            if (TG::GB == 1 || c16 == 1) {
                                __syncwarp();
                                if constexpr (TG::GB == 1)
                                    tx_stage_out<XPC>(p.tx_gx, p.tx_ldgx, gm0, p.M, tx_cols, ((gns >> 4) + c16) * XPC, tx_cn, gs, lane);
                                else
                                    tx_stage_out<2 * XPC>(p.tx_gx, p.tx_ldgx, gm0, p.M, tx_cols, (gns >> 4) * XPC, tx_cn,
                                                          gs - XPC * XP_LD, lane);
                            }
                            if (p.colsum != nullptr) {
                                const float sa = warp_colsum16(v, lane);
                                const int n = gns + c16 * 16 + ((lane >> 1) & 15);
                                if ((lane & 1) == 0 && n < p.N) atomicAdd(p.colsum + n, sa);
                            }
                            emit_chunk(q, v);
                            __syncwarp();
                        }
                        continue;
                    }
                    float va[16], vb[16];
                    load_pair(2 * sub, va, vb);
                    if (TX == 0 && p.c_add) { cp_async_wait_all(); __syncwarp(); }
                    emit_chunk(2 * sub, va);
                    emit_chunk(2 * sub + 1, vb);
                    if (TX == 0 && (p.out_img_t != nullptr || p.colsum != nullptr)) {
                        // lane = column of C = row of the transposed image; its 32 k-values (rows gm0 .. gm0 + 31 of C) are
                        // four 16-byte chunks of consecutive slabs, and consecutive lanes write consecutive chunks
                        __syncwarp();
                        const int n = gns + lane;
                        const int nrows = min(32, p.M - gm0);              // rows beyond M carry the bias only: zero them
                        float cs = 0.f;
                        uint32_t q[16];
                        if (nrows == 32) {
#pragma unroll
                            for (int r = 0; r < 32; r += 2) {
                                const float v0 = xp[lane * XP_LD + r], v1 = xp[lane * XP_LD + r + 1];
                                cs += v0 + v1;
                                q[r >> 1] = pack_bf16(v0, v1);
                            }
                        } else {
#pragma unroll
                            for (int r = 0; r < 32; r += 2) {
                                const float v0 = r < nrows ? xp[lane * XP_LD + r] : 0.f;
                                const float v1 = r + 1 < nrows ? xp[lane * XP_LD + r + 1] : 0.f;
                                cs += v0 + v1;
                                q[r >> 1] = pack_bf16(v0, v1);
                            }
                        }
                        if (p.out_img_t != nullptr && n < p.t_rows_padded && (gm0 >> 6) < p.t_k_blocks) {
                            const size_t block_bytes = (size_t)p.t_rows * 128;
                            uint8_t* blk = p.out_img_t + ((size_t)(n / p.t_rows) * p.t_k_blocks + (gm0 >> 6)) * block_bytes +
                                           (size_t)(n % p.t_rows) * 16;
                            const int slab0 = (gm0 & 63) >> 3;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<uint4*>(blk + (size_t)(slab0 + j) * p.t_rows * 16) =
                                    make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
                        }
                        if (TX == 0 && p.colsum != nullptr && n < p.N) atomicAdd(p.colsum + n, cs);
                        if (p.C == nullptr) __syncwarp();
                    }
                    if (TX == 0 && p.C != nullptr) {
                        __syncwarp();
                        const bool c0 = gns + lane < p.N;
                        if (c0 && !(p.atomic && empty)) {
                            const int nrows = min(32, p.M - gm0);
                            float* cptr = p.C + (int64_t)gm0 * p.ldc + gns + lane;
                            if (p.atomic) {
#pragma unroll 4
                                for (int r = 0; r < nrows; ++r) atomicAdd(cptr + (int64_t)r * p.ldc, xp[lane * XP_LD + r]);
                            } else {
#pragma unroll 8
                                for (int r = 0; r < nrows; ++r) cptr[(int64_t)r * p.ldc] = xp[lane * XP_LD + r];
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&sm->acc_empty[buf]);
            ++tcount;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (cl) cluster_sync_all();               // no multicast / remote arrival may find this CTA gone
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// Weight gradient straight from ROW images: dW[m x n] = dY^T X with dY given as the (k x m) row image the backward-input
// product reads K-major and X as the (k x n) row image the forward product read -- here both are read MN-major and the
// reduction runs over their rows, so no transposed copy of any activation or cotangent exists.
//
// In a row image the 128 rows of a row block are contiguous inside every 8-column slab (2 KB) and the slabs of a row block
// follow each other across its k-blocks, i.e. the operand of a 128-row reduction block is ONE contiguous run: 16 slabs =
// 32 KB for 128 output rows (A), 32 slabs = 64 KB for 256 output columns (B).  A ring stage is therefore a whole 128-row
// block of both operands (96 KB, two stages, two bulk copies per stage; a first version that took 64-row half slabs as 48
// separate 1 KB copies per stage ran 3.3 x slower: the copy engine is bound by the number of copies).  MN-major no-swizzle
// descriptors (cute/atom/mma_traits_sm100.hpp, canonical INTERLEAVE layouts): a core matrix is 8 k x 16 bytes of MN,
// leading offset = 128 B between 8-k groups, stride offset = 2 KB between 8-element MN groups; instruction descriptor
// bits 15 / 16 = A / B MN-major.  Split-K over the row blocks, fp32 vector atomics straight from the accumulator rows.
// ------------------------------------------------------------------------------------------
constexpr int WG_STAGE = 2 * A_BLOCK + 2 * B_BLOCK;       // 32 + 64 KB
constexpr int WG_STAGES = 2;

__global__ void __launch_bounds__(THREADS, 1) tc_wgrad_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* ring = smem_raw;
    Smem* sm = reinterpret_cast<Smem*>(ring + (size_t)WG_STAGES * WG_STAGE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = p.tile_list != nullptr ? p.n_list : p.tiles_m * p.tiles_n;
    const int row_blocks = (p.K + 127) / 128;             // reduction blocks of 128 image rows

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    // tile t: from the compact list of tiles the mask leaves non-zero (all CTAs then carry the same number of real tiles), or
    // the full grid with the row-range test
    auto tile_of = [&](int t, int& tm, int& tn) {
        if (p.tile_list != nullptr) { tm = p.tile_list[2 * t]; tn = p.tile_list[2 * t + 1]; }
        else { tm = t / p.tiles_n; tn = t - tm * p.tiles_n; }
    };
    auto skipped = [&](int tm, int tn) -> bool {
        if (p.row_ranges == nullptr || p.tile_list != nullptr) return false;
        const int rb = p.row_ranges[2 * tn], re = p.row_ranges[2 * tn + 1];
        return tm * BM + BM <= rb || tm * BM >= re;
    };
    // this CTA's slice of the reduction (k_chunk_blocks counts 128-row blocks here)
    int r0 = 0, r1 = row_blocks;
    if (p.k_chunk_blocks > 0) {
        r0 = min(row_blocks, (int)blockIdx.y * p.k_chunk_blocks);
        r1 = min(row_blocks, r0 + p.k_chunk_blocks);
    }
    const int a_slabs = p.a_kblocks * 8, b_slabs = p.b_kblocks * 8;     // slabs per row block of the two images

    if (warp == 0) {
        uint32_t stage = 0, phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int tm, tn;
            tile_of(t, tm, tn);
            if (skipped(tm, tn)) continue;
            // slabs of the tile that exist in the images (the rest of the stage keeps stale bytes: they only feed rows /
            // columns of C beyond m / n, which are never stored)
            const uint32_t a_bytes = (uint32_t)min(16, a_slabs - tm * 16) * 2048u;
            const uint32_t b_bytes = (uint32_t)min(32, b_slabs - tn * 32) * 2048u;
            for (int rb = r0; rb < r1; ++rb) {
                mbar_wait(&sm->empty[stage], phase ^ 1, p.error, 1);
                if (elect_one()) {
                    uint8_t* dst = ring + (size_t)stage * WG_STAGE;
                    mbar_expect_tx(&sm->full[stage], a_bytes + b_bytes);
                    bulk_g2s(dst, p.a_img + ((size_t)rb * a_slabs + (size_t)tm * 16) * 2048, a_bytes, &sm->full[stage]);
                    bulk_g2s(dst + 2 * A_BLOCK, p.b_img + ((size_t)rb * b_slabs + (size_t)tn * 32) * 2048, b_bytes, &sm->full[stage]);
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        uint32_t stage = 0, phase = 0, tcount = 0;
        const uint32_t ring16 = smem_u32(ring) >> 4;
        // kind::f16, A = B = bf16, D = fp32, both MN-major, M = 128, N = 256
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(BM >> 4) << 24);
        constexpr uint32_t HI_MN = (2048u >> 4) | (1u << 14);            // stride offset 2 KB, descriptor version 1
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int tm, tn;
            tile_of(t, tm, tn);
            if (skipped(tm, tn)) continue;
            const uint32_t buf = tcount & 1;
            mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) & 1) ^ 1, p.error, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem + buf * BN;
            uint32_t accumulate = 0u;
            for (int rb = r0; rb < r1; ++rb) {
                mbar_wait(&sm->full[stage], phase, p.error, 3);
                tc_fence_after();
                const uint32_t a16 = ring16 + stage * (WG_STAGE >> 4);
                const uint32_t b16 = a16 + ((2 * A_BLOCK) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (uint32_t ks = 0; ks < 128 / 16; ++ks) {
                        // K = 16 step = two 8-row groups = 256 B further into every slab
                        const uint64_t da = ((uint64_t)HI_MN << 32) | (a16 + ks * 16u + ((128u >> 4) << 16));
                        const uint64_t db = ((uint64_t)HI_MN << 32) | (b16 + ks * 16u + ((128u >> 4) << 16));
                        umma_ss(d_tmem, da, db, idesc, accumulate);
                        accumulate = 1u;
                    }
                    umma_commit(&sm->empty[stage]);
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&sm->acc_full[buf]);
            __syncwarp();
            ++tcount;
        }
    } else {
        // epilogue: the warp's 32 rows x 64 columns go to C with 16-byte vector atomics (a thread owns a row)
        const int ew = warp - 2;
        const int cgroup = ew >> 2;
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const bool vec = (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (p.ldc & 3) == 0;
        uint32_t tcount = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int tm, tn;
            tile_of(t, tm, tn);
            if (skipped(tm, tn)) continue;
            const uint32_t buf = tcount & 1;
            mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1, p.error, 4);
            tc_fence_after();
            const int gm = tm * BM + (warp & 3) * 32 + lane;
            if (r1 > r0) {
#pragma unroll 1
                for (int q = 0; q < 4; q += 2) {
                    const int gn0 = tn * BN + cgroup * 64 + q * 16;
                    if (gn0 >= p.N) break;                               // warp-uniform
                    uint32_t ra[16], rb16[16];
                    tmem_ld16(lane_addr + buf * BN + cgroup * 64 + q * 16, ra);
                    tmem_ld16(lane_addr + buf * BN + cgroup * 64 + q * 16 + 16, rb16);
                    tmem_wait8(ra); tmem_wait8(ra + 8); tmem_wait8(rb16); tmem_wait8(rb16 + 8);
                    if (gm < p.M) {
                        float* crow = p.C + (int64_t)gm * p.ldc + gn0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t* r = h == 0 ? ra : rb16;
                            const int gnh = gn0 + h * 16;
                            if (vec && gnh + 15 < p.N) {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    atomicAdd(reinterpret_cast<float4*>(crow + h * 16) + i,
                                              make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                          __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (gnh + i < p.N) atomicAdd(crow + h * 16 + i, __uint_as_float(r[i]));
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&sm->acc_empty[buf]);
            ++tcount;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// fp32 row-major (or its transpose) -> bf16 image.  One CTA per (row block, k block); 256 threads.
//   transpose == 0: element (row, k) = src[row * ld + k];  transpose == 1: element (row, k) = src[k * ld + row].
// term `t` of the bf16 split of 8 values (in place: v becomes the remainder)
__device__ __forceinline__ uint4 split_term(float (&v)[8], bool more) {
    uint32_t q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
        if (more) {
            v[2 * i] -= __uint_as_float(q[i] << 16);
            v[2 * i + 1] -= __uint_as_float(q[i] & 0xffff0000u);
        }
    }
    return make_uint4(q[0], q[1], q[2], q[3]);
}

__global__ void __launch_bounds__(256) tc_pack_kernel(const float* __restrict__ src, int64_t ld, int rows, int K, int block_rows,
                                                      int transpose, uint8_t* __restrict__ img, int n_split, int64_t split_stride) {
    __shared__ float tile[64][65];
    const int k_blocks = (K + KB - 1) / KB;
    const int rb = blockIdx.y, kb = blockIdx.x;
    uint8_t* blk = img + ((size_t)rb * k_blocks + kb) * (size_t)block_rows * 128;
    const int r0 = rb * block_rows, k0 = kb * KB;
    if (!transpose) {
        // thread -> (row, slab): 8 consecutive k per thread; consecutive threads = consecutive rows (coalesced image write)
        for (int i = threadIdx.x; i < block_rows * 8; i += 256) {
            const int row = i % block_rows, slab = i / block_rows;
            const int gr = r0 + row, gk = k0 + slab * 8;
            float v[8];
            if (gr < rows && gk + 7 < K && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                // 32 contiguous, 16-byte aligned bytes of the row: two vector loads instead of eight scalar ones
                const float4* s4 = reinterpret_cast<const float4*>(src + (int64_t)gr * ld + gk);
                const float4 lo = __ldg(s4), hi = __ldg(s4 + 1);
                v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = (gr < rows && gk + j < K) ? __ldg(src + (int64_t)gr * ld + gk + j) : 0.f;
            }
            for (int t = 0; t < n_split; ++t)
                *reinterpret_cast<uint4*>(blk + (size_t)t * split_stride + (size_t)slab * block_rows * 16 + (size_t)row * 16) =
                    split_term(v, t + 1 < n_split);
        }
    } else {
        // 64 rows of the image at a time: read src[k][row] coalesced along rows, transpose through shared memory
        for (int rs = 0; rs < block_rows; rs += 64) {
            for (int i = threadIdx.x; i < 64 * 64; i += 256) {
                const int kk = i / 64, rr = i % 64;
                const int gr = r0 + rs + rr, gk = k0 + kk;
                tile[kk][rr] = (gr < rows && gk < K) ? __ldg(src + (int64_t)gk * ld + gr) : 0.f;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < 64 * 8; i += 256) {
                const int rr = i % 64, slab = i / 64;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = tile[slab * 8 + j][rr];
                for (int t = 0; t < n_split; ++t)
                    *reinterpret_cast<uint4*>(blk + (size_t)t * split_stride + (size_t)slab * block_rows * 16 +
                                              (size_t)(rs + rr) * 16) = split_term(v, t + 1 < n_split);
            }
            __syncthreads();
        }
    }
}

// One pass over an fp32 matrix (R x C, row-major) -> the image of the matrix (A operand: 128-row blocks, k = columns), the
// image of its transpose (rows = columns, k = rows; block_rows = t_rows) and the column sums: what the backward pass
// needs of grad_y (operand of the backward-input product, operand of the weight gradient, bias gradient) and the forward
// pass of x.  One CTA per 64 x 64 tile; every block of both images is written completely (zero padding included).
__global__ void __launch_bounds__(256) tc_pack_dual_kernel(const float* __restrict__ src, int64_t ld, int R, int C,
                                                           uint8_t* __restrict__ img, int a_rows, uint8_t* __restrict__ img_t,
                                                           int t_rows, float* __restrict__ colsum) {
    __shared__ float tile[64][65];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int rr = i >> 6, cc = i & 63;
        tile[rr][cc] = (r0 + rr < R && c0 + cc < C) ? __ldg(src + (int64_t)(r0 + rr) * ld + c0 + cc) : 0.f;
    }
    __syncthreads();
    const int kb_a = (C + KB - 1) / KB, kb_t = (R + KB - 1) / KB;
    if (img != nullptr && blockIdx.x < kb_a) {
        uint8_t* blk = img + ((size_t)(r0 / a_rows) * kb_a + blockIdx.x) * (size_t)a_rows * 128;
        for (int i = threadIdx.x; i < 64 * 8; i += 256) {
            const int rr = i & 63, slab = i >> 6;
            uint4 q;
            q.x = pack_bf16(tile[rr][slab * 8 + 0], tile[rr][slab * 8 + 1]);
            q.y = pack_bf16(tile[rr][slab * 8 + 2], tile[rr][slab * 8 + 3]);
            q.z = pack_bf16(tile[rr][slab * 8 + 4], tile[rr][slab * 8 + 5]);
            q.w = pack_bf16(tile[rr][slab * 8 + 6], tile[rr][slab * 8 + 7]);
            *reinterpret_cast<uint4*>(blk + (size_t)slab * a_rows * 16 + (size_t)((r0 % a_rows) + rr) * 16) = q;
        }
    }
    if (img_t != nullptr && blockIdx.y < kb_t) {
        uint8_t* blk = img_t + ((size_t)(c0 / t_rows) * kb_t + blockIdx.y) * (size_t)t_rows * 128;
        for (int i = threadIdx.x; i < 64 * 8; i += 256) {
            const int cc = i & 63, slab = i >> 6;
            uint4 q;
            q.x = pack_bf16(tile[slab * 8 + 0][cc], tile[slab * 8 + 1][cc]);
            q.y = pack_bf16(tile[slab * 8 + 2][cc], tile[slab * 8 + 3][cc]);
            q.z = pack_bf16(tile[slab * 8 + 4][cc], tile[slab * 8 + 5][cc]);
            q.w = pack_bf16(tile[slab * 8 + 6][cc], tile[slab * 8 + 7][cc]);
            *reinterpret_cast<uint4*>(blk + (size_t)slab * t_rows * 16 + (size_t)((c0 % t_rows) + cc) * 16) = q;
        }
    }
    if (colsum != nullptr && threadIdx.x < 64 && c0 + threadIdx.x < C) {
        float s = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 64; ++rr) s += tile[rr][threadIdx.x];
        atomicAdd(colsum + c0 + threadIdx.x, s);
    }
}

}  // namespace tcg
}  // namespace tfepb

using namespace tfepb;

extern "C" int64_t tfepb_tc_image_bytes(int64_t rows, int64_t k, int32_t block_rows) {
    if (rows < 0 || k < 0 || (block_rows != 128 && block_rows != 256)) return -1;
    const int64_t rb = (rows + block_rows - 1) / block_rows, kb = (k + tcg::KB - 1) / tcg::KB;
    return rb * kb * (int64_t)block_rows * 128;
}

extern "C" int tfepb_tc_pack_split(const float* src, int64_t ld, int32_t rows, int32_t k, int32_t block_rows, int32_t transpose,
                                   int32_t n_split, void* image, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(src && image, "null buffer");
    TFEPB_CHECK_ARG(rows > 0 && k > 0, "bad sizes");
    TFEPB_CHECK_ARG(n_split >= 1 && n_split <= 3, "n_split must be 1, 2 or 3");
    TFEPB_CHECK_ARG(block_rows == 128 || block_rows == 256, "block_rows must be 128 (A operand) or 256 (B operand)");
    TFEPB_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 15) == 0, "the image must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    dim3 grid((unsigned)((k + tcg::KB - 1) / tcg::KB), (unsigned)((rows + block_rows - 1) / block_rows));
    tcg::tc_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, ld, rows, k, block_rows, transpose, (uint8_t*)image, n_split,
                                                            tfepb_tc_image_bytes(rows, k, block_rows));
    return check_launch("tc_pack");
}

extern "C" int tfepb_tc_pack(const float* src, int64_t ld, int32_t rows, int32_t k, int32_t block_rows, int32_t transpose,
                             void* image, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(src && image, "null buffer");
    TFEPB_CHECK_ARG(rows > 0 && k > 0, "bad sizes");
    TFEPB_CHECK_ARG(block_rows == 128 || block_rows == 256, "block_rows must be 128 (A operand) or 256 (B operand)");
    TFEPB_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 15) == 0, "the image must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    dim3 grid((unsigned)((k + tcg::KB - 1) / tcg::KB), (unsigned)((rows + block_rows - 1) / block_rows));
    tcg::tc_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, ld, rows, k, block_rows, transpose, (uint8_t*)image, 1, 0);
    return check_launch("tc_pack");
}

extern "C" int tfepb_tc_pack_dual(const float* src, int64_t ld, int32_t rows, int32_t cols, void* image, int32_t block_rows,
                                  void* image_t, int32_t t_block_rows, float* column_sums, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(src != nullptr && (image != nullptr || image_t != nullptr), "null buffer");
    TFEPB_CHECK_ARG(rows > 0 && cols > 0 && ld >= cols, "bad sizes");
    TFEPB_CHECK_ARG(image_t == nullptr || t_block_rows == 128 || t_block_rows == 256, "t_block_rows must be 128 or 256");
    TFEPB_CHECK_ARG(image == nullptr || block_rows == 128 || block_rows == 256, "block_rows must be 128 or 256");
    TFEPB_CHECK_ARG(((uintptr_t)image % 16 == 0) && ((uintptr_t)image_t % 16 == 0), "operand images must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    const int tr = image_t != nullptr ? t_block_rows : 128;
    const int ar = image != nullptr ? block_rows : 128;
    // cover every block of both images: rows up to a multiple of the block, columns up to a multiple of the transposed block
    const unsigned gy = (unsigned)((rows + ar - 1) / ar * (ar / 64));
    const unsigned gx = (unsigned)((cols + tr - 1) / tr * (tr / 64));
    tcg::tc_pack_dual_kernel<<<dim3(gx, gy), 256, 0, as_stream(stream)>>>(src, ld, rows, cols, (uint8_t*)image, ar,
                                                                         (uint8_t*)image_t, tr, column_sums);
    return check_launch("tc_pack_dual");
}

extern "C" int tfepb_tc_gemm(const tfepb_tc_gemm_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->a_image && a->b_image, "null operand image");
    TFEPB_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0, "bad sizes");
    const tfepb_tc_tx* tx = a->tx;
    TFEPB_CHECK_ARG(a->c != nullptr || a->out_image != nullptr || a->out_image_t != nullptr || (tx != nullptr && !tx->backward),
                    "no output");
    if (tx != nullptr) {
        TFEPB_CHECK_ARG(tx->kind >= TFEPB_TCTX_AFFINE && tx->kind <= TFEPB_TCTX_SPLINE8, "unknown fused transformer kind %d", tx->kind);
        TFEPB_CHECK_ARG(a->n_split <= 1 && a->split_k <= 1 && a->aux == nullptr && a->aux_image == nullptr &&
                            a->activation == TFEPB_ACT_NONE && a->n % 16 == 0,
                        "fused transformer: plain bf16 product without split-K / aux / activation, n a multiple of 16");
        const int upc = tx->kind == TFEPB_TCTX_AFFINE ? 8 : tx->kind == TFEPB_TCTX_SOS2 ? 3 : 5;
        if (tx->kind == TFEPB_TCTX_SPLINE8) {
            TFEPB_CHECK_ARG(tx->n_units > 0 && (int64_t)a->n >= (int64_t)tx->n_units * 32,
                            "fused spline transformer: n = %d columns do not hold %d units of 32 columns", a->n, tx->n_units);
            TFEPB_CHECK_ARG(tx->spline_table != nullptr && ((uintptr_t)tx->spline_table & 15) == 0,
                            "fused spline transformer: null / misaligned unit table");
        } else {
            TFEPB_CHECK_ARG(tx->n_units > 0 && (int64_t)a->n >= ((int64_t)tx->n_units + upc - 1) / upc * 16,
                            "fused transformer: n = %d columns do not hold %d units", a->n, tx->n_units);
        }
        TFEPB_CHECK_ARG(tx->cols != nullptr && tx->x != nullptr, "fused transformer: null buffer");
        TFEPB_CHECK_ARG(tx->unit_sphere == 0 || tx->unit_sphere == 1, "fused transformer: Moebius variant must be 0 or 1");
        if (!tx->backward) {
            TFEPB_CHECK_ARG(tx->y != nullptr, "fused transformer: null y");
            TFEPB_CHECK_ARG(a->c == nullptr && a->out_image == nullptr && a->out_image_t == nullptr && a->column_sums == nullptr,
                            "fused transformer, forward: the product has no output of its own");
        } else {
            TFEPB_CHECK_ARG(tx->grad_y != nullptr && tx->grad_x != nullptr, "fused transformer: null gradient buffer");
            TFEPB_CHECK_ARG(a->c == nullptr && a->out_image_t == nullptr && a->out_image != nullptr,
                            "fused transformer, backward: the parameter cotangents leave as out_image (+ column_sums) only");
        }
    }
    TFEPB_CHECK_ARG(a->out_image_t == nullptr || a->out_image_t_rows == 128 || a->out_image_t_rows == 256,
                    "out_image_t_rows must be 128 or 256");
    TFEPB_CHECK_ARG((uintptr_t)a->out_image_t % 16 == 0, "operand images must be 16-byte aligned");
    TFEPB_CHECK_ARG(a->c == nullptr || a->ldc >= a->n, "leading dimension smaller than the row length");
    TFEPB_CHECK_ARG(!(a->split_k > 1) || (a->c != nullptr && a->out_image == nullptr && a->out_image_t == nullptr &&
                                          a->column_sums == nullptr && a->bias == nullptr && a->aux == nullptr && a->aux_image == nullptr &&
                                          a->activation == TFEPB_ACT_NONE),
                    "split-K accumulates raw products into a zero-filled fp32 C only");
    TFEPB_CHECK_ARG(((uintptr_t)a->a_image % 16 == 0) && ((uintptr_t)a->b_image % 16 == 0) && ((uintptr_t)a->out_image % 16 == 0),
                    "operand images must be 16-byte aligned");
    TFEPB_CHECK_ARG(a->aux_image == nullptr || (a->aux == nullptr && (uintptr_t)a->aux_image % 16 == 0),
                    "aux and aux_image are alternatives; images are 16-byte aligned");
    TFEPB_CHECK_ARG(a->aux_image == nullptr || (a->activation == TFEPB_ACT_NONE && a->n_split <= 1),
                    "aux_image: plain bf16 product without activation");
    TFEPB_CHECK_ARG(!a->mn_major || (a->n_split <= 1 && a->tx == nullptr && a->k_block_ranges == nullptr && a->c != nullptr &&
                                     a->out_image == nullptr && a->out_image_t == nullptr && a->column_sums == nullptr &&
                                     a->bias == nullptr && a->aux == nullptr && a->aux_image == nullptr &&
                                     a->activation == TFEPB_ACT_NONE),
                    "mn_major: the raw product of two row images is ADDED to c (zero-filled by the caller), nothing else");
    TFEPB_CHECK_ARG(a->tile_list == nullptr || (a->mn_major && a->n_tile_list >= 0), "tile_list goes with mn_major");
    TFEPB_CHECK_ARG(!a->c_accumulate || (a->aux == nullptr && a->tx == nullptr),
                    "c_accumulate shares the staging buffer of aux and of the fused transformer (use aux_image)");
    if (int rc = require_sm100()) return rc;
    tcg::Params p{};
    p.a_img = (const uint8_t*)a->a_image; p.b_img = (const uint8_t*)a->b_image;
    p.M = a->m; p.N = a->n; p.K = a->k; p.k_blocks = (a->k + tcg::KB - 1) / tcg::KB;
    p.C = (float*)a->c; p.ldc = a->ldc; p.bias = (const float*)a->bias; p.act = a->activation;
    p.aux = (const float*)a->aux; p.ldaux = a->ldaux;
    p.aux_img = (const uint8_t*)a->aux_image; p.aux_k_blocks = (a->n + tcg::KB - 1) / tcg::KB;
    p.out_img = (uint8_t*)a->out_image; p.out_k_blocks = (a->n + tcg::KB - 1) / tcg::KB;
    p.out_img_t = (uint8_t*)a->out_image_t; p.t_rows = a->out_image_t != nullptr ? a->out_image_t_rows : 128;
    p.t_k_blocks = (a->m + tcg::KB - 1) / tcg::KB; p.t_rows_padded = (a->n + p.t_rows - 1) / p.t_rows * p.t_rows;
    p.colsum = a->column_sums;
    p.kranges = a->k_block_ranges;
    p.row_ranges = a->row_ranges;
    p.tiles_m = (a->m + tcg::BM - 1) / tcg::BM; p.tiles_n = (a->n + tcg::BN - 1) / tcg::BN;
    p.error = a->error_flag;
    if (tx != nullptr) {
        p.tx_units = tx->n_units; p.tx_unit_sphere = tx->unit_sphere; p.tx_max_radius = (float)tx->max_radius;
        p.tx_cols = tx->cols;
        p.tx_x = (const float*)tx->x; p.tx_ldx = tx->ldx;
        p.tx_y = (float*)tx->y; p.tx_ldy = tx->ldy;
        p.tx_logdet = tx->logdet;
        p.tx_gy = (const float*)tx->grad_y; p.tx_ldgy = tx->ldgy;
        p.tx_gl = tx->grad_logdet;
        p.tx_gx = (float*)tx->grad_x; p.tx_ldgx = tx->ldgx;
        p.sp_table = tx->spline_table;
    }
    int splits = a->split_k > 1 ? a->split_k : 1;
    if (splits > p.k_blocks) splits = p.k_blocks;
    p.atomic = splits > 1 ? 1 : 0;
    p.c_add = (a->c_accumulate != 0 && splits <= 1 && a->c != nullptr) ? 1 : 0;
    p.mn_major = a->mn_major != 0 ? 1 : 0;
    p.a_kblocks = (a->m + tcg::KB - 1) / tcg::KB;        // row images: a_image holds (k x m), b_image (k x n)
    p.b_kblocks = (a->n + tcg::KB - 1) / tcg::KB;
    if (p.mn_major) {
        // dedicated kernel: the reduction is cut in blocks of 128 image rows; always accumulates into C with atomics
        const int row_blocks = (a->k + 127) / 128;
        int sp = a->split_k > 1 ? a->split_k : 1;
        if (sp > row_blocks) sp = row_blocks;
        p.k_chunk_blocks = (row_blocks + sp - 1) / sp;
        sp = (row_blocks + p.k_chunk_blocks - 1) / p.k_chunk_blocks;
        p.atomic = 1;
        const size_t wg_smem = (size_t)tcg::WG_STAGES * tcg::WG_STAGE + sizeof(tcg::Smem) + 256;
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(tcg::tc_wgrad_kernel), wg_smem)) return rc;
        p.tile_list = a->tile_list; p.n_list = a->n_tile_list;
        const int wg_tiles = a->tile_list != nullptr ? a->n_tile_list : p.tiles_m * p.tiles_n;
        if (wg_tiles == 0) return 0;                         // everything masked: c stays as it is
        int wgx = sm_count() / sp;
        if (wgx < 1) wgx = 1;
        if (wgx > wg_tiles) wgx = wg_tiles;
        tcg::tc_wgrad_kernel<<<dim3((unsigned)wgx, (unsigned)sp), tcg::THREADS, wg_smem, as_stream(stream)>>>(p);
        return check_launch("tc_wgrad_kernel");
    }
    p.k_chunk_blocks = splits > 1 ? (p.k_blocks + splits - 1) / splits : 0;
    if (splits > 1) splits = (p.k_blocks + p.k_chunk_blocks - 1) / p.k_chunk_blocks;
    const int n_split = a->n_split > 1 ? a->n_split : 1;
    TFEPB_CHECK_ARG(n_split <= 3, "n_split must be 1, 2 or 3");
    TFEPB_CHECK_ARG(n_split == 1 || (a->out_image_t == nullptr && a->split_k <= 1),
                    "split-precision products are forward products: no transposed image, no split-K");
    p.a_split_stride = n_split > 1 ? tfepb_tc_image_bytes(a->m, a->k, 128) : 0;
    p.b_split_stride = n_split > 1 ? tfepb_tc_image_bytes(a->n, a->k, 256) : 0;
    p.out_split_stride = n_split > 1 ? tfepb_tc_image_bytes(a->m, a->n, 128) : 0;
    const int stage_bytes = n_split == 1 ? tcg::Geo<1>::STAGE : n_split == 2 ? tcg::Geo<2>::STAGE : tcg::Geo<3>::STAGE;
    const int stages = n_split == 3 ? tcg::Geo<3>::N_STAGES : tcg::Geo<1>::N_STAGES;
    const size_t smem = (size_t)stages * stage_bytes + (size_t)tcg::EPI_WARPS * (tcg::XP_FLOATS + 64) * 4 +
                        (tx != nullptr ? (size_t)tcg::TX_COLS_SMEM * 4 : 0) + (size_t)tcg::KR_SMEM * 4 + sizeof(tcg::Smem) + 256;
    using kernel_t = void (*)(const tcg::Params);
    kernel_t kernel = n_split == 1 ? tcg::tc_gemm_kernel<1> : n_split == 2 ? tcg::tc_gemm_kernel<2> : tcg::tc_gemm_kernel<3>;
    if (tx != nullptr) {
        static const kernel_t fused[8] = {
            tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_AFFINE>,   tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_AFFINE + 1>,
            tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_SOS2>,     tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_SOS2 + 1>,
            tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_MOEBIUS3>, tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_MOEBIUS3 + 1>,
            tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_SPLINE8>,  tcg::tc_gemm_kernel<1, 2 * TFEPB_TCTX_SPLINE8 + 1>};
        kernel = fused[2 * (tx->kind - 1) + (tx->backward ? 1 : 0)];
    }
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem)) return rc;
    const int tiles = p.tiles_m * p.tiles_n;
    // clusters of two CTAs sharing the B operand (opt-in, a->cluster): plain bf16 products without split-K / row ranges (the
    // pairs walk the tiles in lock step)
    p.cluster = (a->cluster != 0 && n_split == 1 && splits == 1 && a->row_ranges == nullptr && p.tiles_m >= 2 && sm_count() >= 2) ? 1 : 0;
    if (p.cluster) {
        const int pairs = ((p.tiles_m + 1) / 2) * p.tiles_n;
        int gx = sm_count() / 2;
        if (gx > pairs) gx = pairs;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(2 * gx)); cfg.blockDim = dim3((unsigned)tcg::THREADS);
        cfg.dynamicSmemBytes = smem; cfg.stream = as_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        void* args[] = {&p};
        TFEPB_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), args));
        return check_launch("tc_gemm_kernel (clusters)");
    }
    int gx = sm_count() / splits;
    if (gx < 1) gx = 1;
    if (gx > tiles) gx = tiles;
    dim3 grid((unsigned)gx, (unsigned)splits);
    kernel<<<grid, tcg::THREADS, smem, as_stream(stream)>>>(p);
    return check_launch("tc_gemm_kernel");
}
