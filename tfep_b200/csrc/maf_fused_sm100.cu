// Fused MAF forward for sm_100a: MADE conditioner on tcgen05 tensor cores + neural-spline transformer and
// per-sample log|det J| in the GEMM epilogue, for a chain of MAF layers in one persistent launch.
//
//   y, logdet = T_spline(x ; MADE(x))  per layer    (reference: nn/flows/autoregressive.py:144-177,
//                                                     nn/conditioners/made.py:294-329, nn/transformers/spline.py,
//                                                     chained as nn/flows/sequential.py:50-68)
//
// One persistent CTA per SM walks over WORK ITEMS (layer, tile of 128 samples), layer-major: item j belongs
// to CTA j % grid.  With L layers there are L x n_tiles items, so the last wave of the grid is nearly full
// (2048 items on 148 SMs instead of 512 per launch), and the whole chain is one launch.  A tile of layer
// l + 1 is read back from y (L2-resident) once the CTA that produced it published a per-tile flag; items
// are taken in increasing order by co-resident CTAs, so the wait always ends.
//
// For one item:
//   x tile (fp32)  --bulk copy-->  smem  --bf16, tcgen05.st-->  A0 in TENSOR MEMORY
//   GEMM1  D1 = A0 W1^T   (tcgen05.mma, A from TMEM, B from smem, D in TMEM)  -> epilogue: ELU, bf16 -> A1 (TMEM)
//   GEMM2  D2 = A1 W2^T                                                         -> epilogue: ELU, bf16 -> A2 (TMEM)
//   GEMM3  chunk c: 4 features x 28 columns (25 spline parameters each), THREE accumulator buffers in flight;
//          epilogue: softmax / softplus / bin search / rational-quadratic map + log-det straight out of TMEM,
//          so neither the hidden activations nor the (batch, 1650) parameter tensor ever exist in memory.
//   y tile is written in place over the x tile in smem; a dedicated warp stores it with one bulk copy, adds
//   the log-det of the layer to the running sum and publishes the tile flag.
// Keeping the A operand in tensor memory matters: with both operands in shared memory the MMAs ran at half
// rate (measured, profiles/r01_*), because A and B compete for the shared-memory operand port.
//
// Hidden units are degree-sorted and the output layer is packed feature-major (tfep_b200/_pack.py), so the
// autoregressive masks are block lower-triangular: the host-built schedule only lists the weight blocks
// that contain non-zeros -- masked blocks are neither fetched nor multiplied.  Weight blocks are stored in
// global memory as exact images of their shared-memory layout and streamed by the bulk-copy (TMA) engine
// through an mbarrier ring.
//
// Biases ride in the GEMMs: every A operand carries two constant-one columns and the weight blocks hold
// bf16(b) and bf16(b - bf16(b)) in the matching columns, so the epilogues never load a bias.  Hidden-layer
// rows and the output rows that feed softmax / softplus are pre-multiplied by log2(e), so ELU, softmax and
// softplus use the hardware ex2 / lg2 directly (the hidden activations travel as log2(e) ELU(h); the next
// layer's weights absorb the factor).
//
// Warp roles: warp 0 = TMEM allocator + bulk-copy producer, warps 1 and 3 = MMA issuers, warp 2 = y store /
// log-det / publish, warps 4..19 = epilogue (four warpgroups of 128 threads: thread <-> sample row,
// TMEM lane quadrant = warp % 4; the warpgroups split the columns of a hidden layer / the feature slots of
// a chunk).
#define TFEPB_INLINE_SLOW_WAIT 1
#include "common.cuh"
#include "tc_ptx.cuh"

#include <string.h>

namespace tfepb {
namespace fused {

using namespace tc;

constexpr int EPI_WGS = 4;                      // epilogue warpgroups
constexpr int EPI_THREADS = EPI_WGS * 128;
constexpr int EPI_WARPS = EPI_THREADS / 32;     // hand-over barriers count one arrival per epilogue WARP (lane 0, after a warp sync)
constexpr int AUX_THREADS = 128;                // producer, MMA issuer, store warp, one idle warp
constexpr int THREADS = AUX_THREADS + EPI_THREADS;
constexpr int AUX_REGS = 32, EPI_REGS = 112;    // setmaxnreg budgets: the CTA starts with 96 per thread, 128 x (96 - 32) = 512 x (112 - 96)
constexpr int STAGES = 3;
constexpr int STAGE_BYTES = 49152;              // one weight block (<= 256 rows, bf16)
constexpr int FEATS_PER_CHUNK = 4;
constexpr int PSTRIDE = 28;                     // accumulator columns per feature slot (25 used)
constexpr int CHUNK_N = FEATS_PER_CHUNK * PSTRIDE;   // 112
constexpr int ACC_BUFS = 3;                     // chunk accumulators in flight (3 x 112 = 336 columns)
constexpr int ACC_COLS = 336;                   // TMEM columns [0, 336): accumulators
constexpr int A_COL = 336;                      // TMEM columns [336, 512): bf16 A operand (2 k-values per column)

constexpr int MAX_LAYERS = TFEPB_FUSED_MAX_LAYERS;
constexpr int MAX_OPS = TFEPB_FUSED_MAX_OPS;
struct __align__(16) Op {   // one weight block = one ring stage (16 bytes; the table travels in the kernel parameters,
                            // so that the issuing warps read it through the constant bank into uniform registers)
    uint32_t w_off;         // byte offset into the packed weights of the layer (multiple of 16)
    uint32_t idesc;         // tcgen05 instruction descriptor (kind::f16, bf16 x bf16 -> fp32, M = 128, N = n)
    uint16_t n;             // MMA N of this block (rows of the weight block); the block holds n * ksteps * 32 bytes
    uint16_t tmem_col;      // destination accumulator column
    uint16_t a_col;         // first A column (two k-values each) this block multiplies
    uint8_t ksteps;         // K = 16 steps in this block
    uint8_t flags;
};
enum : uint32_t {
    OP_FIRST = 1u,          // first block of its accumulator: overwrite instead of accumulate
    OP_COMMIT = 2u,         // last block of its accumulator group: commit to acc_full[acc]
    OP_ACC_SHIFT = 2u,      // bits 2-3: accumulator buffer 0..2
    OP_WAIT_A = 16u,        // wait for the A operand (start of a GEMM phase)
    OP_WAIT_EMPTY = 32u,    // wait until the epilogue drained accumulator `acc` (GEMM3 chunks)
    OP_OWNER1 = 64u,        // issued by the second MMA warp (accumulator groups alternate between the two issuers)
    OP_HIDDEN = 128u,       // hidden-layer block: its group commits to hid_full[acc] (acc = column half) instead of acc_full[acc]
};

struct __align__(16) FeatConst {   // per sorted feature, as the epilogue reads it from shared memory (built from
                                   // tfepb_fused_feature at kernel start: the reciprocals are computed once per CTA)
    int col_kind;           // column in x / y, -1 = padding; bit 24: not circular (9 slopes, linear tails)
    float x0, L, invL;
    float iRw, iRh, Rh, y0; // Rw = L - 8 min_bin, Rh = (yf - y0) - 8 min_bin; iRw = 1 / Rw, iRh = 1 / Rh (adjacent: one packed multiply)
};
using FeatIn = tfepb_fused_feature;

struct LayerP {
    const uint8_t* weights;     // packed bf16 weight blocks
    const FeatIn* feats;        // n_chunks * FEATS_PER_CHUNK
    int op_base, n_ops, n_chunks;
    float min_bin, min_slope, slope_offset2;   // slope_offset2 = log2(e) * log(exp(1 - min_slope) - 1)
    const int* input_map;       // NULL, or per conditioner input column: x column | what enters << 16 (tfepb_fused_layer)
    float emb_lower, emb_scale;
    int hsplit[2];              // first column of the second half, per hidden layer (multiple of 16)
};

struct Params {
    const float* x; float* y; float* logdet;
    int batch, D;               // D = row length of x / y
    int K1;                     // D + 2 padded to a multiple of 16
    int HP;                     // hidden width (+2) padded to a multiple of 16
    int n_layers, n_tiles, feat_stride;   // feat_stride: feature slots reserved per layer in shared memory
    int n_halves;               // column halves of a hidden layer (1 or 2), each its own accumulator group and hand-over
    int x_col;                  // first column of the x operand inside the A operand region
    uint32_t* flags;            // (n_layers - 1) x n_tiles: == epoch once the tile of that layer is in y
    uint32_t epoch;
    int* error;                 // device int: set on watchdog timeout
    float* debug_params;        // optional (batch, n_chunks * CHUNK_N): conditioner outputs as seen by the epilogue
    int debug_mode;             // development only (timing experiments)
    LayerP layers[MAX_LAYERS];
    Op ops[MAX_OPS];
};

// ------------------------------------------------------------------------------------------------
// shared memory plan
// ------------------------------------------------------------------------------------------------
struct Smem {
    uint64_t w_full[STAGES], w_empty[STAGES];
    uint64_t x_full[2], x_empty[2], y_ready[2], a_ready;
    uint64_t acc_full[ACC_BUFS], acc_empty[ACC_BUFS], hid_full[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

// Spline epilogue for ONE feature of one sample.  wh[0..7] / wh[8..15] = width / height logits, sl[0..8] = slope
// logits and, for a circular feature, the shift in sl[8] (conditioner outputs, bias included by the GEMM; all but the
// shift pre-multiplied by log2(e)).  K = 8 bins (reference nn/transformers/spline.py:184-241, 319-417, 424-501, 546-650).
//   circular (kind 0): 8 slopes + the shift (not pre-multiplied); after the wrap the far tails cannot be
//     reached, and x exactly on the first knot gives the same value through the first bin;
//   not circular (kind 1, only in MIXED instantiations): 9 slopes; outside [x0, xf] the map is the linear
//     continuation with the boundary slope (the reference's far-tail bins are exactly linear).
// The walk over the knots runs in the UNNORMALISED domain of the softmax numerators: with a = sw / Rw the true knot
// X_k = (sum_{i<k} ew_i) Rw / sw + k min_bin satisfies a X_k = cw_{k-1} + k (a min_bin), so the prefix sums that give
// the softmax denominators also give the knots, neither the eight widths nor the eight heights are ever normalised
// (only the selected bin is), and 1 / sw is not needed at all: 24 MUFU operations per feature (16 ex2 of the two
// softmaxes, 2 x (ex2 + lg2) for the two softplus, rcp of the bin width, of sh and of the denominator, lg2 of dy/dx).
// The logits enter ex2 WITHOUT subtracting their maximum (softmax is shift invariant and fp32 spans 2^+-126): a
// feature whose sums leave [2^-100, 2^100] -- conditioner outputs beyond +-69 -- is recomputed with the maximum
// subtracted (warp-uniform branch, never taken for sane weights).
// Pipelining: `wh` is dead once the exponentials are taken; `handover()` (wait for the slopes, release the
// accumulator, wait for the next one) runs there, and the next chunk's logits are requested into `wh` (next_addr)
// after the knot walk, when the register pressure has dropped.
// Packed arithmetic: the width side and the height side do the same operations on independent values, so the prefix sums, the
// scale factors and the knots are computed as (width, height) pairs with FADD2 / FMUL2 / FFMA2 (17 issue slots fewer per
// feature, identical values; the knot index k enters FFMA2 as a broadcast immediate).
template <bool MIXED, class Handover>
__device__ __forceinline__ float spline8(uint32_t (&wh)[16], uint32_t (&sl)[9], float x, const FeatConst& fc,
                                         float min_bin, float min_slope, float slope_offset2, float& y,
                                         Handover&& handover, uint32_t next_addr, bool has_next) {
    const bool circ = !MIXED || (fc.col_kind >> 24) == 0;
    float ew[8], eh[8];
    f32x2 c2[8];                                    // (cw[k], ch[k])
    auto sums = [&](float mw, float mh) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            ew[k] = ex2(__uint_as_float(wh[k]) - mw);
            eh[k] = ex2(__uint_as_float(wh[8 + k]) - mh);
        }
        c2[0] = pack2(ew[0], eh[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) c2[k] = add2(c2[k - 1], pack2(ew[k], eh[k]));
    };
    sums(0.f, 0.f);
    float sw, sh;
    unpack2(c2[7], sw, sh);
    {
        const float lo = fminf(sw, sh), hi = fmaxf(sw, sh);
        if (__any_sync(0xffffffffu, !(lo > 7.9e-31f && hi < 1.3e30f))) {
            float mw = __uint_as_float(wh[0]), mh = __uint_as_float(wh[8]);
#pragma unroll
            for (int k = 1; k < 8; ++k) { mw = fmaxf(mw, __uint_as_float(wh[k])); mh = fmaxf(mh, __uint_as_float(wh[8 + k])); }
            sums(mw, mh);
            unpack2(c2[7], sw, sh);
        }
    }
    handover();
    float t = x - fc.x0;
    if (circ) {
        // wrap: (x - x0 + shift) mod L, result in [0, L)
        t += __uint_as_float(sl[8]);
        t = t - fc.L * floorf(t * fc.invL);
        t = (t < 0.f) ? t + fc.L : t;
        t = (t >= fc.L) ? t - fc.L : t;
    }
    const f32x2 ab = mul2(c2[7], pack2(fc.iRw, fc.iRh));       // a = sw / Rw, b = sh / Rh
    const f32x2 mb2 = mul2(pack2(min_bin, min_bin), ab);       // (min_bin a, min_bin b)
    float a, b;
    unpack2(ab, a, b);
    const float ts = t * a;
    // last knot below t, knots scaled by a (x side) / b (y side)
    float Us = 0.f, Vs = 0.f, ews = ew[0], ehs = eh[0];
    float raw0 = __uint_as_float(sl[0]), raw1 = __uint_as_float(sl[1]);
    const float raw_last = (MIXED && !circ) ? __uint_as_float(sl[8]) : __uint_as_float(sl[0]);   // tied to knot 0 if circular
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        float Uk, Vk;
        unpack2(fma2(pack2((float)k, (float)k), mb2, c2[k - 1]), Uk, Vk);
        const bool adv = ts > Uk;
        Us = adv ? Uk : Us;
        Vs = adv ? Vk : Vs;
        ews = adv ? ew[k] : ews;
        ehs = adv ? eh[k] : ehs;
        raw0 = adv ? __uint_as_float(sl[k]) : raw0;
        raw1 = adv ? (k == 7 ? raw_last : __uint_as_float(sl[k + 1])) : raw1;
    }
    if (has_next) tmem_ld16(next_addr, wh);      // in flight during the softplus / rational-quadratic tail
    const float dk = softplus_l2(raw0 + slope_offset2) + min_slope;
    const float dk1 = softplus_l2(raw1 + slope_offset2) + min_slope;
    float wsum, hsum;
    unpack2(add2(pack2(ews, ehs), mb2), wsum, hsum);   // a w, b h of the selected bin
    const float iw = rcp(wsum);                      // 1 / (a w)
    const float g = fc.Rh * rcp(sh);                 // 1 / b
    const float e = (ts - Us) * iw;
    const float h_sel = hsum * g;
    const float s = h_sel * (iw * a);                // h / w
    const float ome = 1.f - e, u = e * ome, e2 = e * e;
    const float q = dk1 + dk - 2.f * s;
    const float den = fmaf(q, u, s);
    const float iden = rcp(den);
    y = fmaf(Vs, g, fc.y0) + h_sel * fmaf(s, e2, dk * u) * iden;
    const float nn = fmaf(dk1, e2, fmaf(2.f * s, u, dk * ome * ome));
    const float rr = s * iden;
    float ld = LN2 * lg2(nn * rr * rr);
    if (MIXED && !circ) {
        // linear tails: below x0 the first bin is selected (dk = slope at knot 0), above xf the last one (dk1 = slope at knot 8)
        const bool lo = t < 0.f, hi = t > fc.L;
        const float yt = lo ? fmaf(dk, t, fc.y0) : fmaf(dk1, t - fc.L, fc.y0 + fmaf(8.f, min_bin, fc.Rh));   // yf = y0 + Rh + 8 min_bin
        const float lt = LN2 * lg2(lo ? dk : dk1);
        y = (lo || hi) ? yt : y;
        ld = (lo || hi) ? lt : ld;
    }
    return ld;
}

// development aid: CTA 0 stamps clock64() of key events into debug_params (as long long) when debug_mode & 16
template <bool DEBUG>
__device__ __forceinline__ void trace(const Params& p, int role, int& slot, int tag) {
#ifndef TFEPB_TRACE
    return;                 // compiled in only with -DTFEPB_TRACE (TFEPB_EXTRA_NVCC_FLAGS, scripts/dev_trace.py)
#endif
    if (!DEBUG) return;
    if ((p.debug_mode & 16) && blockIdx.x == 0 && slot < 400) {
        long long* t = reinterpret_cast<long long*>(p.debug_params) + role * 800 + 2 * slot;
        t[0] = clock64();
        t[1] = tag;
        ++slot;
    }
}

template <bool DEBUG, bool MIXED>
__global__ void __launch_bounds__(THREADS, 1) maf_spline_fwd_kernel(const __grid_constant__ Params p) {
    const int dmode = DEBUG ? p.debug_mode : 0;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sW = smem_raw;                                                     // weight ring
    float* sX0 = reinterpret_cast<float*>(sW + (size_t)STAGES * STAGE_BYTES);   // two x / y tiles, row-major [128][D]
    const int x_tile_bytes = TILE_M * p.D * 4;
    const int x_tile_stride = (x_tile_bytes + 127) & ~127;
    FeatConst* sFeat = reinterpret_cast<FeatConst*>(reinterpret_cast<uint8_t*>(sX0) + 2 * x_tile_stride);
    float* sLd = reinterpret_cast<float*>(sFeat + p.n_layers * p.feat_stride);   // [2][EPI_WGS][128] log-det partials
    Smem* sm = reinterpret_cast<Smem*>(sLd + 2 * EPI_WGS * TILE_M);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = p.n_tiles;
    const int n_items = p.n_layers * n_tiles;

    // ---- one-time setup ----
    for (int l = 0; l < p.n_layers; ++l) {
        const int nf = p.layers[l].n_chunks * FEATS_PER_CHUNK;
        for (int i = tid; i < nf; i += THREADS) {
            const FeatIn f = p.layers[l].feats[i];
            FeatConst c;
            c.col_kind = f.col < 0 ? -1 : (f.col | (f.kind << 24));
            c.x0 = f.x0; c.L = f.period; c.invL = f.inv_period;
            c.iRw = 1.f / f.rescaled_width; c.Rh = f.rescaled_height; c.iRh = 1.f / f.rescaled_height; c.y0 = f.y0;
            sFeat[l * p.feat_stride + i] = c;
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm->w_full[s], 1); mbar_init(&sm->w_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm->x_full[b], 1); mbar_init(&sm->x_empty[b], 1); mbar_init(&sm->y_ready[b], EPI_WARPS);
        }
        mbar_init(&sm->a_ready, EPI_WARPS);
        for (int b = 0; b < ACC_BUFS; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], EPI_WARPS); }
        mbar_init(&sm->hid_full[0], 1); mbar_init(&sm->hid_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    // Register budget per role (setmaxnreg moves registers between warpgroups): the four auxiliary warps need few,
    // the epilogue warpgroups take what they give up (the CTA starts with 640 x 96; 128 x 32 + 512 x 112 is the same total).
    // (One instruction per warpgroup, dominating the code it governs, so that the register allocator sees the budget.)
    if (warp < AUX_THREADS / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
    if (warp == 0) {
        // =========================== producer: x tiles + weight blocks ===========================
        // The whole warp walks the schedule (warp-uniform control flow); one elected lane issues the copies.
        uint32_t stage = 0, wphase = 0, tcount = 0;
        int ts = lane == 0 ? 0 : 1000000;
        // x tile of local item number t lives in buffer t & 1; it is requested one item ahead, so its latency
        // (and the drain of the y store that previously used the buffer) is off the critical path
        auto flag_of = [&](int item) -> const uint32_t* {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            return layer == 0 ? nullptr : p.flags + (size_t)(layer - 1) * n_tiles + tile;
        };
        auto request_x = [&](int item, uint32_t t) {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            const int rows = min(TILE_M, p.batch - tile * TILE_M);
            const uint32_t b = t & 1;
            const float* src = layer == 0 ? p.x : p.y;
            if (layer > 0) {
                // the tile must have been published by the CTA that ran the previous layer on it
                const uint32_t* flag = flag_of(item);
                uint32_t polls = 0;
                long long t0 = 0;
                while (ld_acquire_gpu(flag) != p.epoch) {
                    if ((++polls & 0xffu) != 0) continue;
                    if (t0 == 0) t0 = clock64();
                    if ((uint64_t)(clock64() - t0) > WATCHDOG_CYCLES) {
                        if (p.error) atomicExch(p.error, 9);
                        __threadfence_system();
                        __trap();
                    }
                }
                fence_proxy_async();        // the copy below reads through the async proxy
            }
            mbar_wait(&sm->x_empty[b], ((t >> 1) & 1) ^ 1, p.error, 1);
            if (elect_one()) {
                float* dst = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sX0) + b * x_tile_stride);
                if (rows == TILE_M) {
                    mbar_expect_tx(&sm->x_full[b], (uint32_t)x_tile_bytes);
                    bulk_g2s(dst, src + (size_t)tile * TILE_M * p.D, (uint32_t)x_tile_bytes, &sm->x_full[b]);
                } else {
                    mbar_arrive(&sm->x_full[b]);    // ragged last tile: the epilogue warps copy it themselves
                }
            }
            __syncwarp();
        };
        // When every layer has at least two rounds of tiles, the tile an item needs was published a full round
        // ago: its x tile is requested early, during the previous item.  Otherwise the producer of the needed tile
        // may be this very CTA (or a neighbour) still working on it: the request is deferred into the item's own
        // schedule and polled without blocking while the first weight blocks stream in.
        const bool far = n_tiles >= 2 * (int)gridDim.x;
        bool need_x = true;
        int layer = 0, tile = blockIdx.x;            // grid <= n_tiles: the first item of every CTA is in layer 0
        while (layer < p.n_layers) {
            const int item = layer * n_tiles + tile;
            const int n_ops = p.layers[layer].n_ops, op_base = p.layers[layer].op_base;
            const uint8_t* weights = p.layers[layer].weights;
            trace<DEBUG>(p, 0, ts, 1001);
            const int prefetch_at = n_ops > 8 ? 8 : n_ops - 1;   // after the first weight blocks of this item are in flight
            const int must_at = n_ops > STAGES - 1 ? STAGES - 1 : n_ops - 1;   // the ring is full: the MMAs need x to go on
            for (int i = 0; i < n_ops; ++i) {
                if (need_x) {
                    const uint32_t* flag = flag_of(item);
                    if (i == must_at || flag == nullptr || ld_acquire_gpu(flag) == p.epoch) {
                        request_x(item, tcount);
                        need_x = false;
                    }
                }
                const Op op = p.ops[op_base + i];
                const uint32_t bytes = (uint32_t)op.n * op.ksteps * 32u;
                mbar_wait(&sm->w_empty[stage], wphase ^ 1, p.error, 2);
                trace<DEBUG>(p, 0, ts, i);
                if (elect_one()) {
                    if (dmode & 1) {
                        mbar_arrive(&sm->w_full[stage]);
                    } else {
                        mbar_expect_tx(&sm->w_full[stage], bytes);
                        bulk_g2s(sW + (size_t)stage * STAGE_BYTES, weights + op.w_off, bytes, &sm->w_full[stage]);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; wphase ^= 1; }
                if (far && i == prefetch_at && item + (int)gridDim.x < n_items) request_x(item + gridDim.x, tcount + 1);
            }
            need_x = !far;
            ++tcount;
            tile += gridDim.x;
            while (tile >= n_tiles) { tile -= n_tiles; ++layer; }
        }
    } else if (warp == 1 || warp == 3) {
        // =========================== MMA issuers ===========================
        // Warp-uniform walk over the schedule; one elected lane issues the tcgen05.mma / commit instructions.
        // A single warp needs ~100 dependent instructions (~450 cycles) per weight block, more than the MMAs of
        // the block take, so the accumulator groups alternate between TWO issuer warps (OP_OWNER1): each walks
        // the whole schedule to keep the ring position, and issues only its own groups.  Accumulation order
        // inside a group is kept (one warp per group); groups are independent of each other.
        const uint32_t me = warp == 3 ? OP_OWNER1 : 0u;
        uint32_t stage = 0, wphase = 0, a_par = 0, empty_bits = 0;   // empty_bits: next phase parity per accumulator
        int ts = (lane == 0 && warp == 1) ? 0 : 1000000;
        const uint32_t w_base16 = smem_u32(sW) >> 4;
        int layer = 0, tile = blockIdx.x;
        while (layer < p.n_layers) {
            const int n_ops = p.layers[layer].n_ops, op_base = p.layers[layer].op_base;
            for (int i = 0; i < n_ops; ++i) {
                const Op op = p.ops[op_base + i];
                const uint32_t flags = op.flags, n = op.n;
                const uint32_t acc = (flags >> OP_ACC_SHIFT) & 3u;
                if (flags & OP_WAIT_A) {            // both issuers pass every A hand-over (keeps their parity in step)
                    mbar_wait(&sm->a_ready, a_par, p.error, 3);
                    a_par ^= 1u;
                }
                const uint32_t empty_par = ((empty_bits >> acc) & 1u) ^ 1u;
                if (flags & OP_WAIT_EMPTY) empty_bits ^= 1u << acc;      // every use of the buffer, whoever issues it
                if ((flags & OP_OWNER1) != me) {
                    if (++stage == STAGES) { stage = 0; wphase ^= 1; }
                    continue;
                }
                if (flags & OP_WAIT_EMPTY) mbar_wait(&sm->acc_empty[acc], empty_par, p.error, 4);
                trace<DEBUG>(p, 1, ts, 2000 + i);
                mbar_wait(&sm->w_full[stage], wphase, p.error, 5);
                tc_fence_after();
                trace<DEBUG>(p, 1, ts, i);
                // B descriptor: high word constant (SBO = 128 B, version 1); low word = address >> 4 | LBO >> 4 << 16,
                // advanced per k-step by two 8-k slabs of n rows x 16 B.  A: 8 TMEM columns per k-step.
                const uint32_t b_lo = w_base16 + stage * (STAGE_BYTES >> 4) + (n << 16);
                const uint32_t a_tmem = tmem + A_COL + op.a_col;
                const uint32_t ksteps = (dmode & 8) ? 1u : (uint32_t)op.ksteps;
                const uint32_t d_tmem = tmem + op.tmem_col;
                if (elect_one()) {
                    uint32_t accumulate = (flags & OP_FIRST) ? 0u : 1u;
                    for (uint32_t ks = 0; ks < ksteps; ++ks) {
                        const uint64_t db = ((uint64_t)DESC_HI << 32) | (b_lo + ks * 2u * n);
                        umma_ts(d_tmem, a_tmem + ks * 8u, db, op.idesc, accumulate);
                        accumulate = 1u;
                    }
                    umma_commit(&sm->w_empty[stage]);                   // frees the ring stage when the MMAs retire
                    if (flags & OP_COMMIT) umma_commit((flags & OP_HIDDEN) ? &sm->hid_full[acc] : &sm->acc_full[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; wphase ^= 1; }
            }
            tile += gridDim.x;
            while (tile >= n_tiles) { tile -= n_tiles; ++layer; }
        }
    } else if (warp == 2) {
        // =========================== y store, log-det, publish ===========================
        uint32_t tcount = 0;
        int ts = lane == 0 ? 0 : 1000000;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            const int rows = min(TILE_M, p.batch - tile * TILE_M);
            const uint32_t xb = tcount & 1;
            const float* sY = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(sX0) + xb * x_tile_stride);
            mbar_wait(&sm->y_ready[xb], (tcount >> 1) & 1, p.error, 10);
            trace<DEBUG>(p, 3, ts, 4000);
            float* dst = p.y + (size_t)tile * TILE_M * p.D;
            if (rows == TILE_M) {
                if (elect_one()) bulk_s2g(dst, sY, (uint32_t)x_tile_bytes);
                __syncwarp();
            }
            // log-det of this layer: fixed summation order over the warpgroups, added to the running sum
            const float* ldp = sLd + xb * (EPI_WGS * TILE_M);
#pragma unroll
            for (int r = 0; r < TILE_M / 32; ++r) {
                const int row = r * 32 + lane;
                if (row < rows) {
                    float v = ldp[row];
#pragma unroll
                    for (int g = 1; g < EPI_WGS; ++g) v += ldp[g * TILE_M + row];
                    float* out = p.logdet + (size_t)tile * TILE_M + row;
                    if (layer > 0) v += __ldcg(out);
                    *out = v;
                }
            }
            if (rows < TILE_M) {
                for (int i = lane; i < rows * p.D; i += 32) dst[i] = sY[i];
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm->x_empty[xb]);
            } else {
                if (lane == 0) {                 // the lane that issued the bulk store owns its group
                    bulk_wait_read();
                    mbar_arrive(&sm->x_empty[xb]);
                    if (layer + 1 < p.n_layers) bulk_wait_all();
                }
            }
            __syncwarp();
            if (layer + 1 < p.n_layers && lane == 0) {
                fence_proxy_async();
                __threadfence();
                st_release_gpu(p.flags + (size_t)layer * n_tiles + tile, p.epoch);
            }
            trace<DEBUG>(p, 3, ts, 4001);
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
        // =========================== epilogue warps ===========================
        const int et = tid - AUX_THREADS;         // 0..511
        const int wg = et >> 7;                   // warpgroup 0..3
        const int row = (warp & 3) * 32 + lane;   // sample row of the tile = TMEM lane
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t full_bits = 0, hid_par = 0, tcount = 0;   // full_bits: phase parity to wait for, per accumulator
        int ts = (et == 0) ? 0 : 1000000;

        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            const LayerP& L = p.layers[layer];
            const int rows = min(TILE_M, p.batch - tile * TILE_M);
            // ---- x tile -> A0 in tensor memory (bf16 pairs; columns D, D+1 are the constant ones) ----
            const uint32_t xb = tcount & 1;
            float* sX = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sX0) + xb * x_tile_stride);
            float* xrow = sX + row * p.D;
            mbar_wait(&sm->x_full[xb], (tcount >> 1) & 1, p.error, 6);
            trace<DEBUG>(p, 2, ts, 3001);
            if (rows < TILE_M) {
                const float* src = (layer == 0 ? p.x : p.y) + (size_t)tile * TILE_M * p.D;
                for (int i = et; i < TILE_M * p.D; i += EPI_THREADS) sX[i] = i < rows * p.D ? __ldcg(src + i) : 0.f;
                asm volatile("bar.sync 1, 512;" ::: "memory");
            }
            for (int c0 = wg * 8; c0 < p.K1 / 2; c0 += EPI_WGS * 8) {      // 8 TMEM columns = 16 inputs per step
                uint32_t q[8];
                if (L.input_map == nullptr) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = (c0 + i) * 2;
                        const float v0 = k < p.D ? xrow[k] : (k < p.D + 2 ? 1.f : 0.f);
                        const float v1 = k + 1 < p.D ? xrow[k + 1] : (k + 1 < p.D + 2 ? 1.f : 0.f);
                        q[i] = pack_bf16(v0, v1);
                    }
                } else {
                    // PeriodicEmbedding fused into the operand staging (reference mafembed.py:112-142): periodic
                    // features enter the conditioner as cos / sin of their angle
                    auto input = [&](int k) -> float {
                        const int e = __ldg(L.input_map + k);
                        const int what = e >> 16;
                        float v = what >= 3 ? (what == 3 ? 1.f : 0.f) : xrow[e & 0xffff];
                        if (what == 1 || what == 2) {
                            const float ang = (v - L.emb_lower) * L.emb_scale;
                            v = what == 1 ? __cosf(ang) : __sinf(ang);
                        }
                        return v;
                    };
#pragma unroll
                    for (int i = 0; i < 8; ++i) q[i] = pack_bf16(input((c0 + i) * 2), input((c0 + i) * 2 + 1));
                }
                tmem_st8(lane_addr + A_COL + p.x_col + c0, q);
            }
            tmem_st_wait();
            tc_fence_before();
            if (lane == 0) mbar_arrive(&sm->a_ready);
            trace<DEBUG>(p, 2, ts, 3002);
            // ---- two hidden layers: ELU, bf16 -> A operand of the next GEMM (tensor memory) ----
            // Each layer is handed over in two column halves: GEMM2 of the first half (its rows only see the
            // first half of h1) runs while the second half of h1 is still in the ELU, and the first output chunk
            // starts on the first half of h2.
            for (int hl = 0; hl < 2; ++hl) {
                for (int half = 0; half < p.n_halves; ++half) {
                    // The activations overwrite the A operand region.  h2 over h1: same columns, so the second layer
                    // waits for BOTH halves of GEMM2.  h1 over x: the planner puts x at the END of the region
                    // (p.x_col), clear of the first half of h1, so that half is written while the second half of
                    // GEMM1 still reads x; the second half of h1 (which may reach x) waits for its own GEMM1 half,
                    // issued after the first.  Without that placement the first layer waits for both halves too.
                    const bool both = p.n_halves == 2 && (hl == 1 || p.x_col < L.hsplit[0] / 2);
                    if (half == 0 || !both) {
                        mbar_wait(&sm->hid_full[half], (hid_par >> half) & 1u, p.error, 7);
                        hid_par ^= 1u << half;
                        if (half == 0 && both) {
                            mbar_wait(&sm->hid_full[1], (hid_par >> 1) & 1u, p.error, 7);
                            hid_par ^= 2u;
                        }
                    }
                    tc_fence_after();
                    trace<DEBUG>(p, 2, ts, 3010 + 2 * hl + half);
                    const int cbeg = half == 0 ? 0 : L.hsplit[hl];
                    const int cend = (half == 0 && p.n_halves == 2) ? L.hsplit[hl] : p.HP;
                    // 16 accumulator columns per step, two register sets in ping-pong: the load of the next step is
                    // in flight while this one is processed
                    uint32_t ra[16], rb[16];
                    float l2e;                          // opaque to the compiler: kept in ONE register instead of being
                    asm volatile("mov.f32 %0, 0f3FB8AA3B;" : "=f"(l2e));    // re-materialised for every element
                    const f32x2 l2e2 = pack2(l2e, l2e), ml2e2 = pack2(-l2e, -l2e);
                    auto elu16 = [&](uint32_t (&r)[16], int c0) {
                        uint32_t q[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (DEBUG && (dmode & 4)) {
                                q[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                            } else {
                                float v0, v1;
                                elu_l2x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]), l2e2, ml2e2, v0, v1);
                                q[i] = pack_bf16(v0, v1);
                            }
                        }
                        tmem_st8(lane_addr + A_COL + c0 / 2, q);
                    };
                    int c0 = cbeg + wg * 16;
                    if (c0 < cend) tmem_ld16(lane_addr + c0, ra);
                    while (c0 < cend) {
                        tmem_wait8(ra); tmem_wait8(ra + 8);
                        int cn = c0 + EPI_WGS * 16;
                        if (cn < cend) tmem_ld16(lane_addr + cn, rb);
                        elu16(ra, c0);
                        c0 = cn;
                        if (c0 >= cend) break;
                        tmem_wait8(rb); tmem_wait8(rb + 8);
                        cn = c0 + EPI_WGS * 16;
                        if (cn < cend) tmem_ld16(lane_addr + cn, ra);
                        elu16(rb, c0);
                        c0 = cn;
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(&sm->a_ready);
                    trace<DEBUG>(p, 2, ts, 3020 + 2 * hl + half);
                }
            }
            // ---- output layer chunks: spline transformer straight out of TMEM ----
            // Software pipeline over the chunks: the width / height logits of chunk c + 1 are requested from tensor
            // memory as soon as those of chunk c have gone through ex2 (they are dead then), so their latency hides
            // behind the knot walk of chunk c; the slopes of chunk c are requested right before its exponentials.
            float ld = 0.f;
            const float min_bin = L.min_bin, min_slope = L.min_slope, slope_offset2 = L.slope_offset2;
            const int n_chunks = L.n_chunks;
            // addresses of the loop, pinned in registers (the compiler would otherwise re-derive them from threadIdx and
            // the kernel parameters in every iteration): this thread's TMEM lane + feature slot, its row of the x / y tile,
            // the feature records of its slot, the accumulator barriers
            const uint32_t t_slot = pinned(lane_addr + wg * PSTRIDE);
            const uint32_t s_xrow = pinned(smem_u32(xrow));
            uint32_t s_feat = pinned(smem_u32(sFeat + layer * p.feat_stride + wg));     // one feature slot per warpgroup
            const uint32_t s_full = pinned(smem_u32(&sm->acc_full[0])), s_empty = pinned(smem_u32(&sm->acc_empty[0]));
            const uint32_t is_lane0 = pinned(lane == 0 ? 1u : 0u);
            uint32_t b = 0;
            uint32_t wh[16], sl[9];
            mbar_wait_s(s_full, full_bits & 1u, p.error, 8);
            full_bits ^= 1u;
            tc_fence_after();
            tmem_ld16(t_slot, wh);
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c, s_feat += FEATS_PER_CHUNK * (uint32_t)sizeof(FeatConst)) {
                trace<DEBUG>(p, 2, ts, 3100 + c);
                const uint32_t a0 = t_slot + b * CHUNK_N;
                tmem_wait8(wh); tmem_wait8(wh + 8);
                tmem_ld8(a0 + 16, sl);
                tmem_ld1(a0 + 24, sl + 8);
                FeatConst fc;
                {
                    uint32_t f0[4], f1[4];
                    lds_v4(s_feat, f0);
                    lds_v4(s_feat + 16, f1);
                    fc.col_kind = (int)f0[0]; fc.x0 = __uint_as_float(f0[1]); fc.L = __uint_as_float(f0[2]); fc.invL = __uint_as_float(f0[3]);
                    fc.iRw = __uint_as_float(f1[0]); fc.iRh = __uint_as_float(f1[1]); fc.Rh = __uint_as_float(f1[2]); fc.y0 = __uint_as_float(f1[3]);
                }
                const int col = fc.col_kind < 0 ? -1 : (fc.col_kind & 0xffffff);
                const uint32_t bn = (b == ACC_BUFS - 1) ? 0u : b + 1u;
                const bool has_next = c + 1 < n_chunks;
                // hand the accumulator back once the slopes are in registers (TMEM loads are warp-collective: when
                // lane 0 has its values, so has the warp), then wait for the next chunk's accumulator
                auto handover = [&]() {
                    tmem_wait8(sl); tmem_wait1(sl + 8);
                    tc_fence_before();
                    if (is_lane0) mbar_arrive_s(s_empty + b * 8u);
                    if (has_next) {
                        mbar_wait_s(s_full + bn * 8u, (full_bits >> bn) & 1u, p.error, 8);
                        full_bits ^= 1u << bn;
                        tc_fence_after();
                    }
                };
                if (DEBUG && p.debug_params != nullptr && !(dmode & 16) && row < rows && layer == 0 && col >= 0) {
                    float* dbg = p.debug_params + ((size_t)tile * TILE_M + row) * n_chunks * CHUNK_N + c * CHUNK_N + wg * PSTRIDE;
                    tmem_wait8(sl); tmem_wait1(sl + 8);
#pragma unroll
                    for (int i = 0; i < 16; ++i) dbg[i] = __uint_as_float(wh[i]);
#pragma unroll
                    for (int i = 0; i < 9; ++i) dbg[16 + i] = __uint_as_float(sl[i]);
                }
                if (DEBUG && (dmode & 2)) {                 // timing experiment: hand-over skeleton without the spline math
                    handover();
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 9; ++i) acc += __uint_as_float(sl[i]);
                    ld += acc;
                    if (has_next) tmem_ld16(t_slot + bn * CHUNK_N, wh);
                } else if (col >= 0) {
                    float yv;
                    const uint32_t s_x = s_xrow + (uint32_t)col * 4u;
                    ld += spline8<MIXED>(wh, sl, lds_f32(s_x), fc, min_bin, min_slope, slope_offset2, yv, handover,
                                         t_slot + bn * CHUNK_N, has_next);
                    sts_f32(s_x, yv);
                } else {
                    handover();
                    if (has_next) tmem_ld16(t_slot + bn * CHUNK_N, wh);
                }
                trace<DEBUG>(p, 2, ts, 3200 + c);
                b = bn;
            }
            // ---- hand the y tile and the log-det partials to the store warp ----
            sLd[xb * (EPI_WGS * TILE_M) + wg * TILE_M + row] = ld;
            fence_async_smem();                      // y tile writes -> visible to the bulk-copy engine
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->y_ready[xb]);
            trace<DEBUG>(p, 2, ts, 3300);
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

size_t smem_bytes(const Params& p) {
    size_t s = (size_t)STAGES * STAGE_BYTES;
    s += 2 * (((size_t)TILE_M * p.D * 4 + 127) & ~(size_t)127);
    s += (size_t)p.n_layers * p.feat_stride * sizeof(FeatConst);
    s += 2 * EPI_WGS * TILE_M * 4 + sizeof(Smem);
    return s + 256;
}

}  // namespace fused
}  // namespace tfepb

using namespace tfepb;

static_assert(sizeof(fused::Op) == sizeof(tfepb_fused_op), "schedule entry layout mismatch");
static_assert(sizeof(fused::FeatConst) == 32 && sizeof(tfepb_fused_feature) == 32, "feature table layout");
static_assert(sizeof(fused::Params) < 32000, "kernel parameter space");

extern "C" int tfepb_maf_spline_forward_bf16(const tfepb_fused_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->x && a->y && a->logdet && a->layers, "null buffer");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->n_features > 0, "bad sizes");
    TFEPB_CHECK_ARG(a->n_layers >= 1 && a->n_layers <= fused::MAX_LAYERS, "n_layers must be in [1, %d]", fused::MAX_LAYERS);
    TFEPB_CHECK_ARG(a->n_layers == 1 || a->tile_flags != nullptr, "a chain of layers needs the tile_flags workspace");
    const int n_inputs = a->n_inputs > 0 ? a->n_inputs : a->n_features;
    TFEPB_CHECK_ARG(n_inputs >= a->n_features && n_inputs <= 2 * a->n_features, "n_inputs must be in [n_features, 2 n_features]");
    TFEPB_CHECK_ARG(a->k1 % 16 == 0 && a->k1 >= n_inputs + 2, "k1 must hold n_inputs + 2 bias columns, rounded up to 16");
    TFEPB_CHECK_ARG(a->hidden_padded % 16 == 0 && a->hidden_padded > 0 && a->hidden_padded <= fused::ACC_COLS,
                    "hidden width (padded) must be a multiple of 16 and at most 336 (tensor-memory plan)");
    TFEPB_CHECK_ARG(a->k1 <= 2 * (512 - fused::A_COL), "too many input features for the tensor-memory plan");
    TFEPB_CHECK_ARG(a->hidden_halves == 1 || a->hidden_halves == 2, "hidden_halves must be 1 or 2");
    TFEPB_CHECK_ARG((a->n_features * 4 * fused::TILE_M) % 16 == 0, "tile of x must be a multiple of 16 bytes");
    TFEPB_CHECK_ARG(((uintptr_t)a->x % 16 == 0) && ((uintptr_t)a->y % 16 == 0), "x and y must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    fused::Params p{};
    p.x = (const float*)a->x; p.y = (float*)a->y; p.logdet = (float*)a->logdet;
    p.batch = a->batch; p.D = a->n_features; p.K1 = a->k1; p.HP = a->hidden_padded;
    p.n_layers = a->n_layers;
    p.n_tiles = (a->batch + fused::TILE_M - 1) / fused::TILE_M;
    p.flags = a->tile_flags; p.epoch = a->epoch;
    int op_base = 0, feat_stride = 0;
    for (int l = 0; l < a->n_layers; ++l) {
        const tfepb_fused_layer& s = a->layers[l];
        TFEPB_CHECK_ARG(s.ops && s.weights && s.feats, "layer %d: null buffer", l);
        TFEPB_CHECK_ARG(s.n_chunks > 0 && s.n_ops > 0 && op_base + s.n_ops <= fused::MAX_OPS, "layer %d: bad schedule length", l);
        TFEPB_CHECK_ARG((uintptr_t)s.weights % 16 == 0, "layer %d: packed weights must be 16-byte aligned", l);
        fused::LayerP& d = p.layers[l];
        d.weights = (const uint8_t*)s.weights;
        d.feats = s.feats;
        d.op_base = op_base; d.n_ops = s.n_ops; d.n_chunks = s.n_chunks;
        d.min_bin = s.min_bin_size; d.min_slope = s.min_slope; d.slope_offset2 = s.slope_offset * fused::LOG2E;
        TFEPB_CHECK_ARG(s.input_map != nullptr || n_inputs == a->n_features, "layer %d: n_inputs > n_features needs an input_map", l);
        d.input_map = s.input_map; d.emb_lower = s.emb_lower; d.emb_scale = s.emb_scale;
        for (int i = 0; i < 2; ++i) {        // per-layer split, else the chain-wide one
            d.hsplit[i] = s.hidden_split[i] > 0 ? s.hidden_split[i] : a->hidden_split[i];
            TFEPB_CHECK_ARG(a->hidden_halves == 1 || (d.hsplit[i] % 16 == 0 && d.hsplit[i] > 0 && d.hsplit[i] < a->hidden_padded),
                            "layer %d: hidden_split must be a multiple of 16 inside the hidden width", l);
        }
        memcpy(p.ops + op_base, s.ops, sizeof(fused::Op) * (size_t)s.n_ops);
        op_base += s.n_ops;
        if (s.n_chunks * fused::FEATS_PER_CHUNK > feat_stride) feat_stride = s.n_chunks * fused::FEATS_PER_CHUNK;
    }
    p.feat_stride = feat_stride;
    p.n_halves = a->hidden_halves;
    TFEPB_CHECK_ARG(a->x_operand_column >= 0 && a->x_operand_column % 8 == 0 &&
                    a->x_operand_column + a->k1 / 2 <= 512 - fused::A_COL, "x_operand_column outside the A operand region");
    p.x_col = a->x_operand_column;
    p.error = a->error_flag;
    p.debug_params = a->debug_params;
    p.debug_mode = a->debug_mode;
    const size_t smem = fused::smem_bytes(p);
    TFEPB_CHECK_ARG(smem <= 227 * 1024, "shared memory plan of %zu bytes exceeds 227 KB", smem);
    const bool debug = p.debug_mode != 0 || p.debug_params != nullptr;
    const bool mixed = a->mixed_splines != 0;      // some features are not circular: generic spline epilogue
    auto kernel = debug ? (mixed ? fused::maf_spline_fwd_kernel<true, true> : fused::maf_spline_fwd_kernel<true, false>)
                        : (mixed ? fused::maf_spline_fwd_kernel<false, true> : fused::maf_spline_fwd_kernel<false, false>);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem)) return rc;
    // One CTA per SM, never more CTAs than tiles and never more than can be co-resident: the items a CTA waits for
    // belong to CTAs that run (cooperative launch: the driver enforces it instead of a spin that ends in a trap).
    int grid = 0;
    if (int rc = coresident_grid(reinterpret_cast<const void*>(kernel), fused::THREADS, smem, p.n_tiles, &grid)) return rc;
    if (int rc = launch_cooperative(reinterpret_cast<const void*>(kernel), grid, fused::THREADS, smem, &p, as_stream(stream))) return rc;
    return check_launch("maf_spline_fwd_kernel");
}
