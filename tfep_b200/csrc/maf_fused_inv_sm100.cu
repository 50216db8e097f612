// Fused MAF INVERSE for sm_100a: the degree-ordered sweep of MAF.inverse on tcgen05 tensor cores, for a chain
// of MAF layers in one persistent launch (bf16 operands, fp32 accumulation and epilogue).
//
//   x, logdet = MAF.inverse(y)          (reference: nn/flows/autoregressive.py:179-229 -- n_degrees full
//                                         conditioner passes; chained in reverse as nn/flows/sequential.py:50-68)
//
// Inverting an autoregressive flow is sequential over the degrees: the spline parameters of the feature of
// degree d need the hidden units of degree < d, which need x of degree < d.  With degree-sorted hidden units
// every unit is computed once (tfep_b200/_pack.py), so per degree there are three small dependent products
//
//   OUT_d : par_d (25 rows, padded to 32) = W3[feature d] . h2[units of degree < d]     -> spline inverse -> x_d
//   H1_d  : h1[units of degree d] (<= 15, padded to 16) = ELU(W1[those rows] . x)        -> A1 operand
//   H2_d  : h2[units of degree d]                       = ELU(W2[those rows] . h1[<= d]) -> A2 operand
//
// each a handful of tcgen05.mma instructions (M = 128 samples, N = 32 / 16, K up to 336) whose A operands
// -- x, h1 and h2 of the 128-sample tile -- stay RESIDENT IN TENSOR MEMORY for the whole sweep and grow
// in place, column by column, as the epilogue warpgroup produces them.  The weight blocks (bias folded in,
// log2(e) pre-scalings as in the forward kernel, maf_fused_sm100.cu) stream from L2 through the bulk-copy
// ring in the order they are needed; the 3 x D steps of a layer form one dependent chain, so the kernel is
// latency bound and one CTA per SM runs one tile.  Work items are (layer, tile) pairs, layer-major, with
// per-tile flags between the layers exactly as in the forward kernel.
//
// Warp roles: warp 0 = TMEM allocator + bulk-copy producer, warp 1 = MMA issuer, warp 2 = x store / log-det /
// publish, warp 3 idle, warps 4..7 = epilogue (thread <-> sample row <-> TMEM lane).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <string.h>

namespace tfepb {
namespace finv {

using namespace tc;

constexpr int THREADS = 256;
constexpr int EPI_THREADS = 128;
constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 24576;              // largest block: 32 rows x 336 k x 2 B = 21 KB
constexpr int NPAR = 25;
constexpr int ACC_OUT = 0, ACC_HID = 32;        // accumulator columns
constexpr int A0_COL = 64, A1_COL = 128, A2_COL = 320;   // bf16 A operands: x (<= 64 cols), h1, h2 (<= 184 cols)
constexpr int MAX_LAYERS = TFEPB_FUSED_MAX_LAYERS;

struct __align__(16) Op {   // same layout as tfepb_fused_op; a_col = absolute TMEM column of the A operand
    uint32_t w_off, idesc;
    uint16_t n, tmem_col, a_col;
    uint8_t ksteps, flags;
};

struct __align__(16) Step {  // one degree of the sweep (tfepb_fused_inv_step)
    int col;                 // column of the feature in y / x
    float x0, L, invL, Rw, Rh, y0;
    int partner;             // bits 0-3: the other half of the bf16 pair column of the conditioner input: 0 unknown yet
                             // (zero), 1 already inverted (read it back), 2 the constant-one bias column; bit 4: spline
                             // kind (0 circular, 1 not circular: 9 slopes, linear tails); bit 5: enters the conditioner as
                             // (cos, sin) (PeriodicEmbedding); bits 8-15: conditioner input column; bits 16-23: x column
                             // of the pair partner
    int h1_a, h1_n, h2_a, h2_n;   // packed positions of the hidden units that become computable (n = 0: none)
};

struct LayerP {
    const uint8_t* weights;
    const Op* ops;           // device
    const Step* steps;       // device
    int n_ops, n_steps;
    float min_bin, min_slope, slope_offset2;
    float emb_lower, emb_scale;
    const int* init_map;     // NULL, or per conditioner input column what it holds BEFORE the sweep (tfepb_fused_inv_layer)
};

struct Params {
    const float* y; float* x; float* logdet;
    int batch, D, K1, HP, n_layers, n_tiles;
    int Din;                 // conditioner inputs before the two constant ones (D + lifted periodic features)
    uint32_t* flags; uint32_t epoch;
    int* error;
    LayerP layers[MAX_LAYERS];
};

struct Smem {
    uint64_t w_full[STAGES], w_empty[STAGES];
    uint64_t x_full[2], x_empty[2], y_ready[2], a_ready, acc_full;
    uint32_t tmem_base;
    uint32_t pad[3];
};

// Inverse of the 8-bin spline for one feature of one sample (reference nn/transformers/spline.py:504-543,
// 257-259).  r[0..24] as in the forward epilogue (log2-domain widths / heights / slopes; circular: the shift in
// r[24]; not circular (MIXED instantiations): the ninth slope, linear tails outside [y0, yf]).  Returns log|dx/dy|.
template <bool MIXED>
__device__ __forceinline__ float spline8_inverse(const uint32_t (&r)[32], float yv, const Step& fc, float min_bin,
                                                 float min_slope, float slope_offset2, float& x) {
    float p[NPAR];
#pragma unroll
    for (int i = 0; i < NPAR; ++i) p[i] = __uint_as_float(r[i]);
    const bool circ = !MIXED || (fc.partner & 16) == 0;
    const float Ly = fmaf(8.f, min_bin, fc.Rh);                  // yf - y0
    const float t0 = yv - fc.y0;
    const float t = fminf(fmaxf(t0, 0.f), Ly);
    // The knots are built EXACTLY as in the forward kernel (maf_fused_sm100.cu, spline8): exponentials without the
    // maximum subtracted (same fallback), prefix sums, knots in the unnormalised domain -- a X_k = cw_{k-1} + k a min_bin
    // with a = sw / Rw, b Y_k = ch_{k-1} + k b min_bin with b = sh / Rh -- so that both directions see bit-identical
    // bin parameters and the round trip closes to rounding.
    float ew[8], eh[8], cw[8], ch[8];
    auto sums = [&](float mw, float mh) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            ew[k] = ex2(p[k] - mw);
            eh[k] = ex2(p[8 + k] - mh);
        }
        cw[0] = ew[0]; ch[0] = eh[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) { cw[k] = cw[k - 1] + ew[k]; ch[k] = ch[k - 1] + eh[k]; }
    };
    sums(0.f, 0.f);
    {
        const float lo = fminf(cw[7], ch[7]), hi = fmaxf(cw[7], ch[7]);
        if (__any_sync(0xffffffffu, !(lo > 7.9e-31f && hi < 1.3e30f))) {
            float mw = p[0], mh = p[8];
#pragma unroll
            for (int k = 1; k < 8; ++k) { mw = fmaxf(mw, p[k]); mh = fmaxf(mh, p[8 + k]); }
            sums(mw, mh);
        }
    }
    const float sw = cw[7], sh = ch[7];
    const float sa = sw * (1.f / fc.Rw), sb = sh * (1.f / fc.Rh);
    const float ty = t * sb, mbs = min_bin * sa, mbsh = min_bin * sb;
    // walk the knots along y: last bin whose bottom knot is below t
    float Us = 0.f, Vs = 0.f, ews = ew[0], ehs = eh[0];
    float raw0 = p[16], raw1 = p[17];
    const float raw_last = (MIXED && !circ) ? p[24] : p[16];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        const float Uk = fmaf((float)k, mbs, cw[k - 1]);
        const float Vk = fmaf((float)k, mbsh, ch[k - 1]);
        const bool adv = ty > Vk;
        Us = adv ? Uk : Us;
        Vs = adv ? Vk : Vs;
        ews = adv ? ew[k] : ews;
        ehs = adv ? eh[k] : ehs;
        raw0 = adv ? p[16 + k] : raw0;
        raw1 = adv ? (k == 7 ? raw_last : p[16 + k + 1]) : raw1;
    }
    const float iw = rcp(ews + mbs);                 // 1 / (a w)
    const float g = fc.Rh * rcp(sh);                 // 1 / b
    const float h_sel = (ehs + mbsh) * g;
    const float s = h_sel * (iw * sa);               // h / w
    const float yk = Vs * g;
    const float dk = softplus_l2(raw0 + slope_offset2) + min_slope;
    const float dk1 = softplus_l2(raw1 + slope_offset2) + min_slope;
    const float yr = t - yk;
    const float q = dk1 + dk - 2.f * s;
    const float a = fmaf(h_sel, s - dk, yr * q);
    const float b = fmaf(h_sel, dk, -yr * q);
    const float c = -s * yr;
    const float disc = fmaxf(fmaf(b, b, -4.f * a * c), 0.f);
    const float e = 2.f * c * rcp(-b - sqrt_approx(disc));
    const float ome = 1.f - e, u = e * ome, e2 = e * e;
    const float den = fmaf(q, u, s);
    const float iden = rcp(den);
    const float nn = fmaf(dk1, e2, fmaf(2.f * s, u, dk * ome * ome));
    const float rr = s * iden;
    float xr = fmaf(e, ews + mbs, Us) * rcp(sa);     // (a x_k + e a w) / a
    float ld = -LN2 * lg2(nn * rr * rr);
    if (circ) {
        // un-shift and wrap into [x0, x0 + L)
        xr -= p[24];
        xr = xr - fc.L * floorf(xr * fc.invL);
        xr = (xr < 0.f) ? xr + fc.L : xr;
        xr = (xr >= fc.L) ? xr - fc.L : xr;
    } else {
        // linear tails: below y0 the first bin is selected (dk = slope at knot 0), above yf the last one (dk1)
        const bool lo = t0 < 0.f, hi = t0 > Ly;
        const float xt = lo ? t0 * rcp(dk) : fmaf(t0 - Ly, rcp(dk1), fc.L);
        const float lt = -LN2 * lg2(lo ? dk : dk1);
        xr = (lo || hi) ? xt : xr;
        ld = (lo || hi) ? lt : ld;
    }
    x = fc.x0 + xr;
    return ld;
}

__device__ __forceinline__ float elu_l2(float t, float l2e) { return fmaxf(t, fmaf(ex2(fminf(t, 0.f)), l2e, -LOG2E)); }

template <bool MIXED>
__global__ void __launch_bounds__(THREADS, 1) maf_spline_inv_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* sW = smem_raw;
    float* sX0 = reinterpret_cast<float*>(sW + (size_t)STAGES * STAGE_BYTES);   // two y / x tiles, row-major [128][D]
    const int x_tile_bytes = TILE_M * p.D * 4;
    const int x_tile_stride = (x_tile_bytes + 127) & ~127;
    float* sLd = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sX0) + 2 * x_tile_stride);   // [2][128]
    Smem* sm = reinterpret_cast<Smem*>(sLd + 2 * TILE_M);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = p.n_tiles;
    const int n_items = p.n_layers * n_tiles;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm->w_full[s], 1); mbar_init(&sm->w_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm->x_full[b], 1); mbar_init(&sm->x_empty[b], 1); mbar_init(&sm->y_ready[b], EPI_THREADS);
        }
        mbar_init(&sm->a_ready, EPI_THREADS);
        mbar_init(&sm->acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&sm->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    if (warp == 0) {
        // =========================== producer: input tiles + weight blocks ===========================
        uint32_t stage = 0, wphase = 0, tcount = 0;
        auto flag_of = [&](int item) -> const uint32_t* {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            return layer == 0 ? nullptr : p.flags + (size_t)(layer - 1) * n_tiles + tile;
        };
        auto request_x = [&](int item, uint32_t t) {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            const int rows = min(TILE_M, p.batch - tile * TILE_M);
            const uint32_t b = t & 1;
            const float* src = layer == 0 ? p.y : p.x;
            if (layer > 0) {
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
                fence_proxy_async();
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
        const bool far = n_tiles >= 2 * (int)gridDim.x;
        bool need_x = true;
        int layer = 0, tile = blockIdx.x;
        while (layer < p.n_layers) {
            const int item = layer * n_tiles + tile;
            const int n_ops = p.layers[layer].n_ops;
            const Op* ops = p.layers[layer].ops;
            const uint8_t* weights = p.layers[layer].weights;
            const int prefetch_at = n_ops > 8 ? 8 : n_ops - 1;
            const int must_at = n_ops > STAGES - 1 ? STAGES - 1 : n_ops - 1;
            for (int i = 0; i < n_ops; ++i) {
                if (need_x) {
                    const uint32_t* flag = flag_of(item);
                    if (i == must_at || flag == nullptr || ld_acquire_gpu(flag) == p.epoch) {
                        request_x(item, tcount);
                        need_x = false;
                    }
                }
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(ops + i));
                const uint32_t n = raw.z & 0xffffu, ksteps = (raw.w >> 16) & 0xffu;
                const uint32_t bytes = n * ksteps * 32u;
                mbar_wait(&sm->w_empty[stage], wphase ^ 1, p.error, 2);
                if (elect_one()) {
                    mbar_expect_tx(&sm->w_full[stage], bytes);
                    bulk_g2s(sW + (size_t)stage * STAGE_BYTES, weights + raw.x, bytes, &sm->w_full[stage]);
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
    } else if (warp == 1) {
        // =========================== MMA issuer: one dependent chain of small products ===========================
        uint32_t stage = 0, wphase = 0, a_par = 0;
        const uint32_t w_base16 = smem_u32(sW) >> 4;
        int layer = 0, tile = blockIdx.x;
        while (layer < p.n_layers) {
            const int n_ops = p.layers[layer].n_ops;
            const Op* ops = p.layers[layer].ops;
            uint4 nxt = __ldg(reinterpret_cast<const uint4*>(ops));
            for (int i = 0; i < n_ops; ++i) {
                const uint4 raw = nxt;
                if (i + 1 < n_ops) nxt = __ldg(reinterpret_cast<const uint4*>(ops + i + 1));
                const uint32_t idesc = raw.y, n = raw.z & 0xffffu, d_col = raw.z >> 16, a_col = raw.w & 0xffffu;
                const uint32_t ksteps = (raw.w >> 16) & 0xffu;
                // everything that does not depend on the previous product first: the weights and the descriptors ...
                mbar_wait(&sm->w_full[stage], wphase, p.error, 5);
                const uint32_t b_lo = w_base16 + stage * (STAGE_BYTES >> 4) + (n << 16);
                const uint32_t a_tmem = tmem + a_col, d_tmem = tmem + d_col;
                // ... then the hand-over: every product needs what the previous one produced
                mbar_wait(&sm->a_ready, a_par, p.error, 3);
                a_par ^= 1u;
                tc_fence_after();
                if (elect_one()) {
                    uint32_t accumulate = 0u;
                    for (uint32_t ks = 0; ks < ksteps; ++ks) {
                        const uint64_t db = ((uint64_t)DESC_HI << 32) | (b_lo + ks * 2u * n);
                        umma_ts(d_tmem, a_tmem + ks * 8u, db, idesc, accumulate);
                        accumulate = 1u;
                    }
                    umma_commit(&sm->w_empty[stage]);
                    umma_commit(&sm->acc_full);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; wphase ^= 1; }
            }
            tile += gridDim.x;
            while (tile >= n_tiles) { tile -= n_tiles; ++layer; }
        }
    } else if (warp == 2) {
        // =========================== x store, log-det, publish ===========================
        uint32_t tcount = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            const int rows = min(TILE_M, p.batch - tile * TILE_M);
            const uint32_t xb = tcount & 1;
            const float* sY = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(sX0) + xb * x_tile_stride);
            mbar_wait(&sm->y_ready[xb], (tcount >> 1) & 1, p.error, 10);
            float* dst = p.x + (size_t)tile * TILE_M * p.D;
            if (rows == TILE_M) {
                if (elect_one()) bulk_s2g(dst, sY, (uint32_t)x_tile_bytes);
                __syncwarp();
            }
            const float* ldp = sLd + xb * TILE_M;
#pragma unroll
            for (int r = 0; r < TILE_M / 32; ++r) {
                const int row = r * 32 + lane;
                if (row < rows) {
                    float v = ldp[row];
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
                if (lane == 0) {
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
        }
    } else if (warp >= 4) {
        // =========================== epilogue warpgroup ===========================
        const int et = tid - 128;                 // 0..127 = sample row = TMEM lane
        const int row = et;
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc_par = 0, tcount = 0;
        float l2e;
        asm volatile("mov.f32 %0, 0f3FB8AA3B;" : "=f"(l2e));
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tcount) {
            const int layer = item / n_tiles, tile = item - layer * n_tiles;
            const LayerP& L = p.layers[layer];
            const int rows = min(TILE_M, p.batch - tile * TILE_M);
            const uint32_t xb = tcount & 1;
            float* sX = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sX0) + xb * x_tile_stride);
            float* xrow = sX + row * p.D;
            // ---- reset the A operands: zeros, the constant-one columns of x, h1 and h2 ----
            {
                uint32_t z[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) z[i] = 0u;
                for (int c = 0; c < p.K1 / 2; c += 8) tmem_st8(lane_addr + A0_COL + c, z);
                for (int c = 0; c < p.HP / 2; c += 8) {
                    tmem_st8(lane_addr + A1_COL + c, z);
                    tmem_st8(lane_addr + A2_COL + c, z);
                }
                tmem_st_wait();
                const uint32_t one2 = pack_bf16(1.f, 1.f);
                if (p.Din & 1) {
                    tmem_st1(lane_addr + A0_COL + p.Din / 2, pack_bf16(0.f, 1.f));     // columns Din (high half), Din + 1
                    tmem_st1(lane_addr + A0_COL + p.Din / 2 + 1, pack_bf16(1.f, 0.f));
                } else {
                    tmem_st1(lane_addr + A0_COL + p.Din / 2, one2);
                }
                tmem_st1(lane_addr + A1_COL, one2);
                tmem_st1(lane_addr + A2_COL, one2);
                tmem_st_wait();
            }
            mbar_wait(&sm->x_full[xb], (tcount >> 1) & 1, p.error, 6);
            if (rows < TILE_M) {
                const float* src = (layer == 0 ? p.y : p.x) + (size_t)tile * TILE_M * p.D;
                for (int i = et; i < TILE_M * p.D; i += EPI_THREADS) sX[i] = i < rows * p.D ? __ldcg(src + i) : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            if (L.init_map != nullptr) {
                // conditioning features (degree -1) are known from the start: stage them (lifted to (cos, sin) where the
                // embedding says so) next to the constant ones; everything else starts at zero
                auto input = [&](int k) -> float {
                    const int e = __ldg(L.init_map + k);
                    const int what = e >> 16;
                    float v = what >= 3 ? (what == 3 ? 1.f : 0.f) : xrow[e & 0xffff];
                    if (what == 1 || what == 2) {
                        const float ang = (v - L.emb_lower) * L.emb_scale;
                        v = what == 1 ? __cosf(ang) : __sinf(ang);
                    }
                    return v;
                };
                for (int c0 = 0; c0 < p.K1 / 2; c0 += 8) {
                    uint32_t q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) q[i] = pack_bf16(input((c0 + i) * 2), input((c0 + i) * 2 + 1));
                    tmem_st8(lane_addr + A0_COL + c0, q);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(&sm->a_ready);
            // ---- degree sweep ----
            float ld = 0.f;
            float open1 = 0.f, open2 = 0.f;          // last unit written to an odd-length prefix of h1 / h2 (low half of a pair)
            const int n_steps = L.n_steps;
            const float min_bin = L.min_bin, min_slope = L.min_slope, slope_offset2 = L.slope_offset2;
            for (int d = 0; d < n_steps; ++d) {
                const Step st = L.steps[d];
                // one hand-over per product that follows in this item (the last x of a tile has no consumer)
                const bool last = d + 1 == n_steps;
                // OUT_d: parameters of the feature -> x_d  (steps with col < 0 only bring in the hidden units that depend
                // on conditioning features alone: no feature to invert, the hand-over at the start of the item released them)
                if (st.col >= 0) {
                mbar_wait(&sm->acc_full, acc_par, p.error, 8);
                acc_par ^= 1u;
                tc_fence_after();
                {
                    uint32_t r[32];
                    tmem_ld16(lane_addr + ACC_OUT, r);
                    tmem_ld8(lane_addr + ACC_OUT + 16, r + 16);
                    tmem_ld1(lane_addr + ACC_OUT + 24, r + 24);
                    tmem_wait8(r); tmem_wait8(r + 8); tmem_wait8(r + 16); tmem_wait1(r + 24);
                    float xv;
                    ld += spline8_inverse<MIXED>(r, xrow[st.col], st, min_bin, min_slope, slope_offset2, xv);
                    xrow[st.col] = xv;
                    // bf16 pair column of the conditioner input: (cos, sin) of a lifted periodic feature, or the
                    // feature next to its partner, which is known (already x), the constant one, or still zero
                    const int ic = (st.partner >> 8) & 0xff;
                    uint32_t q;
                    if (st.partner & 32) {
                        const float ang = (xv - L.emb_lower) * L.emb_scale;
                        q = pack_bf16(__cosf(ang), __sinf(ang));
                    } else {
                        const int pk = st.partner & 15;
                        const float other = pk == 1 ? xrow[(st.partner >> 16) & 0xff] : (pk == 2 ? 1.f : 0.f);
                        q = (ic & 1) ? pack_bf16(other, xv) : pack_bf16(xv, other);
                    }
                    tmem_st1(lane_addr + A0_COL + (ic >> 1), q);
                    tmem_st_wait();
                }
                tc_fence_before();
                if (st.h1_n != 0 || st.h2_n != 0 || !last) mbar_arrive(&sm->a_ready);
                }
                // H1_d, H2_d: the hidden units of this degree -> A operands (up to 15 units, 8 pair columns; the
                // padding rows of the block give exact zeros, which is what the not-yet-known units must hold)
#pragma unroll
                for (int hl = 0; hl < 2; ++hl) {
                    const int ha = hl == 0 ? st.h1_a : st.h2_a, hn = hl == 0 ? st.h1_n : st.h2_n;
                    if (hn == 0) continue;
                    mbar_wait(&sm->acc_full, acc_par, p.error, 7);
                    acc_par ^= 1u;
                    tc_fence_after();
                    uint32_t r[16];
                    tmem_ld16(lane_addr + ACC_HID, r);
                    tmem_wait8(r); tmem_wait8(r + 8);
                    float v[17];
                    float& open = hl == 0 ? open1 : open2;
                    const int odd = ha & 1;
                    v[0] = open;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i + 1] = elu_l2(__uint_as_float(r[i]), l2e);
                    // units ha .. ha + hn - 1 = v[1 .. hn]; pair columns start at ha / 2
                    uint32_t q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float lo = odd ? v[2 * i] : v[2 * i + 1];
                        const float hi = odd ? v[2 * i + 1] : (2 * i + 2 <= 16 ? v[2 * i + 2] : 0.f);
                        q[i] = pack_bf16(lo, hi);
                    }
                    // remember the last unit if it opens a new pair column (its partner arrives with the next degree)
                    {
                        float last = 0.f;
#pragma unroll
                        for (int i = 1; i <= 16; ++i) last = (i == hn) ? v[i] : last;
                        open = ((ha + hn) & 1) ? last : 0.f;
                    }
                    tmem_st8(lane_addr + (hl == 0 ? A1_COL : A2_COL) + (ha >> 1), q);
                    tmem_st_wait();
                    tc_fence_before();
                    if ((hl == 0 && st.h2_n != 0) || !last) mbar_arrive(&sm->a_ready);
                }
            }
            sLd[xb * TILE_M + row] = ld;
            fence_async_smem();
            mbar_arrive(&sm->y_ready[xb]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

size_t smem_bytes(const Params& p) {
    size_t s = (size_t)STAGES * STAGE_BYTES;
    s += 2 * (((size_t)TILE_M * p.D * 4 + 127) & ~(size_t)127);
    s += 2 * TILE_M * 4 + sizeof(Smem);
    return s + 256;
}

}  // namespace finv
}  // namespace tfepb

using namespace tfepb;

static_assert(sizeof(finv::Op) == sizeof(tfepb_fused_op), "schedule entry layout mismatch");
static_assert(sizeof(finv::Step) == sizeof(tfepb_fused_inv_step), "step table layout mismatch");

extern "C" int tfepb_maf_spline_inverse_bf16(const tfepb_fused_inv_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->y && a->x && a->logdet && a->layers, "null buffer");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->n_features > 0, "bad sizes");
    TFEPB_CHECK_ARG(a->n_layers >= 1 && a->n_layers <= finv::MAX_LAYERS, "n_layers must be in [1, %d]", finv::MAX_LAYERS);
    TFEPB_CHECK_ARG(a->n_layers == 1 || a->tile_flags != nullptr, "a chain of layers needs the tile_flags workspace");
    const int n_inputs = a->n_inputs > 0 ? a->n_inputs : a->n_features;
    TFEPB_CHECK_ARG(n_inputs >= a->n_features && n_inputs <= 2 * a->n_features, "n_inputs must be in [n_features, 2 n_features]");
    TFEPB_CHECK_ARG(a->k1 % 16 == 0 && a->k1 >= n_inputs + 2 && a->k1 <= 2 * (finv::A1_COL - finv::A0_COL),
                    "k1 must hold n_inputs + 2 bias columns, rounded up to 16, and fit the tensor-memory plan");
    TFEPB_CHECK_ARG(a->n_features <= 256, "at most 256 features (step table encoding)");
    TFEPB_CHECK_ARG(a->hidden_padded % 16 == 0 && a->hidden_padded > 0 && a->hidden_padded <= 352,
                    "hidden width (padded) must be a multiple of 16 and at most 352 (tensor-memory plan)");
    TFEPB_CHECK_ARG((a->n_features * 4 * tc::TILE_M) % 16 == 0, "tile of y must be a multiple of 16 bytes");
    TFEPB_CHECK_ARG(((uintptr_t)a->y % 16 == 0) && ((uintptr_t)a->x % 16 == 0), "y and x must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    finv::Params p{};
    p.y = (const float*)a->y; p.x = (float*)a->x; p.logdet = (float*)a->logdet;
    p.batch = a->batch; p.D = a->n_features; p.K1 = a->k1; p.HP = a->hidden_padded;
    p.Din = n_inputs;
    p.n_layers = a->n_layers;
    p.n_tiles = (a->batch + tc::TILE_M - 1) / tc::TILE_M;
    p.flags = a->tile_flags; p.epoch = a->epoch; p.error = a->error_flag;
    for (int l = 0; l < a->n_layers; ++l) {
        const tfepb_fused_inv_layer& s = a->layers[l];
        TFEPB_CHECK_ARG(s.ops && s.steps && s.weights, "layer %d: null buffer", l);
        TFEPB_CHECK_ARG(s.n_ops > 0 && s.n_steps > 0, "layer %d: empty schedule", l);
        TFEPB_CHECK_ARG((uintptr_t)s.weights % 16 == 0 && (uintptr_t)s.ops % 16 == 0 && (uintptr_t)s.steps % 16 == 0,
                        "layer %d: tables and packed weights must be 16-byte aligned", l);
        finv::LayerP& d = p.layers[l];
        d.weights = (const uint8_t*)s.weights;
        d.ops = (const finv::Op*)s.ops; d.steps = (const finv::Step*)s.steps;
        d.n_ops = s.n_ops; d.n_steps = s.n_steps;
        d.min_bin = s.min_bin_size; d.min_slope = s.min_slope; d.slope_offset2 = s.slope_offset * tc::LOG2E;
        d.emb_lower = s.emb_lower; d.emb_scale = s.emb_scale;
        d.init_map = s.init_map;
    }
    const size_t smem = finv::smem_bytes(p);
    TFEPB_CHECK_ARG(smem <= 227 * 1024, "shared memory plan of %zu bytes exceeds 227 KB", smem);
    const bool mixed = a->mixed_splines != 0;      // some features are not circular: generic spline epilogue
    auto kernel = mixed ? finv::maf_spline_inv_kernel<true> : finv::maf_spline_inv_kernel<false>;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem)) return rc;
    int grid = 0;      // co-resident CTAs only (tile flags between CTAs), enforced by a cooperative launch
    if (int rc = coresident_grid(reinterpret_cast<const void*>(kernel), finv::THREADS, smem, p.n_tiles, &grid)) return rc;
    if (int rc = launch_cooperative(reinterpret_cast<const void*>(kernel), grid, finv::THREADS, smem, &p, as_stream(stream))) return rc;
    return check_launch("maf_spline_inv_kernel");
}
