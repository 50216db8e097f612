// Persistent inverse sweep of one MAF layer (exact fp32 / fp64 arithmetic).
//
//   x, logdet = MAF.inverse(y)                      (reference: nn/flows/autoregressive.py:179-229)
//
// The reference inverts an autoregressive flow with n_degrees FULL passes of the conditioner (pass i fixes
// the features of degree i, autoregressive.py:216-227).  With hidden units sorted by degree
// (tfep_b200/_pack.py) every unit can be evaluated exactly once, as soon as its inputs exist:
//
//   for each degree d (ascending):
//       par_d  = W_out[rows of the features of degree d] . h_last[units of degree < d] + b    (A)
//       x_d    = T^-1(y_d ; par_d),  logdet += ...                                            (T)
//       h_l[units of degree d] = ELU(W_l[those rows] . h_{l-1}[units of degree <= d] + b)     (H, l = 1..)
//
// i.e. nnz(masks) multiply-accumulates per sample -- the cost of ONE forward pass -- instead of n_degrees
// passes.  The whole sweep runs in one persistent kernel: a CTA owns a tile of TS samples and keeps x and
// every hidden activation of the tile in shared memory across all degrees (nothing but y, x and logdet
// touches HBM; the weights stream from L2 through L1, each row being read once per tile).  The stages
// of a degree are separated by CTA barriers; inside a stage a warp owns 32 samples (lane = sample, so the
// 16-byte activation loads are conflict-free with the padded leading dimensions chosen by the host) and a
// block of up to 8 output rows, whose weights arrive as warp-uniform 16-byte loads; short stages are split
// along the reduction and combined in a fixed order (deterministic results).
//
// Transformers run through the same device operators as the stand-alone kernels (tx_ops.cuh), reading
// the parameters of the degree group from shared memory.
#include "common.cuh"
#include "tx_math.cuh"
#include "tx_ops.cuh"
#include "tc_ptx.cuh"

namespace tfepb {
namespace sweep {

constexpr int MAXL = TFEPB_SWEEP_MAX_LINEAR;
constexpr int RB = 8;                       // output rows per register block

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

__device__ __forceinline__ float dot_acc(float4 w, float4 h, float a) {
    a = fmaf(w.x, h.x, a); a = fmaf(w.y, h.y, a); a = fmaf(w.z, h.z, a); return fmaf(w.w, h.w, a);
}
__device__ __forceinline__ double dot_acc(double2 w, double2 h, double a) { return fma(w.y, h.y, fma(w.x, h.x, a)); }

template <typename T>
struct Params {
    const T* y; int64_t ldy;
    T* x; int64_t ldx;
    T* logdet;
    int batch, D, n_linear;
    const T* w[MAXL]; const T* b[MAXL];
    int ldw[MAXL];
    int act_ld[MAXL];                 // shared-memory leading dimensions: [0] x tile, [l] hidden layer l
    int par_ld;
    const tfepb_sweep_group* groups; int n_groups;
    const tfepb_sweep_part* parts;
    const tfepb_sweep_group_part* gparts;
    const int* ids;
    const int* fixed_cols; int n_fixed;
    int stage_elems;                  // weights of one degree group (elements), staged kernel only
    // optional PeriodicEmbedding of the conditioner input (nn/embeddings/mafembed.py): act[0] then holds the
    // embedded features (width E) and x lives in its own tile
    const int* emb_out_col; const int* emb_periodic;
    T emb_lower, emb_scale;
    int E, xs_ld;
    int n_hidden[MAXL];               // [l] = units of hidden layer l (= n_out of linear layer l - 1)
    // blocked sweeps of wide conditioners (tfep_b200/_blocked.py): per linear layer an optional per-sample term added to
    // the pre-activations (what the units of EARLIER degree blocks contribute, computed by a plain GEMM), and optional
    // global buffers that receive the hidden activations of the tile (the inputs of the LATER blocks' GEMMs)
    const T* extra[MAXL]; int64_t ldextra[MAXL];
    T* act_out[MAXL]; int64_t ldact_out[MAXL];
};

// out[s][ocol0 + (r - r0)] = act(bias[r] + sum_k W[r][k] in[s][k]) for r in [r0, r1), all TS samples of the tile.
// Ends with a CTA barrier.  K is rounded up to the vector width: the host pads the weight rows with zeros and
// activations that do not exist yet are zero.
template <typename T, int TS, int THREADS, bool WSMEM = false>
__device__ void gemv_stage(const T* __restrict__ W, int ldw, const T* __restrict__ bias, int r0, int r1, int K,
                           const T* in, int ldin, T* out, int ldout, int ocol0, bool act, T* scratch,
                           const T* __restrict__ extra = nullptr, int64_t ldextra = 0, int rows = TS) {
    // `extra` (already offset to the first sample of the tile): extra[s * ldextra + r] is added to row r of sample s
    using V = typename Vec<T>::type;
    constexpr int NV = Vec<T>::N;
    constexpr int SG = TS / 32;                  // warps side by side over the samples
    constexpr int NWG = THREADS / 32 / SG;       // warp groups sharing the rows / the reduction
    const int R = r1 - r0;
    if (R <= 0) return;                          // uniform over the CTA
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sg = warp % SG, wg = warp / SG;
    const int s = sg * 32 + lane;
    const int nrb = (R + RB - 1) / RB;
    const int rbs = (R + nrb - 1) / nrb;         // balanced block size <= RB
    const int Kv = (K + NV - 1) / NV;
    int nks = NWG / nrb;
    if (nks > Kv / 8) nks = Kv / 8;              // at least 8 vector steps per slice
    if (nks < 1) nks = 1;
    const int items = nrb * nks;
    const V* inv = reinterpret_cast<const V*>(in + (size_t)s * ldin);
    for (int it = wg; it < items; it += NWG) {
        const int rb = it / nks, ks = it - rb * nks;
        const int ra = r0 + rb * rbs;
        const int nr = min(rbs, r1 - ra);
        const int kv0 = (Kv * ks) / nks, kv1 = (Kv * (ks + 1)) / nks;
        T acc[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[j] = T(0);
        const T* wrow = W + (size_t)ra * ldw;
#pragma unroll 2
        for (int kv = kv0; kv < kv1; ++kv) {
            const V h = inv[kv];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                if (j < nr) {
                    const V* wp = reinterpret_cast<const V*>(wrow + (size_t)j * ldw) + kv;
                    const V wv = WSMEM ? *wp : __ldg(wp);
                    acc[j] = dot_acc(wv, h, acc[j]);
                }
            }
        }
        if (nks == 1) {
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                if (j < nr) {
                    T v = acc[j] + bias[ra + j];
                    if (extra != nullptr && s < rows) v += extra[(int64_t)s * ldextra + ra + j];
                    out[(size_t)s * ldout + ocol0 + (ra - r0) + j] = act ? tfepb::elu(v) : v;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < RB; ++j)
                if (j < nr) scratch[((size_t)it * RB + j) * TS + s] = acc[j];
        }
    }
    if (nks > 1) {
        __syncthreads();
        for (int o = threadIdx.x; o < R * TS; o += THREADS) {
            const int r = o / TS, ss = o - r * TS;
            const int rb = r / rbs, j = r - rb * rbs;
            T v = bias[r0 + r];
            if (extra != nullptr && ss < rows) v += extra[(int64_t)ss * ldextra + r0 + r];
            for (int ks = 0; ks < nks; ++ks) v += scratch[((size_t)(rb * nks + ks) * RB + j) * TS + ss];
            out[(size_t)ss * ldout + ocol0 + r] = act ? tfepb::elu(v) : v;
        }
    }
    __syncthreads();
}

// conditioner-input columns of x column c of sample s (identity without an embedding: x IS the conditioner input)
template <typename T>
__device__ __forceinline__ void embed_column(const Params<T>& p, T* act0, const T* xs, int s, int c) {
    if (p.emb_out_col == nullptr) return;
    const T v = xs[(size_t)s * p.xs_ld + c];
    T* o = act0 + (size_t)s * p.act_ld[0] + p.emb_out_col[c];
    if (p.emb_periodic[c]) {
        const T a = (v - p.emb_lower) * p.emb_scale;
        o[0] = cos(a);
        o[1] = sin(a);
    } else {
        o[0] = v;
    }
}

template <typename T, int MAXK>
__device__ __forceinline__ SplineOp<T, MAXK> make_spline(const tfepb_sweep_part& d) {
    SplineOp<T, MAXK> op;
    op.K = d.n_bins; op.circular = d.circular; op.idslopes = d.identity_boundary_slopes;
    op.learn_lo = d.learn_lower_bound; op.learn_hi = d.learn_upper_bound;
    op.x0 = (const T*)d.x0; op.xf = (const T*)d.xf; op.y0 = (const T*)d.y0; op.yf = (const T*)d.yf;
    op.min_bin = (T)d.min_bin_size; op.min_slope = (T)d.min_slope;
    op.bins = nullptr; op.ldbins = 0;
    return op;
}

// Inverse transformer of the features of one degree group: thread s handles sample s of the tile.
template <typename T, int TS, int THREADS>
__device__ void transform_stage(const Params<T>& p, const tfepb_sweep_group& g, int64_t tile0, int rows, T* xs, T* act0,
                                const T* par, T* ldacc) {
    const int s = threadIdx.x;
    if (s < TS && s < rows && g.part_count > 0) {
        TxView<T> v{};
        v.x = p.y + tile0 * p.ldy; v.ldx = p.ldy;          // source: y (global), local row index
        v.y = xs; v.ldy = p.xs_ld;                         // destination: the x tile in shared memory
        v.par = par; v.ldp = p.par_ld; v.poff = -(int64_t)g.out_r0; v.sp = 1; v.sf = 0;
        v.logdet = nullptr; v.accumulate = 0; v.B = rows; v.inverse = 1;
        T ld = T(0);
        for (int q = 0; q < g.part_count; ++q) {
            const tfepb_sweep_group_part gp = p.gparts[g.part_first + q];
            const tfepb_sweep_part& d = p.parts[gp.part];
            v.pbase = d.par_base; v.cols = d.cols; v.ids = p.ids + gp.ids_offset; v.F = gp.n_ids;
            if (d.kind == TFEPB_SWEEP_AFFINE) {
                AffineOp<T> op;
                for (int u = 0; u < gp.n_ids; ++u) ld += op.apply(v, s, u);
            } else if (d.kind == TFEPB_SWEEP_SHIFT) {
                const ShiftOp<T> op{(const T*)d.x0, (const T*)d.xf};      // period / lower tables travel in x0 / xf
                for (int u = 0; u < gp.n_ids; ++u) ld += op.apply(v, s, u);
            } else if (d.kind == TFEPB_SWEEP_SPLINE) {
                if (d.n_bins <= 8) {
                    const SplineOp<T, 8> op = make_spline<T, 8>(d);
                    for (int u = 0; u < gp.n_ids; ++u) ld += op.apply(v, s, u);
                } else {
                    const SplineOp<T, 32> op = make_spline<T, 32>(d);
                    for (int u = 0; u < gp.n_ids; ++u) ld += op.apply(v, s, u);
                }
            } else {
                const MoebiusOp<T> op{d.dimension, (T)d.max_radius, d.unit_sphere};
                for (int u = 0; u < gp.n_ids / d.dimension; ++u) ld += op.apply(v, s, u);
            }
        }
        ldacc[s] += ld;
        if (p.emb_out_col != nullptr) {            // lift the freshly inverted features into the conditioner input
            for (int q = 0; q < g.part_count; ++q) {
                const tfepb_sweep_group_part gp = p.gparts[g.part_first + q];
                const tfepb_sweep_part& d = p.parts[gp.part];
                for (int u = 0; u < gp.n_ids; ++u) {
                    const int f = p.ids[gp.ids_offset + u];
                    embed_column<T>(p, act0, xs, s, d.cols ? d.cols[f] : f);
                }
            }
        }
    }
    __syncthreads();
}

template <typename T, int TS, int THREADS>
__global__ void __launch_bounds__(THREADS) maf_inverse_sweep_kernel(const Params<T> p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    T* act[MAXL];
    T* cur = reinterpret_cast<T*>(smem_raw);
    for (int l = 0; l < p.n_linear; ++l) { act[l] = cur; cur += (size_t)TS * p.act_ld[l]; }
    T* par = cur; cur += (size_t)TS * p.par_ld;
    T* scratch = cur; cur += (size_t)(THREADS / 32 / (TS / 32)) * RB * TS;
    T* ldacc = cur; cur += TS;
    T* xs = p.emb_out_col ? cur : act[0];
    const int L = p.n_linear;
    const int n_tiles = (p.batch + TS - 1) / TS;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile0 = (int64_t)tile * TS;
        const int rows = min(TS, p.batch - (int)tile0);
        // ---- reset the tile state: x = 0 (+ conditioning features copied from y), activations = 0 ----
        for (int l = 0; l < L; ++l)
            for (int i = threadIdx.x; i < TS * p.act_ld[l]; i += THREADS) act[l][i] = T(0);
        if (threadIdx.x < TS) ldacc[threadIdx.x] = T(0);
        if (p.emb_out_col)
            for (int i = threadIdx.x; i < TS * p.xs_ld; i += THREADS) xs[i] = T(0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * p.n_fixed; i += THREADS) {
            const int s = i / p.n_fixed, c = p.fixed_cols[i - s * p.n_fixed];
            xs[(size_t)s * p.xs_ld + c] = p.y[(tile0 + s) * p.ldy + c];
            embed_column<T>(p, act[0], xs, s, c);
        }
        __syncthreads();
        // ---- degree sweep ----
        for (int gi = 0; gi < p.n_groups; ++gi) {
            const tfepb_sweep_group g = p.groups[gi];
            if (g.out_r1 > g.out_r0) {
                gemv_stage<T, TS, THREADS>(p.w[L - 1], p.ldw[L - 1], p.b[L - 1], g.out_r0, g.out_r1, g.out_k, act[L - 1], p.act_ld[L - 1],
                                  par, p.par_ld, 0, false, scratch,
                                  p.extra[L - 1] ? p.extra[L - 1] + tile0 * p.ldextra[L - 1] : nullptr, p.ldextra[L - 1], rows);
                transform_stage<T, TS, THREADS>(p, g, tile0, rows, xs, act[0], par, ldacc);
            }
            for (int l = 1; l < L; ++l)
                gemv_stage<T, TS, THREADS>(p.w[l - 1], p.ldw[l - 1], p.b[l - 1], g.h_a[l - 1], g.h_b[l - 1], g.h_k[l - 1], act[l - 1],
                                  p.act_ld[l - 1], act[l], p.act_ld[l], g.h_a[l - 1], true, scratch,
                                  p.extra[l - 1] ? p.extra[l - 1] + tile0 * p.ldextra[l - 1] : nullptr, p.ldextra[l - 1], rows);
        }
        for (int l = 1; l < L; ++l) {           // hidden activations of the tile -> global (blocked sweeps)
            if (p.act_out[l] == nullptr) continue;
            const int width = p.n_hidden[l];
            for (int i = threadIdx.x; i < rows * width; i += THREADS) {
                const int s = i / width, c = i - s * width;
                p.act_out[l][(tile0 + s) * p.ldact_out[l] + c] = act[l][(size_t)s * p.act_ld[l] + c];
            }
        }
        // ---- write the tile ----
        for (int i = threadIdx.x; i < rows * p.D; i += THREADS) {
            const int s = i / p.D, c = i - s * p.D;
            p.x[(tile0 + s) * p.ldx + c] = xs[(size_t)s * p.xs_ld + c];
        }
        if (threadIdx.x < rows) p.logdet[tile0 + threadIdx.x] = ldacc[threadIdx.x];
        __syncthreads();
    }
}

// Same sweep with the weights of every degree group STAGED in shared memory: one elected thread bulk-copies the
// rows of degree d + 1 (output rows + hidden rows, full width, contiguous in the packed matrices) into the other
// half of a double buffer while the CTA works on degree d, so the dot products read their weights as
// conflict-free shared-memory broadcasts instead of waiting on L1 / L2 (the unstaged kernel is latency bound on
// exactly those loads).  Used when the tile state of 32 samples plus two groups of weights fit (fp32, cfg1 / cfg2).
template <typename T, int TS, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) maf_inverse_sweep_staged_kernel(const Params<T> p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    T* act[MAXL];
    T* cur = reinterpret_cast<T*>(smem_raw);
    for (int l = 0; l < p.n_linear; ++l) { act[l] = cur; cur += (size_t)TS * p.act_ld[l]; }
    T* par = cur; cur += (size_t)TS * p.par_ld;
    T* scratch = cur; cur += (size_t)(THREADS / 32 / (TS / 32)) * RB * TS;
    T* ldacc = cur; cur += TS;
    T* xs = p.emb_out_col ? cur : act[0];
    if (p.emb_out_col) cur += (size_t)TS * p.xs_ld;
    cur = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(cur) + 127) & ~(uintptr_t)127);
    T* wbuf[2] = {cur, cur + p.stage_elems};
    uint64_t* bars = reinterpret_cast<uint64_t*>(cur + 2 * (size_t)p.stage_elems);
    const int L = p.n_linear;
    const int n_tiles = (p.batch + TS - 1) / TS;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t issued = 0, waited = 0;          // groups issued / consumed so far (buffer = count & 1, phase = count >> 1)
    // issue the copies of group gi into its buffer; offsets: output rows first, then the hidden layers in order
    auto prefetch = [&](int gi) {
        const tfepb_sweep_group g = p.groups[gi];
        T* dst = wbuf[issued & 1];
        uint64_t* bar = &bars[issued & 1];
        uint32_t bytes = 0;
        const uint32_t ob = (uint32_t)(g.out_r1 - g.out_r0) * p.ldw[L - 1] * sizeof(T);
        bytes += ob;
        for (int l = 1; l < L; ++l) bytes += (uint32_t)(g.h_b[l - 1] - g.h_a[l - 1]) * p.ldw[l - 1] * sizeof(T);
        if (bytes == 0) { tc::mbar_arrive(bar); ++issued; return; }
        tc::mbar_expect_tx(bar, bytes);
        if (ob) tc::bulk_g2s(dst, p.w[L - 1] + (size_t)g.out_r0 * p.ldw[L - 1], ob, bar);
        T* d = dst + (size_t)(g.out_r1 - g.out_r0) * p.ldw[L - 1];
        for (int l = 1; l < L; ++l) {
            const uint32_t hb = (uint32_t)(g.h_b[l - 1] - g.h_a[l - 1]) * p.ldw[l - 1] * sizeof(T);
            if (hb) tc::bulk_g2s(d, p.w[l - 1] + (size_t)g.h_a[l - 1] * p.ldw[l - 1], hb, bar);
            d += (size_t)(g.h_b[l - 1] - g.h_a[l - 1]) * p.ldw[l - 1];
        }
        ++issued;
    };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t tile0 = (int64_t)tile * TS;
        const int rows = min(TS, p.batch - (int)tile0);
        if (threadIdx.x == 0 && p.n_groups > 0) prefetch(0);
        for (int l = 0; l < L; ++l)
            for (int i = threadIdx.x; i < TS * p.act_ld[l]; i += THREADS) act[l][i] = T(0);
        if (threadIdx.x < TS) ldacc[threadIdx.x] = T(0);
        if (p.emb_out_col)
            for (int i = threadIdx.x; i < TS * p.xs_ld; i += THREADS) xs[i] = T(0);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * p.n_fixed; i += THREADS) {
            const int s = i / p.n_fixed, c = p.fixed_cols[i - s * p.n_fixed];
            xs[(size_t)s * p.xs_ld + c] = p.y[(tile0 + s) * p.ldy + c];
            embed_column<T>(p, act[0], xs, s, c);
        }
        __syncthreads();
        for (int gi = 0; gi < p.n_groups; ++gi) {
            const tfepb_sweep_group g = p.groups[gi];
            // every thread is past the stages of group gi - 1 (CTA barrier): its buffer may be refilled
            if (threadIdx.x == 0 && gi + 1 < p.n_groups) prefetch(gi + 1);
            tc::mbar_wait(&bars[waited & 1], (waited >> 1) & 1, nullptr, 0);
            const T* wsm = wbuf[waited & 1];
            ++waited;
            if (g.out_r1 > g.out_r0) {
                gemv_stage<T, TS, THREADS, true>(wsm - (size_t)g.out_r0 * p.ldw[L - 1], p.ldw[L - 1], p.b[L - 1], g.out_r0, g.out_r1,
                                                 g.out_k, act[L - 1], p.act_ld[L - 1], par, p.par_ld, 0, false, scratch,
                                                 p.extra[L - 1] ? p.extra[L - 1] + tile0 * p.ldextra[L - 1] : nullptr,
                                                 p.ldextra[L - 1], rows);
                transform_stage<T, TS, THREADS>(p, g, tile0, rows, xs, act[0], par, ldacc);
            }
            const T* wh = wsm + (size_t)(g.out_r1 - g.out_r0) * p.ldw[L - 1];
            for (int l = 1; l < L; ++l) {
                gemv_stage<T, TS, THREADS, true>(wh - (size_t)g.h_a[l - 1] * p.ldw[l - 1], p.ldw[l - 1], p.b[l - 1], g.h_a[l - 1],
                                                 g.h_b[l - 1], g.h_k[l - 1], act[l - 1], p.act_ld[l - 1], act[l], p.act_ld[l],
                                                 g.h_a[l - 1], true, scratch,
                                                 p.extra[l - 1] ? p.extra[l - 1] + tile0 * p.ldextra[l - 1] : nullptr,
                                                 p.ldextra[l - 1], rows);
                wh += (size_t)(g.h_b[l - 1] - g.h_a[l - 1]) * p.ldw[l - 1];
            }
            __syncthreads();
        }
        for (int l = 1; l < L; ++l) {           // hidden activations of the tile -> global (blocked sweeps)
            if (p.act_out[l] == nullptr) continue;
            const int width = p.n_hidden[l];
            for (int i = threadIdx.x; i < rows * width; i += THREADS) {
                const int s = i / width, c = i - s * width;
                p.act_out[l][(tile0 + s) * p.ldact_out[l] + c] = act[l][(size_t)s * p.act_ld[l] + c];
            }
        }
        for (int i = threadIdx.x; i < rows * p.D; i += THREADS) {
            const int s = i / p.D, c = i - s * p.D;
            p.x[(tile0 + s) * p.ldx + c] = xs[(size_t)s * p.xs_ld + c];
        }
        if (threadIdx.x < rows) p.logdet[tile0 + threadIdx.x] = ldacc[threadIdx.x];
        __syncthreads();
    }
}

template <typename T>
size_t smem_bytes(const Params<T>& p, int TS, int THREADS) {
    size_t n = 0;
    for (int l = 0; l < p.n_linear; ++l) n += (size_t)TS * p.act_ld[l];
    n += (size_t)TS * p.par_ld;
    n += (size_t)(THREADS / 32 / (TS / 32)) * RB * TS;
    n += TS;
    if (p.emb_out_col) n += (size_t)TS * p.xs_ld;
    return n * sizeof(T) + 16;
}

// leading dimension >= n whose 16-byte loads by 8 consecutive lanes (lane = row) hit 8 different bank groups
template <typename T>
int padded_ld(int n) {
    const int nv = Vec<T>::N;                    // elements per 16 bytes
    int ld = (n + nv - 1) / nv * nv;
    while (((ld / nv) & 1) == 0) ld += nv;       // odd number of 16-byte chunks per row
    return ld;
}

template <typename T>
int launch(const tfepb_sweep_args* a, cudaStream_t stream) {
    Params<T> p{};
    p.y = (const T*)a->y; p.ldy = a->ldy; p.x = (T*)a->x; p.ldx = a->ldx; p.logdet = (T*)a->logdet;
    p.batch = a->batch; p.D = a->n_features; p.n_linear = a->n_linear;
    const int nv = Vec<T>::N;
    for (int l = 0; l < a->n_linear; ++l) {
        p.w[l] = (const T*)a->w[l]; p.b[l] = (const T*)a->b[l]; p.ldw[l] = a->ldw[l];
        TFEPB_CHECK_ARG(p.w[l] && p.b[l], "layer %d: null weights", l);
        TFEPB_CHECK_ARG(a->ldw[l] % nv == 0 && ((uintptr_t)a->w[l] % 16) == 0,
                        "layer %d: weight rows must be 16-byte aligned (pad the leading dimension)", l);
        const int width = l == 0 ? (a->emb_out_col ? a->n_embedded : a->n_features) : a->n_out[l - 1];
        TFEPB_CHECK_ARG(a->ldw[l] >= width, "layer %d: leading dimension smaller than the input width", l);
        p.act_ld[l] = padded_ld<T>(a->ldw[l] > width ? a->ldw[l] : width);
    }
    for (int l = 0; l < a->n_linear; ++l) {
        p.extra[l] = (const T*)a->extra[l]; p.ldextra[l] = a->ldextra[l];
        p.act_out[l] = (T*)a->act_out[l]; p.ldact_out[l] = a->ldact_out[l];
        p.n_hidden[l] = l == 0 ? 0 : a->n_out[l - 1];
        TFEPB_CHECK_ARG(a->extra[l] == nullptr || a->ldextra[l] >= a->n_out[l], "layer %d: leading dimension of `extra`", l);
        TFEPB_CHECK_ARG(a->act_out[l] == nullptr || (l > 0 && a->ldact_out[l] >= a->n_out[l - 1]),
                        "layer %d: act_out receives hidden layer l (l >= 1) with a leading dimension >= its width", l);
    }
    p.par_ld = a->max_params < 1 ? 1 : a->max_params;
    p.groups = a->groups; p.n_groups = a->n_groups; p.parts = a->parts; p.gparts = a->group_parts; p.ids = a->ids;
    p.fixed_cols = a->fixed_cols; p.n_fixed = a->n_fixed;
    p.emb_out_col = a->emb_out_col; p.emb_periodic = a->emb_periodic;
    p.emb_lower = (T)a->emb_lower; p.emb_scale = (T)a->emb_scale;
    p.E = a->emb_out_col ? a->n_embedded : a->n_features;
    p.xs_ld = a->emb_out_col ? a->n_features : p.act_ld[0];
    // Tile shapes, in order of preference: (a) 32 samples with the weights of two degree groups staged in shared
    // memory; (b) two independent CTAs of 128 threads x 32 samples per SM (while one is in the single-warp
    // transformer stage or at a barrier the other one multiplies); (c) one CTA of 256 threads with 64 or 32 samples.
    const size_t limit = 227 * 1024, half = 113 * 1024;
    const size_t s32x128 = smem_bytes(p, 32, 128), s64 = smem_bytes(p, 64, 256), s32 = smem_bytes(p, 32, 256);
    TFEPB_CHECK_ARG(s32 <= limit, "the activations of 32 samples (%zu bytes) exceed the shared memory of an SM", s32);
    int TS, per_sm, threads = 256;
    size_t smem;
    void (*kernel)(const Params<T>);
    p.stage_elems = (a->max_group_weight_elems + 31) / 32 * 32;
    const size_t staged = s32 + 256 + 2 * (size_t)p.stage_elems * sizeof(T) + 16;
    if (a->max_group_weight_elems > 0 && staged <= limit) {
        kernel = maf_inverse_sweep_staged_kernel<T, 32, 256>; TS = 32; smem = staged; per_sm = 1;
    } else if (s32x128 <= half) { kernel = maf_inverse_sweep_kernel<T, 32, 128>; TS = 32; smem = s32x128; per_sm = 2; threads = 128; }
    else if (s64 <= limit) { kernel = maf_inverse_sweep_kernel<T, 64, 256>; TS = 64; smem = s64; per_sm = 1; }
    else { kernel = maf_inverse_sweep_kernel<T, 32, 256>; TS = 32; smem = s32; per_sm = 1; }
    TFEPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_tiles = (a->batch + TS - 1) / TS;
    const int cap = sm_count() * per_sm;
    const int grid = n_tiles < cap ? n_tiles : cap;
    kernel<<<grid, threads, smem, stream>>>(p);
    return check_launch("maf_inverse_sweep_kernel");
}

}  // namespace sweep
}  // namespace tfepb

using namespace tfepb;

extern "C" int tfepb_maf_inverse_sweep(const tfepb_sweep_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->dtype == TFEPB_F32 || a->dtype == TFEPB_F64, "unknown dtype %d", a->dtype);
    TFEPB_CHECK_ARG(a->batch >= 0 && a->n_features > 0, "bad sizes");
    TFEPB_CHECK_ARG(a->n_linear >= 1 && a->n_linear <= sweep::MAXL, "n_linear must be in [1, %d]", sweep::MAXL);
    TFEPB_CHECK_ARG(a->y && a->x && a->logdet, "null buffer");
    TFEPB_CHECK_ARG(a->n_groups >= 0 && (a->n_groups == 0 || (a->groups && a->parts && a->group_parts && a->ids)),
                    "null schedule table");
    TFEPB_CHECK_ARG(a->n_fixed == 0 || a->fixed_cols != nullptr, "null fixed_cols");
    TFEPB_CHECK_ARG(a->emb_out_col == nullptr || (a->emb_periodic != nullptr && a->n_embedded >= a->n_features),
                    "embedding tables: emb_periodic and n_embedded >= n_features are required with emb_out_col");
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    if (a->dtype == TFEPB_F32) return sweep::launch<float>(a, as_stream(stream));
    return sweep::launch<double>(a, as_stream(stream));
}
