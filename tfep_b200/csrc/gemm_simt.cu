// Exact-arithmetic (FFMA / DFMA) masked linear layers of the MADE conditioner.
//
// This is the any-shape, fp32/fp64 path: it is what the 1e-5 parity gate against the reference's
// fp32 CPU results is measured on, and what the backward pass and the degree-ordered inverse sweep
// are built from.  The tensor-core path for the headline configuration lives in maf_fused_sm100.cu.
//
// One kernel template covers the three products of nn/masked.py:266-302
//     forward        Y  = act(X W^T + b)                A = X  (k contiguous), B = W (k contiguous)
//     backward input dX = (dY W) * ELU'(H_prev)         A = dY (k contiguous), B = W (n contiguous)
//     backward weight dW += dY^T X, db += sum_b dY      A = dY (m contiguous), B = X (n contiguous)
// with 128x64x16 tiles, 8x4 register micro-tiles, register-staged global loads overlapped with the
// FMA loop, and an optional per-N-tile reduction range [begin, end) so that the all-zero part of a
// degree-sorted (staircase) masked weight is never read or multiplied.
#include "common.cuh"

namespace tfepb {

namespace {

constexpr int BM = 128, BN = TFEPB_GEMM_TILE_N, BK = 16, THREADS = 256;
constexpr int TM = 8, TN = 4, PAD = 4;
static_assert(BN == 64, "tile mapping below assumes 64 output columns per CTA");

template <typename T>
struct GemmParams {
    const T* A; int64_t lda;
    const T* B; int64_t ldb;
    T* C; int64_t ldc;
    int M, N, K;
    const T* bias;          // (N,) or null
    int act;                // TFEPB_ACT_*
    const T* aux; int64_t ldaux;   // ELU'(aux) multiplier or null
    int accumulate;         // C += instead of C =
    int atomic;             // split-K: atomicAdd into C
    const int* ranges;      // per N-tile [begin, end) of the reduction, or null
    int k_chunk;            // split-K chunk length (blockIdx.z), 0 = no split
    T* row_sums;            // (M,) += sum_k A[m, k] (bias gradient), or null; only N-tile 0 contributes
};

template <typename T, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS) gemm_kernel(GemmParams<T> p) {
    __shared__ __align__(16) T As[BK][BM + PAD];
    __shared__ __align__(16) T Bs[BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    int kb = 0, ke = p.K;
    if (p.ranges != nullptr) {
        kb = max(0, p.ranges[2 * blockIdx.x]);
        ke = min(p.K, p.ranges[2 * blockIdx.x + 1]);
    }
    if (p.k_chunk > 0) {
        kb = max(kb, (int)blockIdx.z * p.k_chunk);
        ke = min(ke, ((int)blockIdx.z + 1) * p.k_chunk);
    }

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = T(0);
    T rsum[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) rsum[i] = T(0);
    const bool want_rsum = (p.row_sums != nullptr) && (blockIdx.x == 0) && (tx == 0);

    constexpr int A_PER = BM * BK / THREADS;   // 8
    constexpr int B_PER = BN * BK / THREADS;   // 4
    T ra[A_PER], rb[B_PER];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int idx = tid + i * THREADS;
            const int kk = A_KC ? (idx % BK) : (idx / BM);
            const int mm = A_KC ? (idx / BK) : (idx % BM);
            const int gm = m0 + mm, gk = k0 + kk;
            T v = T(0);
            if (gm < p.M && gk < ke) v = A_KC ? p.A[(int64_t)gm * p.lda + gk] : p.A[(int64_t)gk * p.lda + gm];
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int idx = tid + i * THREADS;
            const int kk = B_KC ? (idx % BK) : (idx / BN);
            const int nn = B_KC ? (idx / BK) : (idx % BN);
            const int gn = n0 + nn, gk = k0 + kk;
            T v = T(0);
            if (gn < p.N && gk < ke) v = B_KC ? p.B[(int64_t)gn * p.ldb + gk] : p.B[(int64_t)gk * p.ldb + gn];
            rb[i] = v;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int idx = tid + i * THREADS;
            const int kk = A_KC ? (idx % BK) : (idx / BM);
            const int mm = A_KC ? (idx / BK) : (idx % BM);
            As[kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int idx = tid + i * THREADS;
            const int kk = B_KC ? (idx % BK) : (idx / BN);
            const int nn = B_KC ? (idx / BK) : (idx % BN);
            Bs[kk][nn] = rb[i];
        }
    };

    if (kb < ke) load_tiles(kb);
    for (int k0 = kb; k0 < ke; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < ke) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            if (want_rsum) {
#pragma unroll
                for (int i = 0; i < TM; ++i) rsum[i] += a[i];
            }
        }
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
            if (gn >= p.N) continue;
            T v = acc[i][j];
            T* c = p.C + (int64_t)gm * p.ldc + gn;
            if (p.atomic) {
                atomicAdd(c, v);
                continue;
            }
            if (p.bias != nullptr) v += p.bias[gn];
            if (p.act == TFEPB_ACT_ELU) v = elu(v);
            if (p.aux != nullptr) {
                const T h = p.aux[(int64_t)gm * p.ldaux + gn];
                v *= (h > T(0)) ? T(1) : (h + T(1));
            }
            if (p.accumulate) v += *c;
            *c = v;
        }
        if (want_rsum) atomicAdd(p.row_sums + gm, rsum[i]);
    }
}

template <typename T, bool A_KC, bool B_KC>
int launch(const GemmParams<T>& p, int splits, cudaStream_t stream, const char* what) {
    if (p.M <= 0 || p.N <= 0) return 0;
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits);
    gemm_kernel<T, A_KC, B_KC><<<grid, THREADS, 0, stream>>>(p);
    return check_launch(what);
}

template <typename T>
int forward_t(const tfepb_linear_fwd_args* a, cudaStream_t s) {
    GemmParams<T> p{};
    p.A = (const T*)a->x; p.lda = a->ldx;
    p.B = (const T*)a->w; p.ldb = a->ldw;
    p.C = (T*)a->y; p.ldc = a->ldy;
    p.M = a->batch; p.N = a->out_features; p.K = a->in_features;
    p.bias = (const T*)a->bias; p.act = a->activation;
    p.ranges = a->k_ranges;
    return launch<T, true, true>(p, 1, s, "masked_linear_forward");
}

template <typename T>
int bwd_input_t(const tfepb_linear_bwd_input_args* a, cudaStream_t s) {
    GemmParams<T> p{};
    p.A = (const T*)a->grad_y; p.lda = a->ldgy;       // (batch, out): reduction over out
    p.B = (const T*)a->w; p.ldb = a->ldw;             // element (n = in, k = out) at w[k * ldw + n]
    p.C = (T*)a->grad_x; p.ldc = a->ldgx;
    p.M = a->batch; p.N = a->in_features; p.K = a->out_features;
    p.aux = (const T*)a->act_out; p.ldaux = a->ldact;
    p.accumulate = a->accumulate;
    p.ranges = a->n_ranges;
    return launch<T, true, false>(p, 1, s, "masked_linear_backward_input");
}

template <typename T>
int bwd_weight_t(const tfepb_linear_bwd_weight_args* a, cudaStream_t s) {
    GemmParams<T> p{};
    p.A = (const T*)a->grad_y; p.lda = a->ldgy;       // element (m = out, k = batch) at gy[k * ld + m]
    p.B = (const T*)a->x; p.ldb = a->ldx;             // element (n = in,  k = batch) at x[k * ld + n]
    p.C = (T*)a->grad_w; p.ldc = a->ldgw;
    p.M = a->out_features; p.N = a->in_features; p.K = a->batch;
    p.atomic = 1;
    p.row_sums = (T*)a->grad_bias;
    // split the batch reduction so that the grid covers the machine a few times over
    const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    int splits = (4 * sm_count() + tiles - 1) / tiles;
    const int max_splits = (p.K + 4 * BK - 1) / (4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int chunk = (p.K + splits - 1) / splits;
    chunk = ((chunk + BK - 1) / BK) * BK;
    splits = (p.K + chunk - 1) / chunk;
    p.k_chunk = chunk;
    return launch<T, false, false>(p, splits, s, "masked_linear_backward_weight");
}

}  // namespace
}  // namespace tfepb

using namespace tfepb;

extern "C" int tfepb_masked_linear_forward(const tfepb_linear_fwd_args* a, tfepb_stream_t stream) {
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->in_features >= 0 && a->out_features > 0, "bad sizes");
    TFEPB_CHECK_ARG((a->in_features == 0 || (a->x && a->w)) && a->y, "null buffer");
    TFEPB_CHECK_ARG(a->ldx >= a->in_features && a->ldw >= a->in_features && a->ldy >= a->out_features,
                    "leading dimension smaller than the row length");
    if (int rc = require_sm100()) return rc;
    if (a->dtype == TFEPB_F32) return forward_t<float>(a, as_stream(stream));
    if (a->dtype == TFEPB_F64) return forward_t<double>(a, as_stream(stream));
    return fail(-1, "unknown dtype %d", a->dtype);
}

extern "C" int tfepb_masked_linear_backward_input(const tfepb_linear_bwd_input_args* a, tfepb_stream_t stream) {
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->in_features > 0 && a->out_features > 0, "bad sizes");
    TFEPB_CHECK_ARG(a->grad_y && a->w && a->grad_x, "null buffer");
    TFEPB_CHECK_ARG(a->ldgy >= a->out_features && a->ldw >= a->in_features && a->ldgx >= a->in_features,
                    "leading dimension smaller than the row length");
    if (int rc = require_sm100()) return rc;
    if (a->dtype == TFEPB_F32) return bwd_input_t<float>(a, as_stream(stream));
    if (a->dtype == TFEPB_F64) return bwd_input_t<double>(a, as_stream(stream));
    return fail(-1, "unknown dtype %d", a->dtype);
}

extern "C" int tfepb_masked_linear_backward_weight(const tfepb_linear_bwd_weight_args* a, tfepb_stream_t stream) {
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->in_features > 0 && a->out_features > 0, "bad sizes");
    TFEPB_CHECK_ARG(a->grad_y && a->x && a->grad_w, "null buffer");
    TFEPB_CHECK_ARG(a->ldgy >= a->out_features && a->ldx >= a->in_features && a->ldgw >= a->in_features,
                    "leading dimension smaller than the row length");
    if (int rc = require_sm100()) return rc;
    if (a->dtype == TFEPB_F32) return bwd_weight_t<float>(a, as_stream(stream));
    if (a->dtype == TFEPB_F64) return bwd_weight_t<double>(a, as_stream(stream));
    return fail(-1, "unknown dtype %d", a->dtype);
}
