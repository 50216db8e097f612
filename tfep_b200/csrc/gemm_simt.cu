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
    const int* skip_ranges; // per 64-column tile of C: [begin, end) rows that can be non-zero (masked weight gradient), or null
};

template <typename T, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS) gemm_kernel(GemmParams<T> p) {
    __shared__ __align__(16) T As[BK][BM + PAD];
    __shared__ __align__(16) T Bs[BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (p.skip_ranges != nullptr && blockIdx.x != 0) {      // tile of the weight gradient that the mask zeroes anyway
        const int rb = p.skip_ranges[2 * blockIdx.x], re = p.skip_ranges[2 * blockIdx.x + 1];
        if (m0 + BM <= rb || m0 >= re) return;
    }

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

// ---------------------------------------------------------------------------------------------------------
// fp32 fast path: 128 x 128 x 16 tiles, 8 x 8 register micro-tiles (64 FFMA per four 16-byte shared-memory loads),
// 16-byte global loads, double-buffered shared memory (one CTA barrier per k-tile).  Same operands, reduction
// ranges, split-K and epilogue semantics as gemm_kernel; requires 16-byte aligned operands (leading dimensions
// multiples of 4, checked by the launcher, which otherwise falls back to gemm_kernel).
// ---------------------------------------------------------------------------------------------------------
constexpr int BM2 = 128, BN2 = 128, LD2 = BM2 + 4;

template <bool KC>
__device__ __forceinline__ void load_tile128(const float* __restrict__ G, int64_t ld, int rows0, int rows, int k0, int K,
                                             int tid, float4 (&r)[2]) {
    // KC: element (row, k) at G[row * ld + k] -> 16-byte loads along k; else element at G[k * ld + row] -> along rows
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int q = tid + i * THREADS;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KC) {
            const int row = rows0 + q / 4, k = k0 + (q % 4) * 4;
            if (row < rows) {
                const float* src = G + (int64_t)row * ld + k;
                if (k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src));
                else {
                    if (k < K) v.x = __ldg(src);
                    if (k + 1 < K) v.y = __ldg(src + 1);
                    if (k + 2 < K) v.z = __ldg(src + 2);
                }
            }
        } else {
            const int k = k0 + q / 32, row = rows0 + (q % 32) * 4;
            if (k < K) {
                const float* src = G + (int64_t)k * ld + row;
                if (row + 3 < rows) v = __ldg(reinterpret_cast<const float4*>(src));
                else {
                    if (row < rows) v.x = __ldg(src);
                    if (row + 1 < rows) v.y = __ldg(src + 1);
                    if (row + 2 < rows) v.z = __ldg(src + 2);
                }
            }
        }
        r[i] = v;
    }
}

template <bool KC>
__device__ __forceinline__ void store_tile128(float (*S)[LD2], int tid, const float4 (&r)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int q = tid + i * THREADS;
        if (KC) {
            const int row = q / 4, k = (q % 4) * 4;
            S[k][row] = r[i].x; S[k + 1][row] = r[i].y; S[k + 2][row] = r[i].z; S[k + 3][row] = r[i].w;
        } else {
            const int k = q / 32, row = (q % 32) * 4;
            *reinterpret_cast<float4*>(&S[k][row]) = r[i];
        }
    }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 2) gemm128_kernel(GemmParams<float> p) {
    __shared__ __align__(16) float As[2][BK][LD2];
    __shared__ __align__(16) float Bs[2][BK][LD2];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * BM2, n0 = blockIdx.x * BN2;
    if (p.skip_ranges != nullptr && blockIdx.x != 0) {      // tile of the weight gradient that the mask zeroes anyway
        int rb = p.skip_ranges[4 * blockIdx.x], re = p.skip_ranges[4 * blockIdx.x + 1];
        if (n0 + 64 < p.N) {
            const int rb1 = p.skip_ranges[4 * blockIdx.x + 2], re1 = p.skip_ranges[4 * blockIdx.x + 3];
            if (re1 > rb1) {
                if (re > rb) { rb = min(rb, rb1); re = max(re, re1); }
                else { rb = rb1; re = re1; }
            }
        }
        if (m0 + BM2 <= rb || m0 >= re) return;
    }

    int kb = 0, ke = p.K;
    if (p.ranges != nullptr) {             // ranges are per 64-column tile: take the union of the two halves
        const int t0 = 2 * blockIdx.x, t1 = t0 + 1;
        int b0 = p.ranges[2 * t0], e0 = p.ranges[2 * t0 + 1];
        if (n0 + 64 < p.N) {
            const int b1 = p.ranges[2 * t1], e1 = p.ranges[2 * t1 + 1];
            if (e1 > b1) {
                if (e0 > b0) { b0 = min(b0, b1); e0 = max(e0, e1); }
                else { b0 = b1; e0 = e1; }
            }
        }
        kb = max(0, b0) & ~3;              // 16-byte aligned start: the extra columns multiply masked (zero) weights
        ke = min(p.K, e0);
    }
    if (p.k_chunk > 0) {
        kb = max(kb, (int)blockIdx.z * p.k_chunk);
        ke = min(ke, ((int)blockIdx.z + 1) * p.k_chunk);
    }

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float rsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rsum[i] = 0.f;
    const bool want_rsum = (p.row_sums != nullptr) && (blockIdx.x == 0) && (tx == 0);

    float4 ra[2], rb[2];
    int buf = 0;
    if (kb < ke) {
        load_tile128<A_KC>(p.A, p.lda, m0, p.M, kb, ke, tid, ra);
        load_tile128<B_KC>(p.B, p.ldb, n0, p.N, kb, ke, tid, rb);
        store_tile128<A_KC>(As[0], tid, ra);
        store_tile128<B_KC>(Bs[0], tid, rb);
    }
    __syncthreads();
    for (int k0 = kb; k0 < ke; k0 += BK) {
        const bool more = k0 + BK < ke;
        if (more) {
            load_tile128<A_KC>(p.A, p.lda, m0, p.M, k0 + BK, ke, tid, ra);
            load_tile128<B_KC>(p.B, p.ldb, n0, p.N, k0 + BK, ke, tid, rb);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            if (want_rsum) {
#pragma unroll
                for (int i = 0; i < 8; ++i) rsum[i] += a[i];
            }
        }
        if (more) {
            store_tile128<A_KC>(As[buf ^ 1], tid, ra);
            store_tile128<B_KC>(Bs[buf ^ 1], tid, rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    // epilogue: rows ty * 4 + {0..3} and 64 + ty * 4 + {0..3}; columns tx * 4 + {0..3} and 64 + tx * 4 + {0..3}
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn >= p.N) continue;
            float v = acc[i][j];
            float* c = p.C + (int64_t)gm * p.ldc + gn;
            if (p.atomic) {
                atomicAdd(c, v);
                continue;
            }
            if (p.bias != nullptr) v += p.bias[gn];
            if (p.act == TFEPB_ACT_ELU) v = elu(v);
            if (p.aux != nullptr) {
                const float h = p.aux[(int64_t)gm * p.ldaux + gn];
                v *= (h > 0.f) ? 1.f : (h + 1.f);
            }
            if (p.accumulate) v += *c;
            *c = v;
        }
        if (want_rsum) atomicAdd(p.row_sums + gm, rsum[i]);
    }
}

template <typename T>
bool aligned16(const GemmParams<T>&) { return false; }
template <>
bool aligned16<float>(const GemmParams<float>& p) {
    return p.N > 64 && p.lda % 4 == 0 && p.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(p.B) & 15) == 0 && (p.k_chunk % 4 == 0);
}

template <typename T, bool A_KC, bool B_KC>
int launch128(const GemmParams<T>&, int, cudaStream_t, const char*) { return -1; }
template <>
int launch128<float, true, true>(const GemmParams<float>& p, int splits, cudaStream_t stream, const char* what) {
    dim3 grid((p.N + BN2 - 1) / BN2, (p.M + BM2 - 1) / BM2, splits);
    gemm128_kernel<true, true><<<grid, THREADS, 0, stream>>>(p);
    return check_launch(what);
}
template <>
int launch128<float, true, false>(const GemmParams<float>& p, int splits, cudaStream_t stream, const char* what) {
    dim3 grid((p.N + BN2 - 1) / BN2, (p.M + BM2 - 1) / BM2, splits);
    gemm128_kernel<true, false><<<grid, THREADS, 0, stream>>>(p);
    return check_launch(what);
}
template <>
int launch128<float, false, false>(const GemmParams<float>& p, int splits, cudaStream_t stream, const char* what) {
    dim3 grid((p.N + BN2 - 1) / BN2, (p.M + BM2 - 1) / BM2, splits);
    gemm128_kernel<false, false><<<grid, THREADS, 0, stream>>>(p);
    return check_launch(what);
}

template <typename T, bool A_KC, bool B_KC>
int launch(const GemmParams<T>& p, int splits, cudaStream_t stream, const char* what) {
    if (p.M <= 0 || p.N <= 0) return 0;
    if (aligned16<T>(p)) return launch128<T, A_KC, B_KC>(p, splits, stream, what);
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
    p.skip_ranges = a->n_ranges;
    // split the batch reduction so that the grid covers the machine a few times over
    const int bn = (sizeof(T) == 4 && p.N > 64) ? BN2 : BN;       // tile width of the kernel the launcher will pick
    const int tiles = ((p.M + BM - 1) / BM) * ((p.N + bn - 1) / bn);
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
    TFEPB_NVTX();
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
    TFEPB_NVTX();
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
    TFEPB_NVTX();
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
