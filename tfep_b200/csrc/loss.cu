// BoltzmannKLDivLoss (reference tfep/loss.py:76-140) as ONE streaming reduction, plus its backward.
//
//   reduced work   rw_i = u_B,i - log|det J|_i - u_A,i                         (loss.py:125-129)
//   unweighted     loss = mean_i rw_i          (ignore_nan: torch.nanmean)      (loss.py:138-140)
//   weighted       loss = sum_i softmax(log_w)_i rw_i   (ignore_nan: nansum)    (loss.py:132-136)
//
// The reference materialises rw, softmax(log_w) and their product as (batch,) temporaries in separate kernels; here
// every input vector is read once (4-16 bytes per sample, HBM bound) and the softmax is folded into the same pass
// as an online (max, sum e, sum e rw) triple -- the (T)FEP estimator's log-sum-exp pattern (analysis.cu) with a
// weighted numerator.  Two launches: per-block partials, then one block that combines them in block order
// (deterministic), both carried in double.  The second kernel also leaves (max, sum e, count, loss) for the backward
// pass, which is one elementwise kernel writing all four cotangents.
#include "common.cuh"

namespace tfepb {
namespace {

constexpr int KL_THREADS = 256;
constexpr int KL_MAX_BLOCKS = 1184;         // 8 blocks per SM on 148 SMs

struct Part {
    double m, s, t, cnt;                    // max of log_w, sum exp(log_w - m), sum exp(log_w - m) rw, terms kept
    double bad;                             // 1 if a log-weight was NaN (softmax is NaN everywhere then)
};

__device__ __forceinline__ Part part_combine(const Part& a, const Part& b) {
    Part r;
    r.bad = (a.bad != 0.0 || b.bad != 0.0) ? 1.0 : 0.0;
    r.cnt = a.cnt + b.cnt;
    if (b.s == 0.0 && b.cnt == 0.0 && b.m == -INFINITY) { r.m = a.m; r.s = a.s; r.t = a.t; return r; }
    if (a.s == 0.0 && a.cnt == 0.0 && a.m == -INFINITY) { r.m = b.m; r.s = b.s; r.t = b.t; return r; }
    r.m = a.m > b.m ? a.m : b.m;
    const double fa = a.m == r.m ? 1.0 : exp(a.m - r.m), fb = b.m == r.m ? 1.0 : exp(b.m - r.m);
    r.s = a.s * fa + b.s * fb;
    r.t = a.t * fa + b.t * fb;
    return r;
}

__device__ __forceinline__ Part part_shfl(const Part& v, int o) {
    Part u;
    u.m = __shfl_xor_sync(0xffffffffu, v.m, o); u.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    u.t = __shfl_xor_sync(0xffffffffu, v.t, o); u.cnt = __shfl_xor_sync(0xffffffffu, v.cnt, o);
    u.bad = __shfl_xor_sync(0xffffffffu, v.bad, o);
    return u;
}

__device__ __forceinline__ Part part_block_reduce(Part v) {
    __shared__ Part sh[KL_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = part_combine(v, part_shfl(v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < KL_THREADS / 32 ? sh[lane] : Part{-INFINITY, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v = part_combine(v, part_shfl(v, o));
    }
    return v;       // valid in thread 0
}

// One thread keeps a running (m, s, t) in the input precision domain rescaled on a new maximum, the sums in double.
template <typename T>
__global__ void __launch_bounds__(KL_THREADS) kl_partial_kernel(const T* __restrict__ ub, const T* __restrict__ ld,
                                                                const T* __restrict__ ua, const T* __restrict__ lw, int64_t n,
                                                                int ignore_nan, Part* __restrict__ partials) {
    Part acc{-INFINITY, 0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * KL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * KL_THREADS + threadIdx.x; i < n; i += stride) {
        T rw = ub[i];
        if (ld != nullptr) rw -= ld[i];
        if (ua != nullptr) rw -= ua[i];
        const bool rw_nan = rw != rw;
        if (lw == nullptr) {
            // unweighted: every kept term has weight 1 (m stays 0 so that partials combine without rescaling)
            acc.m = 0.0;
            if (!(ignore_nan && rw_nan)) { acc.t += (double)rw; acc.cnt += 1.0; }
            acc.s += 1.0;
        } else {
            const double w = (double)lw[i];
            if (w != w) { acc.bad = 1.0; continue; }
            if (w > acc.m) {
                const double f = acc.m == -INFINITY ? 0.0 : exp(acc.m - w);
                acc.s *= f; acc.t *= f;
                acc.m = w;
            }
            const double e = w == acc.m ? 1.0 : exp(w - acc.m);
            acc.s += e;
            if (!(ignore_nan && rw_nan)) { acc.t += e * (double)rw; acc.cnt += 1.0; }
        }
    }
    acc = part_block_reduce(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// out[0] = loss, out[1] = max log_w (0 if unweighted), out[2] = sum exp(log_w - max) (n if unweighted),
// out[3] = number of kept terms, out[4] = 1 if a log-weight was NaN
__global__ void __launch_bounds__(KL_THREADS) kl_final_kernel(const Part* __restrict__ partials, int n_blocks, int weighted,
                                                              int ignore_nan, double* __restrict__ out) {
    // block order: thread k owns a contiguous run of partials, runs are combined by the fixed tree of the block reduce
    Part acc{-INFINITY, 0.0, 0.0, 0.0, 0.0};
    const int per = (n_blocks + KL_THREADS - 1) / KL_THREADS;
    for (int j = 0; j < per; ++j) {
        const int b = threadIdx.x * per + j;
        if (b < n_blocks) acc = part_combine(acc, partials[b]);
    }
    acc = part_block_reduce(acc);
    if (threadIdx.x == 0) {
        double loss;
        if (weighted) {
            if (acc.bad != 0.0) loss = ignore_nan ? 0.0 : NAN;      // softmax of a vector holding a NaN is all NaN
            else loss = acc.t / acc.s;
        } else {
            loss = acc.cnt > 0.0 ? acc.t / acc.cnt : NAN;           // mean over the kept terms (nanmean of all-NaN is NaN)
        }
        out[0] = loss; out[1] = acc.m; out[2] = acc.s; out[3] = acc.cnt; out[4] = acc.bad;
    }
}

template <typename T>
__global__ void __launch_bounds__(KL_THREADS) kl_backward_kernel(const T* __restrict__ ub, const T* __restrict__ ld,
                                                                 const T* __restrict__ ua, const T* __restrict__ lw, int64_t n,
                                                                 int ignore_nan, const double* __restrict__ stats,
                                                                 const T* __restrict__ grad_out, T* __restrict__ g_ub,
                                                                 T* __restrict__ g_ld, T* __restrict__ g_ua, T* __restrict__ g_lw) {
    const double go = (double)grad_out[0];
    const double loss = stats[0], m = stats[1], s = stats[2], cnt = stats[3], bad = stats[4];
    const int64_t stride = (int64_t)gridDim.x * KL_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * KL_THREADS + threadIdx.x; i < n; i += stride) {
        T rw = ub[i];
        if (ld != nullptr) rw -= ld[i];
        if (ua != nullptr) rw -= ua[i];
        const bool dropped = ignore_nan && (rw != rw);
        double g, gw = 0.0;
        if (lw == nullptr) {
            g = dropped ? 0.0 : go / cnt;
        } else if (bad != 0.0) {
            g = ignore_nan ? 0.0 : NAN;
            gw = g;
        } else {
            const double w = exp((double)lw[i] - m) / s;
            g = dropped ? 0.0 : go * w;
            gw = go * w * ((dropped ? 0.0 : (double)rw) - loss);
        }
        if (g_ub != nullptr) g_ub[i] = (T)g;
        if (g_ld != nullptr) g_ld[i] = (T)(-g);
        if (g_ua != nullptr) g_ua[i] = (T)(-g);
        if (g_lw != nullptr) g_lw[i] = (T)gw;
    }
}

int kl_blocks(int64_t n) {
    int64_t blocks = (n + (int64_t)KL_THREADS * 8 - 1) / ((int64_t)KL_THREADS * 8);
    const int64_t cap = (int64_t)sm_count() * 8 < KL_MAX_BLOCKS ? (int64_t)sm_count() * 8 : KL_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace
}  // namespace tfepb

using namespace tfepb;

extern "C" int64_t tfepb_kl_loss_workspace_bytes(void) { return (int64_t)KL_MAX_BLOCKS * sizeof(Part); }

extern "C" int tfepb_kl_loss(int32_t dtype, const void* target_potentials, const void* log_det_J, const void* ref_potentials,
                             const void* log_weights, int64_t n, int32_t ignore_nan, void* workspace, double* out5,
                             tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(n > 0, "empty batch");
    TFEPB_CHECK_ARG(target_potentials && workspace && out5, "null buffer");
    if (int rc = require_sm100()) return rc;
    const int blocks = kl_blocks(n);
    cudaStream_t s = as_stream(stream);
    if (dtype == TFEPB_F32)
        kl_partial_kernel<float><<<blocks, KL_THREADS, 0, s>>>((const float*)target_potentials, (const float*)log_det_J,
                                                              (const float*)ref_potentials, (const float*)log_weights, n,
                                                              ignore_nan, (Part*)workspace);
    else if (dtype == TFEPB_F64)
        kl_partial_kernel<double><<<blocks, KL_THREADS, 0, s>>>((const double*)target_potentials, (const double*)log_det_J,
                                                               (const double*)ref_potentials, (const double*)log_weights, n,
                                                               ignore_nan, (Part*)workspace);
    else
        return fail(-1, "unknown dtype %d", dtype);
    if (int rc = check_launch("kl_partial")) return rc;
    kl_final_kernel<<<1, KL_THREADS, 0, s>>>((const Part*)workspace, blocks, log_weights != nullptr, ignore_nan, out5);
    return check_launch("kl_final");
}

extern "C" int tfepb_kl_loss_backward(int32_t dtype, const void* target_potentials, const void* log_det_J,
                                      const void* ref_potentials, const void* log_weights, int64_t n, int32_t ignore_nan,
                                      const double* stats5, const void* grad_out, void* grad_target, void* grad_log_det_J,
                                      void* grad_ref, void* grad_log_weights, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(n > 0, "empty batch");
    TFEPB_CHECK_ARG(target_potentials && stats5 && grad_out, "null buffer");
    if (int rc = require_sm100()) return rc;
    const int blocks = kl_blocks(n);
    cudaStream_t s = as_stream(stream);
    if (dtype == TFEPB_F32)
        kl_backward_kernel<float><<<blocks, KL_THREADS, 0, s>>>(
            (const float*)target_potentials, (const float*)log_det_J, (const float*)ref_potentials, (const float*)log_weights, n,
            ignore_nan, stats5, (const float*)grad_out, (float*)grad_target, (float*)grad_log_det_J, (float*)grad_ref,
            (float*)grad_log_weights);
    else if (dtype == TFEPB_F64)
        kl_backward_kernel<double><<<blocks, KL_THREADS, 0, s>>>(
            (const double*)target_potentials, (const double*)log_det_J, (const double*)ref_potentials, (const double*)log_weights, n,
            ignore_nan, stats5, (const double*)grad_out, (double*)grad_target, (double*)grad_log_det_J, (double*)grad_ref,
            (double*)grad_log_weights);
    else
        return fail(-1, "unknown dtype %d", dtype);
    return check_launch("kl_backward");
}
