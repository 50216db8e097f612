// Stand-alone transformer kernels (exact fp32 / fp64 arithmetic): elementwise map + per-sample
// log|det J| reduction, and their vector-Jacobian products.
//
// Mapping: one warp owns one sample at a time; lanes stride over the features (or Moebius vector
// blocks), so that for the reference's parameter-major layout (stride_f = 1) every parameter load
// and every x / y access of a warp is one contiguous segment.  The per-sample log-det is reduced
// with warp shuffles and written once.  These kernels are HBM-bound (parameters are read once).
#include "common.cuh"
#include "tx_math.cuh"
#include "tx_ops.cuh"

namespace tfepb {
namespace {

template <typename T>
TxView<T> make_view(const tfepb_tx_io* io, const tfepb_tx_grads* g = nullptr) {
    TxView<T> v{};
    v.x = (const T*)io->x; v.ldx = io->ldx;
    v.y = (T*)io->y; v.ldy = io->ldy;
    v.par = (const T*)io->par; v.ldp = io->ldp;
    v.poff = io->par_offset; v.sp = io->par_stride_p; v.sf = io->par_stride_f;
    v.pbase = io->par_base;
    v.cols = io->cols;
    v.ids = io->feat_ids;
    v.logdet = (T*)io->logdet;
    v.accumulate = io->accumulate_logdet;
    v.B = io->batch; v.F = io->n_features; v.inverse = io->inverse;
    if (g) {
        v.gy = (const T*)g->grad_y; v.ldgy = g->ldgy;
        v.gld = (const T*)g->grad_logdet;
        v.gx = (T*)g->grad_x; v.ldgx = g->ldgx;
        v.gpar = (T*)g->grad_par;
    }
    return v;
}

constexpr int TX_THREADS = 256;

inline int tx_blocks(int batch) {
    const int warps_per_block = TX_THREADS / 32;
    int64_t blocks = ((int64_t)batch + warps_per_block - 1) / warps_per_block;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

// Forward / inverse driver.  Op::units(F) = number of independent work items per sample,
// Op::apply(view, b, unit) performs the map for one item and returns its log-det contribution.
template <typename T, typename Op>
__global__ void __launch_bounds__(TX_THREADS) tx_forward_kernel(TxView<T> v, Op op) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * TX_THREADS + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * TX_THREADS) >> 5;
    const int units = op.units(v.F);
    for (int b = warp; b < v.B; b += nwarps) {
        T ld = T(0);
        for (int u = lane; u < units; u += 32) ld += op.apply(v, b, u);
        ld = warp_sum(ld);
        if (lane == 0 && v.logdet != nullptr) v.logdet[b] = v.accumulate ? v.logdet[b] + ld : ld;
    }
}

template <typename T, typename Op>
__global__ void __launch_bounds__(TX_THREADS) tx_backward_kernel(TxView<T> v, Op op) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * TX_THREADS + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * TX_THREADS) >> 5;
    const int units = op.units(v.F);
    for (int b = warp; b < v.B; b += nwarps) {
        const T gl = v.gld ? v.gld[b] : T(0);
        for (int u = lane; u < units; u += 32) op.backward(v, b, u, gl);
    }
}

template <typename T, typename Op>
int run(const tfepb_tx_io* io, const tfepb_tx_grads* g, Op op, cudaStream_t s, const char* what) {
    if (io->batch == 0 || io->n_features == 0) return 0;
    TxView<T> v = make_view<T>(io, g);
    if (g == nullptr) tx_forward_kernel<T, Op><<<tx_blocks(io->batch), TX_THREADS, 0, s>>>(v, op);
    else tx_backward_kernel<T, Op><<<tx_blocks(io->batch), TX_THREADS, 0, s>>>(v, op);
    return check_launch(what);
}

int check_io(const tfepb_tx_io* io, const tfepb_tx_grads* g) {
    TFEPB_CHECK_ARG(io != nullptr, "null io struct");
    TFEPB_CHECK_ARG(io->batch >= 0 && io->n_features >= 0, "bad sizes");
    TFEPB_CHECK_ARG(io->dtype == TFEPB_F32 || io->dtype == TFEPB_F64, "unknown dtype %d", io->dtype);
    TFEPB_CHECK_ARG(io->x && io->par, "null buffer");
    if (g == nullptr) {
        TFEPB_CHECK_ARG(io->y != nullptr, "null output buffer");
    } else {
        TFEPB_CHECK_ARG(io->inverse == 0, "vector-Jacobian products are implemented for the forward direction only");
        TFEPB_CHECK_ARG(g->grad_y && g->grad_x && g->grad_par, "null gradient buffer");
    }
    return require_sm100();
}

template <typename T>
int spline_dispatch(const tfepb_tx_io* io, const tfepb_spline_cfg* cfg, const tfepb_tx_grads* g, cudaStream_t s) {
    auto fill = [&](auto& op) {
        op.K = cfg->n_bins; op.circular = cfg->circular; op.idslopes = cfg->identity_boundary_slopes;
        op.learn_lo = cfg->learn_lower_bound; op.learn_hi = cfg->learn_upper_bound;
        op.x0 = (const T*)cfg->x0; op.xf = (const T*)cfg->xf; op.y0 = (const T*)cfg->y0; op.yf = (const T*)cfg->yf;
        op.min_bin = (T)cfg->min_bin_size; op.min_slope = (T)cfg->min_slope;
        op.bins = cfg->bins_out; op.ldbins = cfg->ldbins;
    };
    if (cfg->n_bins == 8 && io->par_stride_p == 1 && !cfg->identity_boundary_slopes && !cfg->learn_lower_bound &&
        !cfg->learn_upper_bound && cfg->bins_out == nullptr) {
        // packed layout of the MAF paths: compile-time bin count and parameter stride (same arithmetic)
        SplinePackedOp<T, 8> op;
        op.circular = cfg->circular;
        op.x0 = (const T*)cfg->x0; op.xf = (const T*)cfg->xf; op.y0 = (const T*)cfg->y0; op.yf = (const T*)cfg->yf;
        op.min_bin = (T)cfg->min_bin_size; op.min_slope = (T)cfg->min_slope;
        return run<T>(io, g, op, s, "spline_packed");
    }
    if (cfg->n_bins <= 8) { SplineOp<T, 8> op; fill(op); return run<T>(io, g, op, s, "spline"); }
    if (cfg->n_bins <= 16) { SplineOp<T, 16> op; fill(op); return run<T>(io, g, op, s, "spline"); }
    if (cfg->n_bins <= 32) { SplineOp<T, 32> op; fill(op); return run<T>(io, g, op, s, "spline"); }
    return fail(-1, "n_bins = %d exceeds the supported maximum of 32", cfg->n_bins);
}

int check_spline(const tfepb_tx_io* io, const tfepb_spline_cfg* cfg) {
    TFEPB_CHECK_ARG(cfg != nullptr, "null spline config");
    TFEPB_CHECK_ARG(cfg->n_bins >= 1, "n_bins must be positive");
    TFEPB_CHECK_ARG(cfg->x0 && cfg->xf && cfg->y0 && cfg->yf, "null spline domain buffer");
    TFEPB_CHECK_ARG(!(cfg->circular && (cfg->learn_lower_bound || cfg->learn_upper_bound)),
                    "Cannot instantiate a circular spline with learnable limits.");
    (void)io;
    return 0;
}

// ------------------------------------------------------------------------------------------
// PeriodicEmbedding (nn/embeddings/mafembed.py:112-142): one thread per input element; column c goes to
// out_col[c] (copy) or to the pair out_col[c], out_col[c] + 1 = cos, sin of (x - lower) * scale.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) periodic_embedding_kernel(const T* __restrict__ x, int64_t ldx, int batch, int n_in,
                                                                 const int* __restrict__ out_col,
                                                                 const int* __restrict__ periodic, T lower, T scale,
                                                                 T* __restrict__ out, int64_t ldo) {
    const int64_t total = (int64_t)batch * n_in;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / n_in), c = (int)(i - (int64_t)b * n_in);
        const T v = x[(int64_t)b * ldx + c];
        T* o = out + (int64_t)b * ldo + out_col[c];
        if (periodic[c]) {
            const T a = (v - lower) * scale;
            o[0] = cos(a);
            o[1] = sin(a);
        } else {
            o[0] = v;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) periodic_embedding_backward_kernel(const T* __restrict__ x, int64_t ldx, int batch,
                                                                          int n_in, const int* __restrict__ out_col,
                                                                          const int* __restrict__ periodic, T lower,
                                                                          T scale, const T* __restrict__ go, int64_t ldgo,
                                                                          T* __restrict__ gx, int64_t ldgx) {
    const int64_t total = (int64_t)batch * n_in;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / n_in), c = (int)(i - (int64_t)b * n_in);
        const T* g = go + (int64_t)b * ldgo + out_col[c];
        T r = g[0];
        if (periodic[c]) {
            const T a = (x[(int64_t)b * ldx + c] - lower) * scale;
            r = scale * (g[1] * cos(a) - g[0] * sin(a));
        }
        gx[(int64_t)b * ldgx + c] = r;
    }
}

}  // namespace
}  // namespace tfepb

using namespace tfepb;

extern "C" int tfepb_periodic_embedding(int32_t dtype, const void* x, int64_t ldx, int32_t batch, int32_t n_in,
                                        const int32_t* out_col, const int32_t* periodic, double lower, double scale,
                                        void* out, int64_t ldo, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(dtype == TFEPB_F32 || dtype == TFEPB_F64, "unknown dtype %d", dtype);
    TFEPB_CHECK_ARG(batch >= 0 && n_in > 0, "bad sizes");
    TFEPB_CHECK_ARG(x && out && out_col && periodic, "null buffer");
    if (int rc = require_sm100()) return rc;
    if (batch == 0) return 0;
    int64_t blocks = ((int64_t)batch * n_in + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == TFEPB_F32)
        periodic_embedding_kernel<float><<<(int)blocks, 256, 0, as_stream(stream)>>>(
            (const float*)x, ldx, batch, n_in, out_col, periodic, (float)lower, (float)scale, (float*)out, ldo);
    else
        periodic_embedding_kernel<double><<<(int)blocks, 256, 0, as_stream(stream)>>>(
            (const double*)x, ldx, batch, n_in, out_col, periodic, lower, scale, (double*)out, ldo);
    return check_launch("periodic_embedding");
}

extern "C" int tfepb_periodic_embedding_backward(int32_t dtype, const void* x, int64_t ldx, int32_t batch, int32_t n_in,
                                                 const int32_t* out_col, const int32_t* periodic, double lower, double scale,
                                                 const void* grad_out, int64_t ldgo, void* grad_x, int64_t ldgx,
                                                 tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(dtype == TFEPB_F32 || dtype == TFEPB_F64, "unknown dtype %d", dtype);
    TFEPB_CHECK_ARG(batch >= 0 && n_in > 0, "bad sizes");
    TFEPB_CHECK_ARG(x && grad_out && grad_x && out_col && periodic, "null buffer");
    if (int rc = require_sm100()) return rc;
    if (batch == 0) return 0;
    int64_t blocks = ((int64_t)batch * n_in + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == TFEPB_F32)
        periodic_embedding_backward_kernel<float><<<(int)blocks, 256, 0, as_stream(stream)>>>(
            (const float*)x, ldx, batch, n_in, out_col, periodic, (float)lower, (float)scale, (const float*)grad_out, ldgo,
            (float*)grad_x, ldgx);
    else
        periodic_embedding_backward_kernel<double><<<(int)blocks, 256, 0, as_stream(stream)>>>(
            (const double*)x, ldx, batch, n_in, out_col, periodic, lower, scale, (const double*)grad_out, ldgo,
            (double*)grad_x, ldgx);
    return check_launch("periodic_embedding_backward");
}

extern "C" int tfepb_affine(const tfepb_tx_io* io, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_io(io, nullptr)) return rc;
    if (io->dtype == TFEPB_F32) return run<float>(io, nullptr, AffineOp<float>{}, as_stream(stream), "affine");
    return run<double>(io, nullptr, AffineOp<double>{}, as_stream(stream), "affine");
}

extern "C" int tfepb_affine_backward(const tfepb_tx_io* io, const tfepb_tx_grads* g, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(g != nullptr, "null gradient struct");
    if (int rc = check_io(io, g)) return rc;
    if (io->dtype == TFEPB_F32) return run<float>(io, g, AffineOp<float>{}, as_stream(stream), "affine_backward");
    return run<double>(io, g, AffineOp<double>{}, as_stream(stream), "affine_backward");
}

extern "C" int tfepb_shift(const tfepb_tx_io* io, const void* period, const void* lower, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_io(io, nullptr)) return rc;
    TFEPB_CHECK_ARG(period && lower, "null period / lower table");
    if (io->dtype == TFEPB_F32)
        return run<float>(io, nullptr, ShiftOp<float>{(const float*)period, (const float*)lower}, as_stream(stream), "shift");
    return run<double>(io, nullptr, ShiftOp<double>{(const double*)period, (const double*)lower}, as_stream(stream), "shift");
}

extern "C" int tfepb_shift_backward(const tfepb_tx_io* io, const void* period, const void* lower, const tfepb_tx_grads* g,
                                    tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(g != nullptr, "null gradient struct");
    if (int rc = check_io(io, g)) return rc;
    TFEPB_CHECK_ARG(period && lower, "null period / lower table");
    if (io->dtype == TFEPB_F32)
        return run<float>(io, g, ShiftOp<float>{(const float*)period, (const float*)lower}, as_stream(stream), "shift_backward");
    return run<double>(io, g, ShiftOp<double>{(const double*)period, (const double*)lower}, as_stream(stream),
                       "shift_backward");
}

extern "C" int tfepb_sos(const tfepb_tx_io* io, int32_t n_polynomials, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_io(io, nullptr)) return rc;
    TFEPB_CHECK_ARG(n_polynomials >= 2, "n_polynomials must be strictly greater than 1.");
    TFEPB_CHECK_ARG(io->inverse == 0, "Inversion of SOS polynomial transformer has not been implemented yet.");
    if (n_polynomials == 2 && io->par_stride_p == 1) {
        if (io->dtype == TFEPB_F32) return run<float>(io, nullptr, SosPackedOp<float, 2>{}, as_stream(stream), "sos_packed");
        return run<double>(io, nullptr, SosPackedOp<double, 2>{}, as_stream(stream), "sos_packed");
    }
    if (io->dtype == TFEPB_F32) return run<float>(io, nullptr, SosOp<float>{n_polynomials}, as_stream(stream), "sos");
    return run<double>(io, nullptr, SosOp<double>{n_polynomials}, as_stream(stream), "sos");
}

extern "C" int tfepb_sos_backward(const tfepb_tx_io* io, int32_t n_polynomials, const tfepb_tx_grads* g,
                                  tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(g != nullptr, "null gradient struct");
    if (int rc = check_io(io, g)) return rc;
    TFEPB_CHECK_ARG(n_polynomials >= 2, "n_polynomials must be strictly greater than 1.");
    if (n_polynomials == 2 && io->par_stride_p == 1) {
        if (io->dtype == TFEPB_F32) return run<float>(io, g, SosPackedOp<float, 2>{}, as_stream(stream), "sos_packed_backward");
        return run<double>(io, g, SosPackedOp<double, 2>{}, as_stream(stream), "sos_packed_backward");
    }
    if (io->dtype == TFEPB_F32)
        return run<float>(io, g, SosOp<float>{n_polynomials}, as_stream(stream), "sos_backward");
    return run<double>(io, g, SosOp<double>{n_polynomials}, as_stream(stream), "sos_backward");
}

extern "C" int tfepb_moebius(const tfepb_tx_io* io, int32_t dimension, double max_radius, int32_t unit_sphere,
                             tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_io(io, nullptr)) return rc;
    TFEPB_CHECK_ARG(dimension >= 1 && dimension <= 16, "dimension must be in [1, 16]");
    TFEPB_CHECK_ARG(io->n_features % dimension == 0, "n_features must be a multiple of the vector dimension");
    TFEPB_CHECK_ARG(unit_sphere >= 0 && unit_sphere <= 2, "unit_sphere (variant) must be 0, 1 or 2");
    if (dimension == 3 && unit_sphere != 2) {
        if (io->dtype == TFEPB_F32)
            return run<float>(io, nullptr, MoebiusPackedOp<float, 3>{(float)max_radius, unit_sphere}, as_stream(stream), "moebius3");
        return run<double>(io, nullptr, MoebiusPackedOp<double, 3>{max_radius, unit_sphere}, as_stream(stream), "moebius3");
    }
    if (io->dtype == TFEPB_F32)
        return run<float>(io, nullptr, MoebiusOp<float>{dimension, (float)max_radius, unit_sphere}, as_stream(stream), "moebius");
    return run<double>(io, nullptr, MoebiusOp<double>{dimension, max_radius, unit_sphere}, as_stream(stream), "moebius");
}

extern "C" int tfepb_moebius_backward(const tfepb_tx_io* io, int32_t dimension, double max_radius, int32_t unit_sphere,
                                      const tfepb_tx_grads* g, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(g != nullptr, "null gradient struct");
    if (int rc = check_io(io, g)) return rc;
    TFEPB_CHECK_ARG(dimension >= 1 && dimension <= 16, "dimension must be in [1, 16]");
    TFEPB_CHECK_ARG(io->n_features % dimension == 0, "n_features must be a multiple of the vector dimension");
    TFEPB_CHECK_ARG(unit_sphere >= 0 && unit_sphere <= 2, "unit_sphere (variant) must be 0, 1 or 2");
    if (dimension == 3 && unit_sphere != 2) {
        if (io->dtype == TFEPB_F32)
            return run<float>(io, g, MoebiusPackedOp<float, 3>{(float)max_radius, unit_sphere}, as_stream(stream),
                              "moebius3_backward");
        return run<double>(io, g, MoebiusPackedOp<double, 3>{max_radius, unit_sphere}, as_stream(stream), "moebius3_backward");
    }
    if (io->dtype == TFEPB_F32)
        return run<float>(io, g, MoebiusOp<float>{dimension, (float)max_radius, unit_sphere}, as_stream(stream),
                          "moebius_backward");
    return run<double>(io, g, MoebiusOp<double>{dimension, max_radius, unit_sphere}, as_stream(stream), "moebius_backward");
}

extern "C" int tfepb_spline(const tfepb_tx_io* io, const tfepb_spline_cfg* cfg, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_io(io, nullptr)) return rc;
    if (int rc = check_spline(io, cfg)) return rc;
    if (io->dtype == TFEPB_F32) return spline_dispatch<float>(io, cfg, nullptr, as_stream(stream));
    return spline_dispatch<double>(io, cfg, nullptr, as_stream(stream));
}

extern "C" int tfepb_spline_backward(const tfepb_tx_io* io, const tfepb_spline_cfg* cfg, const tfepb_tx_grads* g,
                                     tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(g != nullptr, "null gradient struct");
    if (int rc = check_io(io, g)) return rc;
    if (int rc = check_spline(io, cfg)) return rc;
    if (io->dtype == TFEPB_F32) return spline_dispatch<float>(io, cfg, g, as_stream(stream));
    return spline_dispatch<double>(io, cfg, g, as_stream(stream));
}
