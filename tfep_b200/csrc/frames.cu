// Fused pre / post kernels of the wrapper flows around the MAF kernels (inference; under autograd the same algebra
// runs as differentiable tensor operations, tfep_b200/nn/flows/{centroid,oriented}.py):
//
//   CenteredCentroidFlow  reference tfep/nn/flows/centroid.py:194-263   translate the (weighted) centroid to an origin,
//                         drop the coordinates of one fixed point, ... flow ..., place the fixed point so that the
//                         centroid is preserved, translate back
//   OrientedFlow          reference tfep/nn/flows/oriented.py:182-225 + utils/geometry.py:296-411   rotate every sample
//                         into the frame where one point lies on an axis and a second one on a plane, drop the three
//                         constrained coordinates, ... flow ..., rotate back
//
// One warp per sample, lanes stride over the features (coalesced rows); each pre kernel gathers the propagated
// features into the contiguous tensor the wrapped flow consumes and leaves the per-sample frame (translation or
// rotation matrix) in a side buffer for the post kernel, which scatters the flow's output back.  HBM bound:
// 4 D bytes read + written per sample and pass.
#include "common.cuh"

namespace tfepb {
namespace {

constexpr int FR_THREADS = 256;
constexpr int MAX_DIM = 4;

template <typename T>
struct CentroidP {
    const T* x; int64_t ldx;            // input row (pre) / original input row (post)
    const T* yprop; int64_t ldy;        // post: output of the wrapped flow (batch, n_prop)
    T* out; int64_t ldout;              // pre: (batch, n_prop); post: (batch, D)
    T* shift;                           // (batch, dim): origin - centroid
    int batch, D, dim, n_prop;
    const int* prop_cols;               // n_prop columns of x that go through the flow
    const int* full_to_prop;            // D entries: position among the propagated features, -1 for the fixed ones
    const int* points; int n_points;    // points defining the centroid (NULL = all D / dim points)
    const T* weights;                   // normalised weights of those points, or NULL (plain mean)
    T origin[MAX_DIM];
    int fixed_point;                    // point index (in the full row) of the fixed point
    int fixed_slot;                     // its position inside `points` (or its index if points == NULL)
    int restore, translate_back;
};

template <typename T>
__global__ void __launch_bounds__(FR_THREADS) centroid_pre_kernel(const CentroidP<T> p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (FR_THREADS / 32) + (threadIdx.x >> 5);
    if (b >= p.batch) return;
    const T* x = p.x + (int64_t)b * p.ldx;
    T c[MAX_DIM] = {T(0), T(0), T(0), T(0)};
    const T mean_w = T(1) / T(p.n_points);
    for (int i = lane; i < p.n_points; i += 32) {
        const int pt = p.points ? p.points[i] : i;
        const T w = p.weights ? p.weights[i] : mean_w;
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k)
            if (k < p.dim) c[k] += w * x[pt * p.dim + k];
    }
    T s[MAX_DIM];
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) s[k] = k < p.dim ? p.origin[k] - warp_sum(c[k]) : T(0);
    if (lane < p.dim) p.shift[(int64_t)b * p.dim + lane] = lane == 0 ? s[0] : (lane == 1 ? s[1] : (lane == 2 ? s[2] : s[3]));
    T* out = p.out + (int64_t)b * p.ldout;
    for (int j = lane; j < p.n_prop; j += 32) {
        const int col = p.prop_cols[j], k = col % p.dim;
        out[j] = x[col] + (k == 0 ? s[0] : (k == 1 ? s[1] : (k == 2 ? s[2] : s[3])));
    }
}

template <typename T>
__global__ void __launch_bounds__(FR_THREADS) centroid_post_kernel(const CentroidP<T> p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (FR_THREADS / 32) + (threadIdx.x >> 5);
    if (b >= p.batch) return;
    const T* y = p.yprop + (int64_t)b * p.ldy;
    T s[MAX_DIM], fixed[MAX_DIM];
#pragma unroll
    for (int k = 0; k < MAX_DIM; ++k) s[k] = k < p.dim ? p.shift[(int64_t)b * p.dim + k] : T(0);
    if (p.restore) {
        // fixed point = (origin - sum over the other defining points of w_i y_i) / w_fixed
        T c[MAX_DIM] = {T(0), T(0), T(0), T(0)};
        const T mean_w = T(1) / T(p.n_points);
        for (int i = lane; i < p.n_points; i += 32) {
            const int pt = p.points ? p.points[i] : i;
            if (pt == p.fixed_point) continue;
            const T w = p.weights ? p.weights[i] : mean_w;
#pragma unroll
            for (int k = 0; k < MAX_DIM; ++k)
                if (k < p.dim) c[k] += w * y[p.full_to_prop[pt * p.dim + k]];
        }
        const T wf = p.weights ? p.weights[p.fixed_slot] : mean_w;
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k) fixed[k] = k < p.dim ? (p.origin[k] - warp_sum(c[k])) / wf : T(0);
    } else {
        // the fixed point keeps its (centred) input coordinates
#pragma unroll
        for (int k = 0; k < MAX_DIM; ++k)
            fixed[k] = k < p.dim ? p.x[(int64_t)b * p.ldx + p.fixed_point * p.dim + k] + s[k] : T(0);
    }
    T* out = p.out + (int64_t)b * p.ldout;
    for (int col = lane; col < p.D; col += 32) {
        const int j = p.full_to_prop[col], k = col % p.dim;
        T v = j >= 0 ? y[j] : (k == 0 ? fixed[0] : (k == 1 ? fixed[1] : (k == 2 ? fixed[2] : fixed[3])));
        if (p.translate_back) v -= (k == 0 ? s[0] : (k == 1 ? s[1] : (k == 2 ? s[2] : s[3])));
        out[col] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// reference frame: R = R2 R1 (Rodrigues), utils/geometry.py:296-411 with project_on_positive_axis = False
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void rodrigues(const T k[3], T c, T s, T R[9]) {
    const T t = T(1) - c;
    R[0] = c + t * k[0] * k[0];        R[1] = t * k[0] * k[1] - s * k[2]; R[2] = t * k[0] * k[2] + s * k[1];
    R[3] = t * k[1] * k[0] + s * k[2]; R[4] = c + t * k[1] * k[1];        R[5] = t * k[1] * k[2] - s * k[0];
    R[6] = t * k[2] * k[0] - s * k[1]; R[7] = t * k[2] * k[1] + s * k[0]; R[8] = c + t * k[2] * k[2];
}

template <typename T>
__device__ void frame_rotation(const T a[3], const T q[3], int axis, int plane_axis, T R[9]) {
    T e[3] = {T(0), T(0), T(0)}, f[3] = {T(0), T(0), T(0)};
    e[axis] = T(1);
    f[plane_axis] = T(1);
    const T n[3] = {e[1] * f[2] - e[2] * f[1], e[2] * f[0] - e[0] * f[2], e[0] * f[1] - e[1] * f[0]};   // plane normal
    // first rotation: about a x axis, by the angle between a and the axis folded into [-pi/2, pi/2] (nearest half-axis)
    T v[3] = {a[1] * e[2] - a[2] * e[1], a[2] * e[0] - a[0] * e[2], a[0] * e[1] - a[1] * e[0]};
    const T an = Math<T>::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    T vn = Math<T>::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const T tiny = T(1e-8);
    if (Math<T>::abs(v[0]) <= tiny && Math<T>::abs(v[1]) <= tiny && Math<T>::abs(v[2]) <= tiny) {
        // parallel to the axis: any direction orthogonal to it (plane_axis x axis), angle 0 after folding
        v[0] = f[1] * e[2] - f[2] * e[1]; v[1] = f[2] * e[0] - f[0] * e[2]; v[2] = f[0] * e[1] - f[1] * e[0];
        vn = T(1);
    }
    const T k1[3] = {v[0] / vn, v[1] / vn, v[2] / vn};
    T cos1 = (a[0] * e[0] + a[1] * e[1] + a[2] * e[2]) / an;
    cos1 = cos1 < T(-1) ? T(-1) : (cos1 > T(1) ? T(1) : cos1);
    T sin1 = Math<T>::sqrt(T(1) - cos1 * cos1 > T(0) ? T(1) - cos1 * cos1 : T(0));
    if (cos1 < T(0)) { cos1 = -cos1; sin1 = -sin1; }
    T R1[9];
    rodrigues<T>(k1, cos1, sin1, R1);
    // second rotation: about the axis, zeroing the out-of-plane component of the second point
    T p1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p1[i] = R1[3 * i] * q[0] + R1[3 * i + 1] * q[1] + R1[3 * i + 2] * q[2];
    const T along = p1[0] * e[0] + p1[1] * e[1] + p1[2] * e[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) p1[i] -= e[i] * along;
    const T pn = Math<T>::sqrt(p1[0] * p1[0] + p1[1] * p1[1] + p1[2] * p1[2]);
    T sin2 = (p1[0] * n[0] + p1[1] * n[1] + p1[2] * n[2]) / pn;        // |n| = 1
    sin2 = sin2 < T(-1) ? T(-1) : (sin2 > T(1) ? T(1) : sin2);
    const T side = p1[0] * f[0] + p1[1] * f[1] + p1[2] * f[2];
    const T sgn = side > T(0) ? T(-1) : (side < T(0) ? T(1) : T(0));
    T cos2 = Math<T>::sqrt(T(1) - sin2 * sin2 > T(0) ? T(1) - sin2 * sin2 : T(0));
    if (sgn == T(0)) { cos2 = T(1); }
    T R2[9];
    rodrigues<T>(e, cos2, sgn * sin2, R2);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[3 * i + j] = R2[3 * i] * R1[j] + R2[3 * i + 1] * R1[3 + j] + R2[3 * i + 2] * R1[6 + j];
}

template <typename T>
struct OrientedP {
    const T* x; int64_t ldx;
    const T* yprop; int64_t ldy;
    T* out; int64_t ldout;
    T* rot;                              // (batch, 9)
    int batch, D, n_prop;
    const int* prop_cols;
    const int* full_to_prop;
    int axis_point, plane_point, axis, plane_axis;
    int round_off, rotate_back;
};

template <typename T>
__global__ void __launch_bounds__(FR_THREADS) oriented_pre_kernel(const OrientedP<T> p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (FR_THREADS / 32) + (threadIdx.x >> 5);
    if (b >= p.batch) return;
    const T* x = p.x + (int64_t)b * p.ldx;
    T a[3], q[3], R[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = x[3 * p.axis_point + k]; q[k] = x[3 * p.plane_point + k]; }
    frame_rotation<T>(a, q, p.axis, p.plane_axis, R);
    if (lane < 9) {
        T r = R[0];
#pragma unroll
        for (int i = 1; i < 9; ++i) r = lane == i ? R[i] : r;
        p.rot[(int64_t)b * 9 + lane] = r;
    }
    T* out = p.out + (int64_t)b * p.ldout;
    for (int j = lane; j < p.n_prop; j += 32) {
        const int col = p.prop_cols[j], pt = col / 3, k = col - 3 * pt;
        const T r0 = k == 0 ? R[0] : (k == 1 ? R[3] : R[6]), r1 = k == 0 ? R[1] : (k == 1 ? R[4] : R[7]),
                r2 = k == 0 ? R[2] : (k == 1 ? R[5] : R[8]);
        out[j] = r0 * x[3 * pt] + r1 * x[3 * pt + 1] + r2 * x[3 * pt + 2];
    }
}

template <typename T>
__global__ void __launch_bounds__(FR_THREADS) oriented_post_kernel(const OrientedP<T> p) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (FR_THREADS / 32) + (threadIdx.x >> 5);
    if (b >= p.batch) return;
    const T* x = p.x + (int64_t)b * p.ldx;
    const T* y = p.yprop + (int64_t)b * p.ldy;
    T R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = p.rot[(int64_t)b * 9 + i];
    T* out = p.out + (int64_t)b * p.ldout;
    for (int pt = lane; pt < p.D / 3; pt += 32) {
        T v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = p.full_to_prop[3 * pt + k];
            if (j >= 0) v[k] = y[j];
            else v[k] = p.round_off ? T(0) : R[3 * k] * x[3 * pt] + R[3 * k + 1] * x[3 * pt + 1] + R[3 * k + 2] * x[3 * pt + 2];
        }
        if (p.rotate_back) {
            // y R: out_k = sum_m y_m R[m][k]
#pragma unroll
            for (int k = 0; k < 3; ++k) out[3 * pt + k] = v[0] * R[k] + v[1] * R[3 + k] + v[2] * R[6 + k];
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) out[3 * pt + k] = v[k];
        }
    }
}

template <typename T>
int launch_centroid(const tfepb_centroid_args* a, bool post, cudaStream_t s) {
    CentroidP<T> p{};
    p.x = (const T*)a->x; p.ldx = a->ldx; p.yprop = (const T*)a->y_propagated; p.ldy = a->ldy;
    p.out = (T*)a->out; p.ldout = a->ldout; p.shift = (T*)a->shift;
    p.batch = a->batch; p.D = a->n_features; p.dim = a->space_dimension; p.n_prop = a->n_propagated;
    p.prop_cols = a->propagated_columns; p.full_to_prop = a->column_to_propagated;
    p.points = a->centroid_points; p.n_points = a->n_centroid_points; p.weights = (const T*)a->weights;
    for (int k = 0; k < MAX_DIM; ++k) p.origin[k] = k < a->space_dimension ? (T)a->origin[k] : T(0);
    p.fixed_point = a->fixed_point; p.fixed_slot = a->fixed_slot;
    p.restore = a->restore_fixed_point; p.translate_back = a->translate_back;
    const int grid = (a->batch + FR_THREADS / 32 - 1) / (FR_THREADS / 32);
    if (post) centroid_post_kernel<T><<<grid, FR_THREADS, 0, s>>>(p);
    else centroid_pre_kernel<T><<<grid, FR_THREADS, 0, s>>>(p);
    return check_launch(post ? "centroid_post" : "centroid_pre");
}

template <typename T>
int launch_oriented(const tfepb_oriented_args* a, bool post, cudaStream_t s) {
    OrientedP<T> p{};
    p.x = (const T*)a->x; p.ldx = a->ldx; p.yprop = (const T*)a->y_propagated; p.ldy = a->ldy;
    p.out = (T*)a->out; p.ldout = a->ldout; p.rot = (T*)a->rotation;
    p.batch = a->batch; p.D = a->n_features; p.n_prop = a->n_propagated;
    p.prop_cols = a->propagated_columns; p.full_to_prop = a->column_to_propagated;
    p.axis_point = a->axis_point; p.plane_point = a->plane_point; p.axis = a->axis; p.plane_axis = a->plane_axis;
    p.round_off = a->round_off_imprecisions; p.rotate_back = a->rotate_back;
    const int grid = (a->batch + FR_THREADS / 32 - 1) / (FR_THREADS / 32);
    if (post) oriented_post_kernel<T><<<grid, FR_THREADS, 0, s>>>(p);
    else oriented_pre_kernel<T><<<grid, FR_THREADS, 0, s>>>(p);
    return check_launch(post ? "oriented_post" : "oriented_pre");
}

}  // namespace
}  // namespace tfepb

using namespace tfepb;

static int check_centroid(const tfepb_centroid_args* a, bool post) {
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->dtype == TFEPB_F32 || a->dtype == TFEPB_F64, "dtype must be TFEPB_F32 or TFEPB_F64");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->n_features > 0, "bad sizes");
    TFEPB_CHECK_ARG(a->space_dimension >= 1 && a->space_dimension <= 4 && a->n_features % a->space_dimension == 0,
                    "space_dimension must be in [1, 4] and divide n_features");
    TFEPB_CHECK_ARG(a->n_centroid_points > 0 && a->n_propagated == a->n_features - a->space_dimension, "bad point counts");
    TFEPB_CHECK_ARG(a->x && a->out && a->shift && a->propagated_columns && a->column_to_propagated, "null buffer");
    TFEPB_CHECK_ARG(!post || a->y_propagated != nullptr, "null buffer");
    return 0;
}

extern "C" int tfepb_centroid_pre(const tfepb_centroid_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_centroid(a, false)) return rc;
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    return a->dtype == TFEPB_F32 ? launch_centroid<float>(a, false, as_stream(stream)) : launch_centroid<double>(a, false, as_stream(stream));
}

extern "C" int tfepb_centroid_post(const tfepb_centroid_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_centroid(a, true)) return rc;
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    return a->dtype == TFEPB_F32 ? launch_centroid<float>(a, true, as_stream(stream)) : launch_centroid<double>(a, true, as_stream(stream));
}

static int check_oriented(const tfepb_oriented_args* a, bool post) {
    TFEPB_CHECK_ARG(a != nullptr, "null argument struct");
    TFEPB_CHECK_ARG(a->dtype == TFEPB_F32 || a->dtype == TFEPB_F64, "dtype must be TFEPB_F32 or TFEPB_F64");
    TFEPB_CHECK_ARG(a->batch >= 0 && a->n_features > 0 && a->n_features % 3 == 0, "n_features must be a multiple of 3");
    TFEPB_CHECK_ARG(a->axis >= 0 && a->axis < 3 && a->plane_axis >= 0 && a->plane_axis < 3 && a->axis != a->plane_axis, "bad axes");
    TFEPB_CHECK_ARG(a->axis_point >= 0 && a->plane_point >= 0 && a->axis_point != a->plane_point &&
                    3 * a->axis_point < a->n_features && 3 * a->plane_point < a->n_features, "bad point indices");
    TFEPB_CHECK_ARG(a->n_propagated == a->n_features - 3, "three coordinates are constrained");
    TFEPB_CHECK_ARG(a->x && a->out && a->rotation && a->propagated_columns && a->column_to_propagated, "null buffer");
    TFEPB_CHECK_ARG(!post || a->y_propagated != nullptr, "null buffer");
    return 0;
}

extern "C" int tfepb_oriented_pre(const tfepb_oriented_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_oriented(a, false)) return rc;
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    return a->dtype == TFEPB_F32 ? launch_oriented<float>(a, false, as_stream(stream)) : launch_oriented<double>(a, false, as_stream(stream));
}

extern "C" int tfepb_oriented_post(const tfepb_oriented_args* a, tfepb_stream_t stream) {
    TFEPB_NVTX();
    if (int rc = check_oriented(a, true)) return rc;
    if (int rc = require_sm100()) return rc;
    if (a->batch == 0) return 0;
    return a->dtype == TFEPB_F32 ? launch_oriented<float>(a, true, as_stream(stream)) : launch_oriented<double>(a, true, as_stream(stream));
}
