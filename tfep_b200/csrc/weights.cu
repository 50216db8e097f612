// Packed effective weights of a weight-normalised masked linear layer, forward and backward, one launch each.
//
// The reference recomputes  W = M o (g v / |v|_row)  in a forward pre-hook (nn/masked.py:369-371, 433-439, mask multiply
// :270) and lets autograd differentiate it; the degree-sorted conditioner then permutes rows and columns
// (tfep_b200/_pack.py).  As tensor algebra that is ~35 tiny launches per layer and training step (norm, where, divide,
// multiply, two gathers, padding, and their backward: scatters, reductions, ...), a few milliseconds per step for a
// six-layer flow.  Here: one CTA per PACKED output row.
//
//   forward   out[r][c] = M[i][j] v[i][j] g[i] / s_i,   i = row_perm[r] (-1: zero row), j = col_perm[c] (NULL: c),
//             s_i = |v[i]| over ALL columns (as the reference: the norm is taken before masking), 1 if the norm is zero;
//             bias_out[r] = bias[i].
//   backward  S_i = sum_c M v dW;  grad_g[i] = S_i / s_i;
//             grad_v[i][j] = g_i / s_i (M dW[r][c] - [|v_i| > 0] v[i][j] S_i / s_i^2);  grad_bias[i] = grad_bias_out[r].
//             (what autograd gives for effective_weight(): the `where(norm > 0, norm, 1)` branch carries no gradient.)
#include "common.cuh"

namespace tfepb {
namespace {

constexpr int WN_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                          // red may still be read from a previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WN_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(WN_THREADS) wn_pack_kernel(const float* __restrict__ v, int64_t ldv, const float* __restrict__ g,
                                                             const float* __restrict__ mask, int64_t ldm,
                                                             const float* __restrict__ bias, int cols,
                                                             const int* __restrict__ row_perm, const int* __restrict__ col_perm,
                                                             float* __restrict__ out, int64_t ldo, int out_cols,
                                                             float* __restrict__ bias_out) {
    __shared__ float red[WN_THREADS / 32];
    const int r = blockIdx.x;
    const int i = row_perm != nullptr ? row_perm[r] : r;
    float* orow = out + (int64_t)r * ldo;
    if (i < 0) {
        for (int c = threadIdx.x; c < out_cols; c += WN_THREADS) orow[c] = 0.f;
        if (bias_out != nullptr && threadIdx.x == 0) bias_out[r] = 0.f;
        return;
    }
    const float* vrow = v + (int64_t)i * ldv;
    float ss = 0.f;
    for (int k = threadIdx.x; k < cols; k += WN_THREADS) ss += vrow[k] * vrow[k];
    ss = block_sum(ss, red);
    const float n = sqrtf(ss);
    const float scale = g[i] / (n > 0.f ? n : 1.f);
    for (int c = threadIdx.x; c < out_cols; c += WN_THREADS) {
        float w = 0.f;
        if (c < cols) {
            const int j = col_perm != nullptr ? col_perm[c] : c;
            w = vrow[j] * scale;
            if (mask != nullptr) w *= mask[(int64_t)i * ldm + j];
        }
        orow[c] = w;                          // columns [cols, out_cols): zero padding of the leading dimension
    }
    if (bias_out != nullptr && threadIdx.x == 0) bias_out[r] = bias[i];
}

__global__ void __launch_bounds__(WN_THREADS) wn_pack_backward_kernel(const float* __restrict__ v, int64_t ldv, const float* __restrict__ g,
                                                                      const float* __restrict__ mask, int64_t ldm, int cols,
                                                                      const int* __restrict__ row_perm, const int* __restrict__ col_perm,
                                                                      const float* __restrict__ gout, int64_t ldgo,
                                                                      const float* __restrict__ gbias_out,
                                                                      float* __restrict__ gv, int64_t ldgv, float* __restrict__ gg,
                                                                      float* __restrict__ gbias) {
    __shared__ float red[WN_THREADS / 32];
    const int r = blockIdx.x;
    const int i = row_perm != nullptr ? row_perm[r] : r;
    if (i < 0) return;
    const float* vrow = v + (int64_t)i * ldv;
    const float* grow = gout + (int64_t)r * ldgo;
    float ss = 0.f, sd = 0.f;
    for (int c = threadIdx.x; c < cols; c += WN_THREADS) {
        const int j = col_perm != nullptr ? col_perm[c] : c;
        const float vj = vrow[j];
        const float m = mask != nullptr ? mask[(int64_t)i * ldm + j] : 1.f;
        ss += vj * vj;
        sd += m * vj * grow[c];
    }
    ss = block_sum(ss, red);
    sd = block_sum(sd, red);
    const float n = sqrtf(ss);
    const bool pos = n > 0.f;
    const float inv = pos ? 1.f / n : 1.f;
    const float gi = g[i];
    const float a = gi * inv;                              // g / s
    const float b = pos ? sd * inv * inv : 0.f;            // S / s^2 (no gradient through the guarded norm)
    for (int c = threadIdx.x; c < cols; c += WN_THREADS) {
        const int j = col_perm != nullptr ? col_perm[c] : c;
        const float m = mask != nullptr ? mask[(int64_t)i * ldm + j] : 1.f;
        gv[(int64_t)i * ldgv + j] = a * (m * grow[c] - vrow[j] * b);
    }
    if (threadIdx.x == 0) {
        gg[i] = sd * inv;
        if (gbias != nullptr) gbias[i] = gbias_out[r];
    }
}

}  // namespace
}  // namespace tfepb

using namespace tfepb;

extern "C" int tfepb_wn_pack(const float* v, int64_t ldv, const float* g, const float* mask, int64_t ldm, const float* bias,
                             int32_t rows, int32_t cols, const int32_t* row_perm, int32_t out_rows, const int32_t* col_perm,
                             float* out, int64_t ldo, int32_t out_cols, float* bias_out, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(v && g && out, "null buffer");
    TFEPB_CHECK_ARG(rows > 0 && cols > 0 && out_rows > 0 && out_cols >= cols && ldo >= out_cols && ldv >= cols, "bad sizes");
    TFEPB_CHECK_ARG(mask == nullptr || ldm >= cols, "mask leading dimension smaller than the row length");
    TFEPB_CHECK_ARG(row_perm != nullptr || out_rows == rows, "without a row permutation out_rows must equal rows");
    TFEPB_CHECK_ARG((bias == nullptr) == (bias_out == nullptr), "bias and bias_out go together");
    if (int rc = require_sm100()) return rc;
    wn_pack_kernel<<<out_rows, WN_THREADS, 0, as_stream(stream)>>>(v, ldv, g, mask, ldm, bias, cols, row_perm, col_perm, out, ldo,
                                                                  out_cols, bias_out);
    return check_launch("wn_pack");
}

extern "C" int tfepb_wn_pack_backward(const float* v, int64_t ldv, const float* g, const float* mask, int64_t ldm, int32_t rows,
                                      int32_t cols, const int32_t* row_perm, int32_t out_rows, const int32_t* col_perm,
                                      const float* grad_out, int64_t ldgo, const float* grad_bias_out, float* grad_v,
                                      int64_t ldgv, float* grad_g, float* grad_bias, tfepb_stream_t stream) {
    TFEPB_NVTX();
    TFEPB_CHECK_ARG(v && g && grad_out && grad_v && grad_g, "null buffer");
    TFEPB_CHECK_ARG(rows > 0 && cols > 0 && out_rows > 0 && ldgo >= cols && ldv >= cols && ldgv >= cols, "bad sizes");
    TFEPB_CHECK_ARG(mask == nullptr || ldm >= cols, "mask leading dimension smaller than the row length");
    TFEPB_CHECK_ARG(row_perm != nullptr || out_rows == rows, "without a row permutation out_rows must equal rows");
    TFEPB_CHECK_ARG((grad_bias == nullptr) == (grad_bias_out == nullptr), "grad_bias and grad_bias_out go together");
    if (int rc = require_sm100()) return rc;
    wn_pack_backward_kernel<<<out_rows, WN_THREADS, 0, as_stream(stream)>>>(v, ldv, g, mask, ldm, cols, row_perm, col_perm, grad_out,
                                                                           ldgo, grad_bias_out, grad_v, ldgv, grad_g, grad_bias);
    return check_launch("wn_pack_backward");
}
