/*
 * tfep_b200 -- C ABI of the B200 (sm_100a) implementation of the tfep MAF / (T)FEP hot path.
 *
 * The reference (andrrizzi/tfep) is pure Python on PyTorch and has no FFI layer: its boundary for
 * this path is the Python class / function API.  This header is the contract
 * a reference-side binding would load with ctypes; every entry point names the
 * reference code it replaces (paths relative to the reference's tfep/ package).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, explicit sizes and leading dimensions (in elements), no torch types;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing here allocates device memory;
 *   - calls enqueue work on `stream` (a cudaStream_t passed as void*) and return without synchronising;
 *   - return value: 0 = ok, <0 = invalid argument, >0 = cudaError_t; tfepb_last_error() gives the
 *     thread-local message of the last failure;
 *   - `dtype` selects the arithmetic type of every floating-point buffer of that call
 *     (TFEPB_F32 or TFEPB_F64); the tensor-core entry points are fp32-in/fp32-out with bf16 operands;
 *   - a transformer parameter (sample b, parameter p, feature f) lives at
 *         par[b * ldp + par_offset + p * par_stride_p + f * par_stride_f]
 *     or, when the optional table par_base is given, at
 *         par[b * ldp + par_offset + par_base[f] + p * par_stride_p].
 *     This covers the reference's parameter-major layout (stride_p = n_features, stride_f = 1;
 *     nn/transformers/spline.py:352, affine.py:140, sos.py:108), Moebius' native layout
 *     (moebius.py:102) and the degree-sorted feature-major packing of the fused paths;
 *   - `cols` (optional, int32, length n_features) maps transformer feature f to its column in x / y
 *     (conditioning features, nn/flows/autoregressive.py:164-175; MixedTransformer groups, mixed.py:168-186).
 */
#ifndef TFEP_B200_H
#define TFEP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFEPB_ABI_VERSION 5

enum { TFEPB_F32 = 0, TFEPB_F64 = 1 };
enum { TFEPB_ACT_NONE = 0, TFEPB_ACT_ELU = 1 };

typedef void* tfepb_stream_t;

/* ----------------------------------------------------------------------------------------------
 * status / device
 * -------------------------------------------------------------------------------------------- */
int tfepb_abi_version(void);
const char* tfepb_last_error(void);
/* Fills SM count and compute capability of the current device.  The library refuses to run
 * compute entry points on anything but compute capability 10.x (no fallback path). */
int tfepb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ----------------------------------------------------------------------------------------------
 * MADE conditioner, exact-arithmetic SIMT path (any shape, fp32 / fp64)
 * Replaces MaskedLinearFunc.forward / .backward (nn/masked.py:266-302) and the ELU of MADE
 * (nn/conditioners/made.py:320).  `w` is the EFFECTIVE weight M o (g v/|v|) (nn/masked.py:369-371),
 * so the mask is already folded in; `k_ranges` lets the kernel skip the all-zero part of the
 * reduction for a tile of TFEPB_GEMM_TILE_N consecutive output columns.
 * -------------------------------------------------------------------------------------------- */
#define TFEPB_GEMM_TILE_N 64

typedef struct {
    int32_t dtype;
    int32_t batch, in_features, out_features;
    const void* x;    int64_t ldx;     /* (batch, in)  */
    const void* w;    int64_t ldw;     /* (out, in) effective weights */
    const void* bias;                  /* (out,) or NULL */
    void* y;          int64_t ldy;     /* (batch, out) */
    int32_t activation;                /* TFEPB_ACT_* applied to y */
    int32_t reserved;
    const int32_t* k_ranges;           /* NULL or (ceil(out/TILE_N), 2) int32 on the device: [begin,end) */
} tfepb_linear_fwd_args;
/* y = act(x w^T + bias) */
int tfepb_masked_linear_forward(const tfepb_linear_fwd_args* a, tfepb_stream_t stream);

typedef struct {
    int32_t dtype;
    int32_t batch, in_features, out_features;
    const void* grad_y;  int64_t ldgy;   /* (batch, out): cotangent of the PRE-activation output */
    const void* w;       int64_t ldw;    /* (out, in) effective weights */
    void* grad_x;        int64_t ldgx;   /* (batch, in) */
    const void* act_out; int64_t ldact;  /* NULL, or (batch, in) post-ELU activations of the PREVIOUS layer:
                                            grad_x is multiplied by ELU'(.) = (h > 0 ? 1 : h + 1) */
    int32_t accumulate;                  /* 0: grad_x = ..., 1: grad_x += ... */
    int32_t reserved;
    const int32_t* n_ranges;             /* NULL or (ceil(in/TILE_N), 2): nonzero range of the out-reduction */
} tfepb_linear_bwd_input_args;
/* grad_x = (grad_y w) [* ELU'(act_out)] */
int tfepb_masked_linear_backward_input(const tfepb_linear_bwd_input_args* a, tfepb_stream_t stream);

typedef struct {
    int32_t dtype;
    int32_t batch, in_features, out_features;
    const void* grad_y;  int64_t ldgy;   /* (batch, out) */
    const void* x;       int64_t ldx;    /* (batch, in) */
    void* grad_w;        int64_t ldgw;   /* (out, in); must be zero-filled by the caller (split-batch atomics) */
    void* grad_bias;                     /* (out,) zero-filled, or NULL */
    const int32_t* n_ranges;             /* NULL or (ceil(in/TILE_N), 2): per tile of TILE_N input columns the rows
                                            [begin, end) of grad_w the mask can leave non-zero; tiles outside are
                                            skipped and stay zero (the bias gradient is always complete) */
} tfepb_linear_bwd_weight_args;
/* grad_w += grad_y^T x ; grad_bias += sum_b grad_y   (the caller applies the mask, nn/masked.py:296-297) */
int tfepb_masked_linear_backward_weight(const tfepb_linear_bwd_weight_args* a, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * transformers: elementwise map + per-sample log|det J| reduction
 * -------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t dtype;
    int32_t batch, n_features;
    int32_t inverse;                    /* 0 forward, 1 inverse */
    const void* x;   int64_t ldx;       /* input  (batch, *) */
    void* y;         int64_t ldy;       /* output (batch, *) */
    const void* par; int64_t ldp;       /* parameters (batch, *) */
    int64_t par_offset, par_stride_p, par_stride_f;
    const int32_t* par_base;            /* NULL, or (n_features,) int32: offset of parameter 0 of feature f in a
                                           row of `par`, replacing f * par_stride_f (degree-sorted
                                           feature-major packing, MixedTransformer groups) */
    const int32_t* cols;                /* NULL or int32 table: column of feature f in x / y */
    const int32_t* feat_ids;            /* NULL (features 0..n_features-1), or (n_features,) int32 ids of the
                                           features to process; cols, par_base and the spline domain tables
                                           are indexed by the id (one degree group of the inverse sweep) */
    void* logdet;                       /* (batch,) */
    int32_t accumulate_logdet;          /* 0: logdet = ld, 1: logdet += ld */
    int32_t reserved;
} tfepb_tx_io;

/* AffineTransformer: y = x exp(a) + b, ld = sum a; inverse x = (y - b) exp(-a), ld = -sum a.
 * Parameter 0 = shift, 1 = log-scale.  nn/transformers/affine.py:281-363. */
int tfepb_affine(const tfepb_tx_io* io, tfepb_stream_t stream);

typedef struct {
    int32_t n_bins;
    int32_t circular, identity_boundary_slopes, learn_lower_bound, learn_upper_bound;
    int32_t reserved;
    const void* x0; const void* xf; const void* y0; const void* yf;   /* (n_features,) each, dtype of io */
    double min_bin_size, min_slope;
    int32_t* bins_out;                  /* NULL, or (batch, n_features) int32: 0 left tail, 1..K bins, K+1 right tail */
    int64_t ldbins;
} tfepb_spline_cfg;
/* NeuralSplineTransformer forward / inverse incl. softmax/softplus parameter normalisation, the
 * circular shift + wrap, identity boundary slopes and learnable bounds.
 * nn/transformers/spline.py:184-261 (module), :319-417 (_get_parameters), :424-650 (functional). */
int tfepb_spline(const tfepb_tx_io* io, const tfepb_spline_cfg* cfg, tfepb_stream_t stream);

/* SOSPolynomialTransformer forward (no inverse in the reference, sos.py:111-114).
 * nn/transformers/sos.py:207-235, 271-306. */
int tfepb_sos(const tfepb_tx_io* io, int32_t n_polynomials, tfepb_stream_t stream);

/* MoebiusTransformer forward; the inverse is the same map with -par (moebius.py:142-147),
 * selected with io->inverse.  n_features must be a multiple of `dimension` (<= 16).
 * nn/transformers/moebius.py:374-478.
 * `unit_sphere` selects the variant: 0 sphere of radius |x|, 1 unit sphere, 2 SymmetrizedMoebiusTransformer
 * (nn/transformers/moebius.py:481-629: y = |x| s / |s|, s = f(x; w) + f(x; -w), closed-form log-det, analytic
 * inverse with io->inverse). */
int tfepb_moebius(const tfepb_tx_io* io, int32_t dimension, double max_radius, int32_t unit_sphere,
                  tfepb_stream_t stream);

/* Vector-Jacobian products of the forward maps above (what PyTorch autograd computes for the
 * reference; SOS follows the reference's hand-written backward, sos.py:237-268, which drops the
 * log-det cotangent).  grad_par uses the same (ldp, offset, strides) addressing as par. */
typedef struct {
    const void* grad_y;      int64_t ldgy;    /* (batch, *) cotangent of y, addressed through cols */
    const void* grad_logdet;                  /* (batch,) cotangent of logdet, or NULL (= 0) */
    void* grad_x;            int64_t ldgx;    /* (batch, *) written through cols */
    void* grad_par;                           /* same addressing as io->par */
} tfepb_tx_grads;
int tfepb_affine_backward(const tfepb_tx_io* io, const tfepb_tx_grads* g, tfepb_stream_t stream);
/* Volume-preserving shift y = x + b (VolumePreservingShiftTransformer, nn/transformers/affine.py:148-275, 366-456);
 * `period[f]` > 0 wraps feature f as (x + b) % period + lower[f] (tables of n_features values, dtype of the call);
 * one parameter per feature, log-det = 0. */
int tfepb_shift(const tfepb_tx_io* io, const void* period, const void* lower, tfepb_stream_t stream);
int tfepb_shift_backward(const tfepb_tx_io* io, const void* period, const void* lower, const tfepb_tx_grads* g,
                         tfepb_stream_t stream);

int tfepb_spline_backward(const tfepb_tx_io* io, const tfepb_spline_cfg* cfg, const tfepb_tx_grads* g,
                          tfepb_stream_t stream);
int tfepb_sos_backward(const tfepb_tx_io* io, int32_t n_polynomials, const tfepb_tx_grads* g,
                       tfepb_stream_t stream);
int tfepb_moebius_backward(const tfepb_tx_io* io, int32_t dimension, double max_radius, int32_t unit_sphere,
                           const tfepb_tx_grads* g, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Packed effective weight of a weight-normalised masked linear layer, one launch (fp32).  Replaces, for the training path,
 * MaskedWeightNorm.compute_weight + _ApplyMask + the mask multiply of MaskedLinearFunc.forward (nn/masked.py:369-371,
 * 433-439, 270) followed by the row / column permutation of the degree-sorted conditioner, and their autograd backward:
 *   out[r][c] = mask[i][j] v[i][j] g[i] / s_i,  i = row_perm[r] (-1: a zero row), j = col_perm[c] (NULL: c),
 *   s_i = |v[i]| over all `cols` columns (1 if the norm is zero); columns [cols, out_cols) of `out` are zero padding;
 *   bias_out[r] = bias[i] (both NULL to skip).
 * Backward: grad_v (rows x cols, ldgv), grad_g (rows), grad_bias (rows) from grad_out (out_rows x cols, ldgo) and
 * grad_bias_out; rows of v that no packed row refers to are left untouched (zero-fill them if row_perm is not onto).
 * -------------------------------------------------------------------------------------------- */
int tfepb_wn_pack(const float* v, int64_t ldv, const float* g, const float* mask, int64_t ldm, const float* bias,
                  int32_t rows, int32_t cols, const int32_t* row_perm, int32_t out_rows, const int32_t* col_perm,
                  float* out, int64_t ldo, int32_t out_cols, float* bias_out, tfepb_stream_t stream);
int tfepb_wn_pack_backward(const float* v, int64_t ldv, const float* g, const float* mask, int64_t ldm, int32_t rows,
                           int32_t cols, const int32_t* row_perm, int32_t out_rows, const int32_t* col_perm,
                           const float* grad_out, int64_t ldgo, const float* grad_bias_out, float* grad_v, int64_t ldgv,
                           float* grad_g, float* grad_bias, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * General masked linear layers on the tensor cores (tcgen05, bf16 operands, fp32 accumulation): any MADE shape,
 * forward and backward.  C[m x n] = A[m x k] . B[n x k]^T with operand IMAGES made by tfepb_tc_pack: blocks of
 * (block_rows x 64 k) bf16, block (rb, kb) at (rb * ceil(k / 64) + kb) * block_rows * 128 bytes, inside a block the
 * shared-memory operand layout (for each slab of 8 k-values: block_rows x 16 bytes), so a block is one bulk copy.
 * A operands use block_rows = 128, B operands 256.  The three products of nn/masked.py:266-302:
 *   forward          A = image(x),           B = image(w_eff)              (+ bias, ELU, image of the result = next A)
 *   backward input   A = image(grad_y),      B = image(w_eff^T, transpose) (+ ELU' multiplier `aux`)
 *   backward weight  A = image(grad_y^T),    B = image(x^T), both transposed, split_k > 1 (fp32 atomics into a
 *                    zero-filled C; the reduction runs over the batch)
 * -------------------------------------------------------------------------------------------- */
int64_t tfepb_tc_image_bytes(int64_t rows, int64_t k, int32_t block_rows);
/* src fp32: element (row, k) at src[row * ld + k], or src[k * ld + row] if transpose */
int tfepb_tc_pack(const float* src, int64_t ld, int32_t rows, int32_t k, int32_t block_rows, int32_t transpose,
                  void* image, tfepb_stream_t stream);
/* Split-precision images: n_split (1, 2 or 3) consecutive images of tfepb_tc_image_bytes() bytes each, term t holding
 * bf16(src - sum of the terms before it), i.e. 8 / 16 / 24 significant bits of every element (nn/masked.py:266-277 at
 * fp32-class accuracy on bf16 tensor cores; see n_split of tfepb_tc_gemm_args). */
int tfepb_tc_pack_split(const float* src, int64_t ld, int32_t rows, int32_t k, int32_t block_rows, int32_t transpose,
                        int32_t n_split, void* image, tfepb_stream_t stream);
/* One pass over src (rows x cols fp32): its image with block_rows = 128 / 256 (or NULL), the image of its transpose with
 * t_block_rows = 128 / 256 (or NULL) and, if column_sums != NULL (zero-filled by the caller), += the sums over the rows.
 * (x and grad_y of a layer: 128 / 256 or 128 / 128; a weight matrix: 256 / 256 = the B operands of the forward and of the
 * backward-input product.) */
int tfepb_tc_pack_dual(const float* src, int64_t ld, int32_t rows, int32_t cols, void* image, int32_t block_rows,
                       void* image_t, int32_t t_block_rows, float* column_sums, tfepb_stream_t stream);
/* Transformer fused into the epilogue of the OUTPUT-layer product of a MAF (AutoregressiveFlow.forward,
 * nn/flows/autoregressive.py:144-177: parameters = conditioner(x); y, log_det = transformer(x, parameters)) so that the
 * (batch x n_parameters) matrix never reaches memory.  The n output columns of the product are 16-column chunks; chunk q
 * holds the parameters of units q * U .. q * U + U - 1, P consecutive columns per unit, pad columns after them:
 *   TFEPB_TCTX_AFFINE   U = 8 features, P = 2 (shift, log_scale)                 affine.py:83-117
 *   TFEPB_TCTX_SOS2     U = 3 features, P = 5 (a0, a10, a11, a20, a21), 1 pad    sos.py:163-268
 *   TFEPB_TCTX_MOEBIUS3 U = 5 three-vectors, P = 3 (the vector w), 1 pad         moebius.py:120-260 (variants 0 / 1 of
 *                                                                                tfepb_moebius' unit_sphere argument)
 * (the caller permutes / pads the rows of the packed weight and bias; pad rows are zero).
 * backward == 0: y[:, cols] = T(x[:, cols]; parameters), logdet[b] += log|det J| (atomic: zero-fill or carry the sum of
 * earlier layers); the product needs no c / out_image.
 * backward == 1: the product is recomputed, grad_x[:, cols] = direct term of the VJP, and the values handed to
 * out_image / out_image_t / column_sums are the parameter cotangents (zero in pad columns): the operands of the
 * backward-input and weight-gradient products of the output layer and its bias gradient (nn/masked.py:279-302). */
#define TFEPB_TCTX_AFFINE 1
#define TFEPB_TCTX_SOS2 2
#define TFEPB_TCTX_MOEBIUS3 3
#define TFEPB_TCTX_SPLINE8 4             /* neural spline with 8 bins (spline.py:184-241; circular or not, identity boundary slopes,
                                            learnable bounds): one feature per 32 columns = its 23..27 parameters + padding */
typedef struct {
    int32_t kind;                        /* TFEPB_TCTX_* */
    int32_t backward;
    int32_t n_units;                     /* features (3-vectors for Moebius) */
    int32_t unit_sphere;                 /* Moebius: 0 sphere of radius |x|, 1 unit sphere */
    double max_radius;                   /* Moebius */
    const int32_t* cols;                 /* device (n_units * x columns per unit): columns of x / y of every unit */
    const void* x; int64_t ldx;          /* fp32 */
    void* y; int64_t ldy;                /* forward */
    float* logdet;                       /* forward, or NULL */
    const void* grad_y; int64_t ldgy;    /* backward */
    const float* grad_logdet;            /* backward, or NULL (ignored by SOS: sos.py:233) */
    void* grad_x; int64_t ldgx;          /* backward */
    /* TFEPB_TCTX_SPLINE8 only: device, 16-byte aligned, 8 floats per unit (unit order) = x0, xf, y0, yf (domain [x0, xf] ->
     * [y0, yf]), min_bin_size, min_slope, flags as int32 bits (0 circular, 1 identity_boundary_slopes, 2 learn_lower_bound,
     * 3 learn_upper_bound), unused: the units of a MixedTransformer of splines may differ in every option */
    const float* spline_table;
} tfepb_tc_tx;

typedef struct {
    const void* a_image; const void* b_image;
    int32_t m, n, k;
    int32_t activation;                  /* TFEPB_ACT_* applied after the bias */
    void* c; int64_t ldc;                /* fp32 (m, n) row-major, or NULL */
    const void* bias;                    /* fp32 (n,) or NULL */
    const void* aux; int64_t ldaux;      /* NULL, or fp32 (m, n): the result is multiplied by ELU'(aux) */
    void* out_image;                     /* NULL, or the bf16 image (block_rows = 128, k = n) of the result */
    const int32_t* k_block_ranges;       /* device, NULL or (ceil(n / 256), 2): [first, end) 64-wide k-blocks that are
                                            non-zero for each tile of 256 output columns (staircase masks) */
    int32_t split_k;                     /* > 1: the reduction is split over that many CTAs per tile */
    int32_t out_image_t_rows;            /* block_rows of out_image_t: 128 (it will be an A operand) or 256 (a B operand) */
    int32_t* error_flag;
    const int32_t* row_ranges;           /* device, NULL or (ceil(n / 256), 2): rows [begin, end) of C that can be
                                            non-zero for each tile of 256 columns; tiles outside are skipped and
                                            left untouched (masked weight gradient into a zero-filled C) */
    void* out_image_t;                   /* NULL, or the bf16 image of the TRANSPOSED result (rows = the n columns, k = the m
                                            rows; entries with k >= m are zero): the operand the weight-gradient product
                                            needs, written by the epilogue instead of a separate pack launch */
    float* column_sums;                  /* NULL, or fp32 (n,), zero-filled by the caller: += sum over the m rows of the
                                            result (the bias gradient of nn/masked.py:298-300 when the result is grad_y) */
    int32_t n_split;                     /* 0 / 1: plain bf16 operands.  2 / 3: a_image, b_image (and out_image) are
                                            split-precision images (tfepb_tc_pack_split); every k-step accumulates the
                                            3 / 6 products A_i B_j with i + j < n_split in fp32 -- operands carried to
                                            16 / 24 significant bits (forward products only) */
    int32_t c_accumulate;                /* != 0: c += result instead of c = result (without split_k, which always adds) */
    const tfepb_tc_tx* tx;               /* NULL, or the transformer fused into the epilogue (n_split <= 1, no split_k,
                                            n a multiple of 16, activation NONE, no aux) */
    int32_t mn_major;                    /* != 0: a_image / b_image are the ROW images (block_rows = 128) of a (k x m) and a (k x n)
                                            matrix and the product reduces over their rows: C = A^T B read MN-major, e.g. the
                                            weight gradient dW = dY^T X straight from the images of dY and X that the other
                                            products use -- no transposed images.  Rows >= k of the images must be zero
                                            (tfepb_tc_pack / out_image write them so).  The product is ADDED to c (fp32 atomics,
                                            zero-fill c first); split_k cuts the reduction in blocks of 128 rows; row_ranges
                                            applies; no bias / activation / images. */
    int32_t cluster;                     /* != 0: launch as clusters of two CTAs that work on two m-tiles of the same n-tile and share
                                            the B operand (each fetches half of every B block, multicast to both): -33 % L2 -> SM
                                            traffic; plain bf16 products without split_k / row_ranges, else ignored */
    const int32_t* tile_list;            /* mn_major only: NULL, or device (n_tile_list, 2) = the (128-row tile, 256-column tile) pairs of
                                            c that the mask leaves non-zero; only those are computed, evenly spread over the CTAs
                                            (instead of row_ranges, whose skipped tiles leave CTAs idle) */
    int32_t n_tile_list; int32_t reserved3;
    const void* aux_image;               /* alternative to aux: the same (m, n) operand h given as its bf16 image (block_rows =
                                            128, k = n; e.g. the out_image a forward product wrote): the result is multiplied
                                            by ELU'(h) of the bf16 values -- no fp32 copy of the activations is needed */
} tfepb_tc_gemm_args;
int tfepb_tc_gemm(const tfepb_tc_gemm_args* a, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * PeriodicEmbedding of the MADE input (nn/embeddings/mafembed.py:112-142): input column c is copied to output
 * column out_col[c], or, if periodic[c], lifted to out[out_col[c]], out[out_col[c] + 1] =
 * cos, sin((x - lower) * scale) with scale = 2 pi / (upper - lower).  Backward: grad_x from grad_out.
 * -------------------------------------------------------------------------------------------- */
int tfepb_periodic_embedding(int32_t dtype, const void* x, int64_t ldx, int32_t batch, int32_t n_in,
                             const int32_t* out_col, const int32_t* periodic, double lower, double scale,
                             void* out, int64_t ldo, tfepb_stream_t stream);
int tfepb_periodic_embedding_backward(int32_t dtype, const void* x, int64_t ldx, int32_t batch, int32_t n_in,
                                      const int32_t* out_col, const int32_t* periodic, double lower, double scale,
                                      const void* grad_out, int64_t ldgo, void* grad_x, int64_t ldgx,
                                      tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Fused MAF forward on the tensor cores (tcgen05 / TMEM / bulk-copy engine), bf16 operands with fp32
 * accumulation: y, logdet = T(x ; MADE(x)) for MADEs with two hidden layers and a circular neural-spline
 * transformer with 8 bins, for a CHAIN of n_layers >= 1 MAF layers in ONE launch.  Replaces
 * SequentialFlow._pass (nn/flows/sequential.py:50-68) over AutoregressiveFlow.forward
 * (nn/flows/autoregressive.py:144-177): per layer the three masked linear layers (nn/masked.py:266-277),
 * the two ELUs (nn/conditioners/made.py:320) and NeuralSplineTransformer.forward
 * (nn/transformers/spline.py:184-241); logdet is the sum over the layers.  The caller packs the
 * degree-sorted effective weights into shared-memory-image blocks and lists the non-zero blocks in `ops`
 * (tfep_b200/_fused.py documents the format); masked blocks are simply absent from the schedule.  Biases
 * travel inside the weight blocks (two constant-one columns in every A operand); hidden-layer rows and
 * softmax / softplus rows are pre-scaled by log2(e).
 * The persistent grid walks (layer, tile-of-128-samples) work items layer-major; a tile of layer l+1 is
 * released by a per-tile flag in `tile_flags` once layer l published it to `y` (which doubles as the
 * inter-layer buffer; x == y is allowed).
 * -------------------------------------------------------------------------------------------- */
#define TFEPB_FUSED_MAX_LAYERS 8
#define TFEPB_FUSED_MAX_OPS 512          /* over all layers of one launch */

typedef struct {
    uint32_t w_off;                      /* block position in the layer's `weights` (bytes, multiple of 16);
                                            the block holds n x (16 ksteps) bf16 = n * ksteps * 32 bytes */
    uint32_t idesc;                      /* tcgen05 instruction descriptor: kind::f16, bf16 x bf16 -> fp32, M = 128, N = n */
    uint16_t n, tmem_col, a_col;         /* MMA N, accumulator column, first A column (2 k-values per column) */
    uint8_t ksteps;                      /* K = 16 steps in the block */
    uint8_t flags;                       /* 1 first block of accumulator, 2 commit, bits 2-3 accumulator buffer
                                            (0..2), 16 wait for A operand, 32 wait for drained accumulator,
                                            64 issued by the second MMA warp, 128 hidden-layer block */
} tfepb_fused_op;

typedef struct {
    int32_t col;                         /* column of the feature in x / y; -1 = padding slot */
    float x0, period;                    /* lower limit and width xf - x0 of the spline domain */
    float inv_period;                    /* 1 / period */
    float rescaled_width, rescaled_height, y0;   /* period - 8 min_bin, (yf - y0) - 8 min_bin, lower limit of y */
    int32_t kind;                        /* 0 circular (8 slopes + shift), 1 not circular (9 slopes, linear tails) */
} tfepb_fused_feature;

typedef struct {
    const tfepb_fused_op* ops;           /* HOST array of n_ops entries (copied into the launch) */
    int32_t n_ops, n_chunks;             /* n_chunks: output-layer chunks of 4 feature slots x 28 columns */
    const void* weights;                 /* device, packed bf16 blocks */
    const tfepb_fused_feature* feats;    /* device: n_chunks * 4 */
    float min_bin_size, min_slope, slope_offset;   /* slope_offset = log(exp(1 - min_slope) - 1) */
    int32_t reserved;
    const int32_t* input_map;            /* device, k1 entries, or NULL (conditioner input k = x column k, then the two
                                            constant ones).  Entry k: bits 0-15 the x column, bits 16-18 what enters the
                                            conditioner: 0 x, 1 cos(a), 2 sin(a) with a = (x - emb_lower) * emb_scale
                                            (PeriodicEmbedding fused into the operand staging, reference
                                            nn/embeddings/mafembed.py:112-142), 3 the constant one, 4 zero */
    float emb_lower, emb_scale;          /* emb_scale = 2 pi / (upper - lower) */
    int32_t hidden_split[2];             /* this layer's split of the two hidden layers (see tfepb_fused_args); 0 = the
                                            chain-wide value */
} tfepb_fused_layer;

typedef struct {
    const void* x; void* y; void* logdet;          /* fp32 (batch, n_features), (batch, n_features), (batch,) */
    int32_t batch, n_features;
    int32_t k1, hidden_padded;                     /* padded widths shared by all layers of the chain */
    int32_t n_layers;
    int32_t hidden_halves;                         /* 1 or 2: column halves a hidden layer is computed and handed over in */
    int32_t hidden_split[2];                       /* per hidden layer: first column of the second half (multiple of 16);
                                                      the ops of a half carry its index (0 / 1) in flag bits 2-3 and each
                                                      half ends with a commit (flag 2); every flag-16 op consumes the next
                                                      hand-over, in the order x, h1 first half, h1 second half, h2 first
                                                      half, h2 second half */
    const tfepb_fused_layer* layers;               /* HOST array of n_layers entries */
    uint32_t* tile_flags;                          /* device, (n_layers - 1) * ceil(batch / 128) words owned by the
                                                      caller and private to launches in flight; a word equal to
                                                      `epoch` marks a published tile.  May be NULL if n_layers == 1 */
    uint32_t epoch;                                /* any value the words do not hold yet (e.g. a call counter) */
    int32_t debug_mode;                            /* development only, 0 */
    int32_t mixed_splines;                         /* 1 if any feature has kind != 0 (selects the generic spline epilogue) */
    int32_t n_inputs;                              /* conditioner inputs before the two constant ones (n_features plus one
                                                      per lifted periodic feature); 0 = n_features.  k1 >= n_inputs + 2 */
    int32_t x_operand_column;                      /* first tensor-memory column of the x operand inside the 176-column A
                                                      operand region (multiple of 8; the first-layer blocks carry it in
                                                      a_col).  The hidden activations are written from column 0: with x at
                                                      the end of the region, clear of the first half of h1
                                                      (>= hidden_split[0] / 2), that half is written while the second half
                                                      of the first product still reads x */
    int32_t reserved3;
    int32_t* error_flag;                           /* device int, set if an internal wait times out; may be NULL */
    float* debug_params;                           /* NULL, or (batch, n_chunks * 112): conditioner outputs (+bias)
                                                      of layer 0 in packed order, for parity tests of the GEMM chain */
} tfepb_fused_args;
int tfepb_maf_spline_forward_bf16(const tfepb_fused_args* a, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Fused MAF INVERSE on the tensor cores, bf16 operands with fp32 accumulation, for a chain of n_layers >= 1
 * MAF layers (given in the order they are inverted) in ONE launch: x, logdet = MAF.inverse(y) with a MADE of
 * two hidden layers and a circular 8-bin neural spline, one feature per degree.  Replaces
 * AutoregressiveFlow.inverse (nn/flows/autoregressive.py:179-229; n_degrees conditioner passes) by the
 * degree-ordered sweep: per degree three small tcgen05 products whose A operands (x, h1, h2 of a 128-sample
 * tile) stay resident in tensor memory; `ops` lists one weight block per product (same entry layout and
 * block image as the forward kernel; a_col is the tensor-memory column of the operand: 64 x, 128 h1, 320 h2;
 * tmem_col the accumulator: 0 output rows, 32 hidden units), `steps` one entry per degree.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t col;                         /* column of the feature of this degree in y / x; -1: no feature, the step only
                                            brings in hidden units that depend on conditioning features alone */
    float x0, period, inv_period, rescaled_width, rescaled_height, y0;
    int32_t partner;                     /* bits 0-3: other half of the bf16 pair column of the conditioner input: 0 not
                                            known yet, 1 known, 2 constant one; bit 4: spline kind (0 circular, 1 not
                                            circular); bit 5: the feature enters the conditioner as (cos, sin) in the
                                            input columns (c, c + 1), c even (PeriodicEmbedding); bits 8-15: input
                                            column c of the feature; bits 16-23: x column of the pair partner */
    int32_t h1_first, h1_count, h2_first, h2_count;   /* packed hidden units computable after the feature (<= 15) */
} tfepb_fused_inv_step;

typedef struct {
    const tfepb_fused_op* ops;           /* DEVICE array, 16-byte aligned */
    const tfepb_fused_inv_step* steps;   /* DEVICE array, 16-byte aligned */
    int32_t n_ops, n_steps;
    const void* weights;                 /* device, packed bf16 blocks */
    float min_bin_size, min_slope, slope_offset;
    int32_t reserved;
    float emb_lower, emb_scale;          /* as in tfepb_fused_layer (used by steps with bit 5 set) */
    const int32_t* init_map;             /* NULL, or device, k1 entries in the encoding of tfepb_fused_layer.input_map: what
                                            every conditioner input column holds BEFORE the sweep -- conditioning features
                                            (degree -1, nn/flows/autoregressive.py:204), the constant ones, zero for the
                                            features still to be inverted */
} tfepb_fused_inv_layer;

typedef struct {
    const void* y; void* x; void* logdet;          /* fp32 (batch, n_features), (batch, n_features), (batch,) */
    int32_t batch, n_features;
    int32_t k1, hidden_padded;
    int32_t n_layers;
    int32_t mixed_splines;                         /* 1 if any step has the not-circular kind bit (generic spline epilogue) */
    const tfepb_fused_inv_layer* layers;           /* HOST array of n_layers entries */
    uint32_t* tile_flags;                          /* as for the forward kernel */
    uint32_t epoch;
    int32_t n_inputs;                              /* as in tfepb_fused_args; 0 = n_features */
    int32_t* error_flag;
} tfepb_fused_inv_args;
int tfepb_maf_spline_inverse_bf16(const tfepb_fused_inv_args* a, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Persistent inverse sweep of one MAF layer, exact fp32 / fp64 arithmetic: x, logdet = MAF.inverse(y).
 * Replaces AutoregressiveFlow.inverse (nn/flows/autoregressive.py:179-229), which runs n_degrees full
 * conditioner passes: with degree-sorted packed weights every unit is evaluated once, degree by degree,
 * by one persistent kernel that keeps x and all hidden activations of a sample tile in shared memory.
 * `w[l]` (n_out[l] rows, leading dimension ldw[l] = a multiple of 16 bytes, zero padded) and `b[l]` are the
 * packed effective weights of linear layer l (l = n_linear - 1 is the output layer, rows grouped by degree);
 * `groups` lists, in ascending degree, the output rows and the hidden units that become computable.
 * -------------------------------------------------------------------------------------------- */
#define TFEPB_SWEEP_MAX_LINEAR 5
enum { TFEPB_SWEEP_AFFINE = 0, TFEPB_SWEEP_SPLINE = 1, TFEPB_SWEEP_MOEBIUS = 2, TFEPB_SWEEP_SHIFT = 3 };

typedef struct {
    int32_t out_r0, out_r1, out_k;       /* output-layer rows of the group and the reduction length they need */
    int32_t h_a[TFEPB_SWEEP_MAX_LINEAR - 1], h_b[TFEPB_SWEEP_MAX_LINEAR - 1], h_k[TFEPB_SWEEP_MAX_LINEAR - 1];
                                         /* per hidden layer l (index l - 1): units [h_a, h_b) computable after
                                            the group, and their reduction length */
    int32_t part_first, part_count;      /* entries of `group_parts` */
} tfepb_sweep_group;

typedef struct {
    int32_t kind;                        /* TFEPB_SWEEP_* */
    int32_t n_bins, circular, identity_boundary_slopes, learn_lower_bound, learn_upper_bound;   /* spline */
    int32_t dimension, unit_sphere;      /* moebius (unit_sphere: variant as in tfepb_moebius) */
    const void *x0, *xf, *y0, *yf;       /* spline domain per feature of the part (dtype of the call);
                                            shift: x0 = period table (0 = not periodic), xf = lower-limit table */
    double min_bin_size, min_slope, max_radius;
    const int32_t* cols;                 /* feature -> column of x / y, or NULL */
    const int32_t* par_base;             /* feature -> first packed output row of its parameters */
} tfepb_sweep_part;

typedef struct {
    int32_t part, ids_offset, n_ids;     /* features `ids[ids_offset .. + n_ids)` of `part` belong to the group */
} tfepb_sweep_group_part;

typedef struct {
    int32_t dtype, batch, n_features, n_linear;
    const void* y; int64_t ldy;
    void* x; int64_t ldx;
    void* logdet;
    const void* w[TFEPB_SWEEP_MAX_LINEAR];
    const void* b[TFEPB_SWEEP_MAX_LINEAR];
    int32_t n_out[TFEPB_SWEEP_MAX_LINEAR], ldw[TFEPB_SWEEP_MAX_LINEAR];
    const tfepb_sweep_group* groups;             /* device */
    int32_t n_groups, max_params;                /* max_params = max (out_r1 - out_r0) */
    const tfepb_sweep_part* parts;               /* device */
    const tfepb_sweep_group_part* group_parts;   /* device */
    const int32_t* ids;                          /* device */
    const int32_t* fixed_cols;                   /* device: conditioning features copied from y (may be NULL) */
    int32_t n_fixed;
    int32_t n_embedded;                          /* width of the conditioner input: n_features, or more with an embedding */
    const int32_t* emb_out_col;                  /* device, per column of x: first conditioner-input column (NULL = no
                                                    embedding: the conditioner reads x itself) */
    const int32_t* emb_periodic;                 /* device, per column of x: 1 = lifted to (cos, sin) */
    double emb_lower, emb_scale;                 /* (x - lower) * scale is the angle of a lifted feature */
    int32_t reserved;
    int32_t max_group_weight_elems;              /* max over groups of the elements of all weight rows the group uses
                                                    (rows x leading dimension, output + hidden layers); 0 = unknown:
                                                    the kernel then reads weights through L1 instead of staging them */
    /* Blocked sweeps of wide conditioners (tfep_b200/_blocked.py: the degrees are cut into blocks, the kernel runs the
     * sequential part INSIDE a block on the block's own small weight matrices, plain GEMMs supply what earlier blocks
     * contribute): */
    const void* extra[TFEPB_SWEEP_MAX_LINEAR];   /* per linear layer l: NULL, or (batch, n_out[l]) added to the
                                                    pre-activations of the layer (before ELU / the transformer) */
    int64_t ldextra[TFEPB_SWEEP_MAX_LINEAR];
    void* act_out[TFEPB_SWEEP_MAX_LINEAR];       /* per hidden layer l >= 1: NULL, or (batch, n_out[l - 1]) receiving the
                                                    activations of the layer */
    int64_t ldact_out[TFEPB_SWEEP_MAX_LINEAR];
} tfepb_sweep_args;
int tfepb_maf_inverse_sweep(const tfepb_sweep_args* a, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * (T)FEP estimator and bootstrap
 * -------------------------------------------------------------------------------------------- */
/* One pass over v_i = scale * w_i (+ logw_i): writes out[0] = max_i v_i, out[1] = sum_i exp(v_i - max)
 * as doubles on the device.  fep_estimator = -kT (out[0] + log out[1] - log n) with scale = -1/kT
 * (analysis/estimator.py:61-86).  `partials` is a caller-owned workspace of
 * tfepb_lse_workspace_bytes() bytes.  The pair (max, sum) is what ranks exchange at N > 1. */
int64_t tfepb_lse_workspace_bytes(void);
int tfepb_lse(int32_t dtype, const void* w, const void* logw, int64_t n, double scale,
              void* partials, double* out2, tfepb_stream_t stream);
/* The estimator in one call (analysis/estimator.py:61-86): the pass above with scale = -1 / kT, and its final reduction
 * also writes result[0] = -kT ((max + log sum) - log_norm) in the dtype of w -- log_norm = log n for plain work values, 0
 * with log-weights (Bayesian bootstrap), logsumexp(bias / kT) for biased data.  out2 receives (max, sum) as in
 * tfepb_lse.  Two launches, no scalar tensor arithmetic on the host side. */
int tfepb_fep_estimate(int32_t dtype, const void* w, const void* logw, int64_t n, double kT, double log_norm,
                       void* partials, double* out2, void* result, tfepb_stream_t stream);

/* Raw MT19937 stream of torch's CPU generator (analysis/bootstrap.py:214-218 draws its indices from
 * it): fills idx[i] = u32[skip + i] % max_idx for i < count, continuing from `state` (624 words +
 * position, exactly at::mt19937's data) which is updated in place on the device. */
int tfepb_mt19937_seed(uint32_t seed, uint32_t* state625_host);
int tfepb_mt19937_indices(uint32_t* state625_dev, int64_t count, uint32_t max_idx, int32_t* idx,
                          tfepb_stream_t stream);

/* The same stream generated by all SMs (MT19937 jump-ahead, tfep_b200/csrc/mt19937_jump.cu): the request is cut into
 * at most n_streams sub-streams whose start states are obtained from `state` by polynomial jumps (t^J mod the
 * characteristic polynomial, evaluated at the state transition on the device), one CTA per sub-stream; idx and the
 * updated `state` are bit-identical to tfepb_mt19937_indices.  `workspace`: tfepb_mt19937_parallel_workspace_bytes(
 * n_streams) bytes on the device.  tfepb_mt19937_jump_polynomial (host only) returns t^steps mod phi as 19968 bits. */
int64_t tfepb_mt19937_parallel_workspace_bytes(int32_t n_streams);
int tfepb_mt19937_indices_parallel(uint32_t* state625_dev, int64_t count, uint32_t max_idx, int32_t* idx,
                                   int32_t n_streams, void* workspace, tfepb_stream_t stream);
int tfepb_mt19937_jump_polynomial(uint64_t steps, uint32_t* poly624_host);

/* Fused resample + exponential average: for resample r (row r of idx, or a counter-based Philox
 * stream when idx == NULL) out_sums[r] = sum_j e[idx[r, j]] in double, where e_i = exp(v_i - max)
 * was produced by tfepb_exp_table.  analysis/bootstrap.py:185-233 with statistic = fep_estimator.
 * `e` holds the n entries [shard_lo, shard_lo + n) of the global table (shard_lo = 0, n = everything on
 * one GPU); draws outside the shard contribute zero, so the sums of batch-sharded ranks add up.
 * `sample_sizes` (device, n_resamples entries, Philox stream only) or NULL: per-resample number of draws, at most
 * `sample_size` -- stratified resampling: a resample of n uniform draws is the same as Multinomial(n; tile sizes / n)
 * counts per table tile plus that many uniform draws INSIDE each tile, so each call can work on an L2-resident
 * tile (e + its max_idx = the tile) and no draw is generated twice. */
int tfepb_exp_table(int32_t dtype, const void* w, int64_t n, double scale, const double* max_dev,
                    float* e, tfepb_stream_t stream);
int tfepb_bootstrap_sums(const float* e, int64_t n, int64_t shard_lo, uint32_t max_idx, const int32_t* idx, int64_t ldidx,
                         int32_t n_resamples, int64_t sample_size, uint64_t philox_seed,
                         uint64_t philox_offset, const int64_t* sample_sizes, double* out_sums, tfepb_stream_t stream);
/* Bayesian bootstrap of the FEP estimator (analysis/bootstrap.py:236-262, estimator.py:78-79): per resample r
 * out_sums[r] = sum_i e[i] g_ri and out_weight_sums[r] = sum_i g_ri with g_ri ~ Exp(1) from Philox4x32-10
 * (counter = philox_offset + r * ceil(n / 4) + i / 4), i.e. Dirichlet(1, ..., 1) weights g_ri / sum_i g_ri.
 * One coalesced pass over the exp table `e` (tfepb_exp_table) per resample; n_resamples <= 65535 per call. */
int tfepb_bayesian_bootstrap_sums(const float* e, int64_t n, int32_t n_resamples, uint64_t philox_seed,
                                  uint64_t philox_offset, double* out_sums, double* out_weight_sums,
                                  tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * BoltzmannKLDivLoss (tfep/loss.py:76-140) as one streaming reduction: rw_i = target_i - log_det_J_i - ref_i
 * (loss.py:125-129), then mean_i rw_i (loss.py:140; ignore_nan: torch.nanmean, loss.py:139) or, with log_weights,
 * sum_i softmax(log_weights)_i rw_i (loss.py:132-136; ignore_nan: torch.nansum), the softmax folded into the same
 * pass as an online (max, sum e, sum e rw) triple.  log_det_J, ref_potentials and log_weights may be NULL.
 * out5 (device doubles): loss, max log-weight, sum exp(log_w - max), number of kept terms, 1 if a log-weight was NaN;
 * the backward entry point reads them back as `stats5` and writes the cotangents of the four inputs (any may be NULL)
 * given the device scalar grad_out.  `workspace`: tfepb_kl_loss_workspace_bytes() bytes.
 * -------------------------------------------------------------------------------------------- */
int64_t tfepb_kl_loss_workspace_bytes(void);
int tfepb_kl_loss(int32_t dtype, const void* target_potentials, const void* log_det_J, const void* ref_potentials,
                  const void* log_weights, int64_t n, int32_t ignore_nan, void* workspace, double* out5,
                  tfepb_stream_t stream);
int tfepb_kl_loss_backward(int32_t dtype, const void* target_potentials, const void* log_det_J,
                           const void* ref_potentials, const void* log_weights, int64_t n, int32_t ignore_nan,
                           const double* stats5, const void* grad_out, void* grad_target, void* grad_log_det_J,
                           void* grad_ref, void* grad_log_weights, tfepb_stream_t stream);

/* ----------------------------------------------------------------------------------------------
 * Fused pre / post kernels of the wrapper flows (inference): the frame change, the gather of the propagated features
 * into the contiguous tensor the wrapped flow consumes, and on the way back the scatter, the constrained coordinates
 * and the inverse frame change -- one launch each side of the flow.
 *   CenteredCentroidFlow._transform  nn/flows/centroid.py:194-263 (over PartialFlow._pass nn/flows/partial.py:88-121)
 *   OrientedFlow._transform          nn/flows/oriented.py:182-225, frame: utils/geometry.py:296-411
 * pre:  out (batch, n_propagated) = propagated features in the new frame; `shift` / `rotation` = the per-sample frame.
 * post: out (batch, n_features) from y_propagated (the flow's output), the frame and the original input x.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t dtype, batch, n_features, space_dimension;      /* space_dimension <= 4 */
    const void* x; int64_t ldx;
    const void* y_propagated; int64_t ldy;                 /* post only */
    void* out; int64_t ldout;
    void* shift;                                           /* (batch, space_dimension): origin - centroid */
    int32_t n_propagated;
    const int32_t* propagated_columns;                     /* device, n_propagated */
    const int32_t* column_to_propagated;                   /* device, n_features: position among the propagated, -1 if fixed */
    const int32_t* centroid_points; int32_t n_centroid_points;   /* device point indices defining the centroid, or NULL = all */
    const void* weights;                                   /* device, normalised weights of those points, or NULL (mean) */
    double origin[4];
    int32_t fixed_point, fixed_slot;                       /* point index of the fixed point; its position in centroid_points */
    int32_t restore_fixed_point;                           /* post: place the fixed point so that the centroid is preserved */
    int32_t translate_back;                                /* post: undo the translation */
} tfepb_centroid_args;
int tfepb_centroid_pre(const tfepb_centroid_args* a, tfepb_stream_t stream);
int tfepb_centroid_post(const tfepb_centroid_args* a, tfepb_stream_t stream);

typedef struct {
    int32_t dtype, batch, n_features, n_propagated;        /* points of 3 coordinates; n_propagated = n_features - 3 */
    const void* x; int64_t ldx;
    const void* y_propagated; int64_t ldy;                 /* post only */
    void* out; int64_t ldout;
    void* rotation;                                        /* (batch, 9) row-major rotation matrices */
    const int32_t* propagated_columns;
    const int32_t* column_to_propagated;
    int32_t axis_point, plane_point;                       /* point put on the axis / on the plane */
    int32_t axis, plane_axis;                              /* 0 / 1 / 2 = x / y / z: the axis, the second axis spanning the plane */
    int32_t round_off_imprecisions;                        /* the constrained coordinates are exactly zero in the frame */
    int32_t rotate_back;                                   /* post: rotate back into the original frame */
} tfepb_oriented_args;
int tfepb_oriented_pre(const tfepb_oriented_args* a, tfepb_stream_t stream);
int tfepb_oriented_post(const tfepb_oriented_args* a, tfepb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TFEP_B200_H */
