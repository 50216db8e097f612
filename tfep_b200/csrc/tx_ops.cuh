// Transformer device operators shared by the stand-alone kernels (transformers.cu) and the persistent
// inverse sweep (maf_inverse.cu): a view of x / y / parameters / log-det with the layout conventions of
// include/tfep_b200.h, and one operator per transformer kind built on tx_math.cuh.
#pragma once

#include "common.cuh"
#include "tx_math.cuh"

namespace tfepb {

template <typename T>
struct TxView {
    const T* x; int64_t ldx;
    T* y; int64_t ldy;
    const T* par; int64_t ldp, poff, sp, sf;
    const int* pbase;
    const int* cols;
    const int* ids;
    T* logdet;
    int accumulate, B, F, inverse;
    // gradients (backward kernels only)
    const T* gy; int64_t ldgy;
    const T* gld;
    T* gx; int64_t ldgx;
    T* gpar;

    __device__ __forceinline__ int col(int f) const { return cols ? cols[f] : f; }
    __device__ __forceinline__ int fid(int u) const { return ids ? ids[u] : u; }
    __device__ __forceinline__ int64_t poffset(int b, int f) const {
        return (int64_t)b * ldp + poff + (pbase ? (int64_t)pbase[f] : (int64_t)f * sf);
    }
    __device__ __forceinline__ ParIn<T> pin(int b, int f) const { return ParIn<T>{par + poffset(b, f), sp}; }
    __device__ __forceinline__ ParOut<T> pout(int b, int f) const { return ParOut<T>{gpar + poffset(b, f), sp}; }
};

// ------------------------------------------------------------------------------------------
template <typename T>
struct AffineOp {
    __device__ int units(int F) const { return F; }
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f = v.fid(u), c = v.col(f);
        T out, ld;
        if (v.inverse) affine_eval<T, true>(v.pin(b, f), v.x[(int64_t)b * v.ldx + c], out, ld);
        else affine_eval<T, false>(v.pin(b, f), v.x[(int64_t)b * v.ldx + c], out, ld);
        v.y[(int64_t)b * v.ldy + c] = out;
        return ld;
    }
    __device__ void backward(const TxView<T>& v, int b, int u, T gl) const {
        const int f = v.fid(u), c = v.col(f);
        T gx;
        affine_vjp<T>(v.pin(b, f), v.x[(int64_t)b * v.ldx + c], v.gy[(int64_t)b * v.ldgy + c], gl, gx, v.pout(b, f));
        v.gx[(int64_t)b * v.ldgx + c] = gx;
    }
};

template <typename T>
struct ShiftOp {
    const T* period;     // per feature of the part: period, or 0 for a non-periodic feature
    const T* lower;      // per feature: lower limit of the periodic range
    __device__ int units(int F) const { return F; }
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f = v.fid(u), c = v.col(f);
        T out;
        if (v.inverse) shift_eval<T, true>(v.pin(b, f), v.x[(int64_t)b * v.ldx + c], period[f], lower[f], out);
        else shift_eval<T, false>(v.pin(b, f), v.x[(int64_t)b * v.ldx + c], period[f], lower[f], out);
        v.y[(int64_t)b * v.ldy + c] = out;
        return T(0);
    }
    // the wrap is piecewise constant: dy/dx = 1, dy/db = 1, log-det = 0
    __device__ void backward(const TxView<T>& v, int b, int u, T) const {
        const int f = v.fid(u), c = v.col(f);
        const T gy = v.gy[(int64_t)b * v.ldgy + c];
        v.gx[(int64_t)b * v.ldgx + c] = gy;
        v.pout(b, f).set(0, gy);
    }
};

template <typename T>
struct SosOp {
    int n_poly;
    __device__ int units(int F) const { return F; }
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f = v.fid(u), c = v.col(f);
        T out, ld;
        sos_eval<T>(v.pin(b, f), n_poly, v.x[(int64_t)b * v.ldx + c], out, ld);
        v.y[(int64_t)b * v.ldy + c] = out;
        return ld;
    }
    __device__ void backward(const TxView<T>& v, int b, int u, T) const {
        const int f = v.fid(u), c = v.col(f);
        T gx;
        sos_vjp<T>(v.pin(b, f), n_poly, v.x[(int64_t)b * v.ldx + c], v.gy[(int64_t)b * v.ldgy + c], gx, v.pout(b, f));
        v.gx[(int64_t)b * v.ldgx + c] = gx;
    }
};

template <typename T>
struct MoebiusOp {
    int d;
    T max_radius;
    int unit_sphere;         // variant: 0 sphere of radius |x|, 1 unit sphere, 2 symmetrized (moebius.py:481-600)
    __device__ int units(int F) const { return F / d; }
    // Vector blocks are `d` consecutive TRANSFORMER features; their columns go through `cols`.
    // With cols == nullptr the block is contiguous in x / y (stride 1).
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f0 = v.fid(u * d);
        if (unit_sphere == 2) {
            T xs[16], vs[16], ys[16];
            for (int i = 0; i < d; ++i) {
                xs[i] = v.x[(int64_t)b * v.ldx + v.col(f0 + i)];
                vs[i] = v.par[v.poffset(b, f0 + i)];
            }
            const T ld = v.inverse ? symmoebius_inverse<T>(xs, 1, vs, 1, d, max_radius, ys, 1)
                                   : symmoebius_eval<T>(xs, 1, vs, 1, d, max_radius, ys, 1);
            for (int i = 0; i < d; ++i) v.y[(int64_t)b * v.ldy + v.col(f0 + i)] = ys[i];
            return ld;
        }
        if (v.cols == nullptr && v.pbase == nullptr) {
            return moebius_eval<T>(v.x + (int64_t)b * v.ldx + f0, 1, v.par + v.poffset(b, f0), v.sf,
                                   v.inverse ? T(-1) : T(1), d, max_radius, unit_sphere != 0,
                                   v.y + (int64_t)b * v.ldy + f0, 1);
        }
        // gather through the column map (d <= 16)
        T xs[16], vs[16], ys[16];
        for (int i = 0; i < d; ++i) {
            xs[i] = v.x[(int64_t)b * v.ldx + v.col(f0 + i)];
            vs[i] = v.par[v.poffset(b, f0 + i)];
        }
        const T ld = moebius_eval<T>(xs, 1, vs, 1, v.inverse ? T(-1) : T(1), d, max_radius, unit_sphere != 0, ys, 1);
        for (int i = 0; i < d; ++i) v.y[(int64_t)b * v.ldy + v.col(f0 + i)] = ys[i];
        return ld;
    }
    __device__ void backward(const TxView<T>& v, int b, int u, T gl) const {
        const int f0 = v.fid(u * d);
        T xs[16], vs[16], gys[16], gxs[16], gvs[16];
        for (int i = 0; i < d; ++i) {
            const int c = v.col(f0 + i);
            xs[i] = v.x[(int64_t)b * v.ldx + c];
            gys[i] = v.gy[(int64_t)b * v.ldgy + c];
            vs[i] = v.par[v.poffset(b, f0 + i)];
        }
        if (unit_sphere == 2) symmoebius_vjp<T>(xs, 1, vs, 1, d, max_radius, gys, 1, gl, gxs, 1, gvs, 1);
        else moebius_vjp<T>(xs, 1, vs, 1, d, max_radius, unit_sphere != 0, gys, 1, gl, gxs, 1, gvs, 1);
        for (int i = 0; i < d; ++i) {
            v.gx[(int64_t)b * v.ldgx + v.col(f0 + i)] = gxs[i];
            v.gpar[v.poffset(b, f0 + i)] = gvs[i];
        }
    }
};

// ------------------------------------------------------------------------------------------
// Fast variants for the PACKED parameter layout of the MAF paths (stride_p == 1: the parameters of a feature are
// consecutive): they are copied into registers once, the sizes are compile-time constants, so after inlining every
// index of the SOS / Moebius math is a constant and nothing is re-read from memory (splines: constant stride only).  The generic operators above do
// 64-bit strided address arithmetic per parameter access and re-read parameters inside their loops (ncu, r02: 200
// issued instructions per SOS feature, 1250 per spline feature, issue bound); the arithmetic itself is the same code.
// ------------------------------------------------------------------------------------------
template <typename T, int NPOLY>
struct SosPackedOp {
    static constexpr int P = 1 + 2 * NPOLY;
    __device__ int units(int F) const { return F; }
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f = v.fid(u), c = v.col(f);
        const T* p0 = v.par + v.poffset(b, f);
        T loc[P];
#pragma unroll
        for (int i = 0; i < P; ++i) loc[i] = p0[i];
        T out, ld;
        sos_eval<T>(ParIn<T>{loc, 1}, NPOLY, v.x[(int64_t)b * v.ldx + c], out, ld);
        v.y[(int64_t)b * v.ldy + c] = out;
        return ld;
    }
    __device__ void backward(const TxView<T>& v, int b, int u, T) const {
        const int f = v.fid(u), c = v.col(f);
        const int64_t off = v.poffset(b, f);
        const T* p0 = v.par + off;
        T loc[P], gloc[P];
#pragma unroll
        for (int i = 0; i < P; ++i) loc[i] = p0[i];
        T gx;
        sos_vjp<T>(ParIn<T>{loc, 1}, NPOLY, v.x[(int64_t)b * v.ldx + c], v.gy[(int64_t)b * v.ldgy + c], gx, ParOut<T>{gloc, 1});
        v.gx[(int64_t)b * v.ldgx + c] = gx;
        T* g0 = v.gpar + off;
#pragma unroll
        for (int i = 0; i < P; ++i) g0[i] = gloc[i];
    }
};

template <typename T, int D>
struct MoebiusPackedOp {                 // variants 0 / 1 (not the symmetrized map)
    T max_radius;
    int unit_sphere;
    __device__ int units(int F) const { return F / D; }
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f0 = v.fid(u * D);
        T xs[D], vs[D], ys[D];
        int cs[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            cs[i] = v.col(f0 + i);
            xs[i] = v.x[(int64_t)b * v.ldx + cs[i]];
            vs[i] = v.par[v.poffset(b, f0 + i)];
        }
        const T ld = moebius_eval<T>(xs, 1, vs, 1, v.inverse ? T(-1) : T(1), D, max_radius, unit_sphere != 0, ys, 1);
#pragma unroll
        for (int i = 0; i < D; ++i) v.y[(int64_t)b * v.ldy + cs[i]] = ys[i];
        return ld;
    }
    __device__ void backward(const TxView<T>& v, int b, int u, T gl) const {
        const int f0 = v.fid(u * D);
        T xs[D], vs[D], gys[D], gxs[D], gvs[D];
        int cs[D];
        int64_t po[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            cs[i] = v.col(f0 + i);
            po[i] = v.poffset(b, f0 + i);
            xs[i] = v.x[(int64_t)b * v.ldx + cs[i]];
            gys[i] = v.gy[(int64_t)b * v.ldgy + cs[i]];
            vs[i] = v.par[po[i]];
        }
        moebius_vjp<T>(xs, 1, vs, 1, D, max_radius, unit_sphere != 0, gys, 1, gl, gxs, 1, gvs, 1);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            v.gx[(int64_t)b * v.ldgx + cs[i]] = gxs[i];
            v.gpar[po[i]] = gvs[i];
        }
    }
};

template <typename T, int KBINS>
struct SplinePackedOp {                  // K bins, free boundary slopes, fixed limits (circular or not); the slope
                                         // parameters are indexed by the bin found at run time, so they stay in memory
    int circular;
    const T *x0, *xf, *y0, *yf;
    T min_bin, min_slope;
    __device__ int units(int F) const { return F; }
    __device__ SplineFeat<T> feat(int f) const {
        SplineFeat<T> c;
        c.K = KBINS; c.circular = circular; c.idslopes = 0; c.learn_lo = 0; c.learn_hi = 0;
        c.x0 = x0[f]; c.xf = xf[f]; c.y0 = y0[f]; c.yf = yf[f];
        c.min_bin = min_bin; c.min_slope = min_slope;
        return c;
    }
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f = v.fid(u), col = v.col(f);
        const ParIn<T> pin{v.par + v.poffset(b, f), 1};
        T out, ld;
        int bin;
        if (v.inverse) spline_eval<T, KBINS, true>(feat(f), pin, v.x[(int64_t)b * v.ldx + col], out, ld, bin);
        else spline_eval<T, KBINS, false>(feat(f), pin, v.x[(int64_t)b * v.ldx + col], out, ld, bin);
        v.y[(int64_t)b * v.ldy + col] = out;
        return ld;
    }
    __device__ void backward(const TxView<T>& v, int b, int u, T gl) const {
        const int f = v.fid(u), col = v.col(f);
        const int64_t off = v.poffset(b, f);
        T gx;
        spline_vjp<T, KBINS>(feat(f), ParIn<T>{v.par + off, 1}, v.x[(int64_t)b * v.ldx + col], v.gy[(int64_t)b * v.ldgy + col],
                             gl, gx, ParOut<T>{v.gpar + off, 1});
        v.gx[(int64_t)b * v.ldgx + col] = gx;
    }
};

template <typename T, int MAXK>
struct SplineOp {
    int K, circular, idslopes, learn_lo, learn_hi;
    const T *x0, *xf, *y0, *yf;
    T min_bin, min_slope;
    int* bins; int64_t ldbins;

    __device__ int units(int F) const { return F; }
    __device__ SplineFeat<T> feat(int f) const {
        SplineFeat<T> c;
        c.K = K; c.circular = circular; c.idslopes = idslopes; c.learn_lo = learn_lo; c.learn_hi = learn_hi;
        c.x0 = x0[f]; c.xf = xf[f]; c.y0 = y0[f]; c.yf = yf[f];
        c.min_bin = min_bin; c.min_slope = min_slope;
        return c;
    }
    __device__ T apply(const TxView<T>& v, int b, int u) const {
        const int f = v.fid(u), col = v.col(f);
        T out, ld;
        int bin;
        if (v.inverse) spline_eval<T, MAXK, true>(feat(f), v.pin(b, f), v.x[(int64_t)b * v.ldx + col], out, ld, bin);
        else spline_eval<T, MAXK, false>(feat(f), v.pin(b, f), v.x[(int64_t)b * v.ldx + col], out, ld, bin);
        v.y[(int64_t)b * v.ldy + col] = out;
        if (bins != nullptr) bins[(int64_t)b * ldbins + f] = bin;
        return ld;
    }
    __device__ void backward(const TxView<T>& v, int b, int u, T gl) const {
        const int f = v.fid(u), col = v.col(f);
        T gx;
        spline_vjp<T, MAXK>(feat(f), v.pin(b, f), v.x[(int64_t)b * v.ldx + col], v.gy[(int64_t)b * v.ldgy + col], gl, gx,
                            v.pout(b, f));
        v.gx[(int64_t)b * v.ldgx + col] = gx;
    }
};

}  // namespace tfepb
