// Per-element math of the MAF transformers: forward, inverse and vector-Jacobian products.
//
// Host+device inline functions (no CUDA runtime) so the very same code is exercised on the CPU by
// tests/hostcheck against the oracle and used by the SIMT kernels in transformers.cu.
//
// Reference lines (relative to the reference's tfep/nn/transformers/):
//   affine.py:281-363, 366-456   spline.py:319-417, 424-650   sos.py:207-306   moebius.py:374-478
#pragma once

#include "hd_math.cuh"

namespace tfepb {

// Strided view of the P parameters of one (sample, feature) pair.
template <typename T>
struct ParIn {
    const T* p;
    int64_t s;
    TFEPB_HD T operator[](int i) const { return p[(int64_t)i * s]; }
};
template <typename T>
struct ParOut {
    T* p;
    int64_t s;
    TFEPB_HD void set(int i, T v) const { p[(int64_t)i * s] = v; }
};

// ---------------------------------------------------------------------------------------------
// affine: y = x e^a + b
// ---------------------------------------------------------------------------------------------
template <typename T, bool INVERSE>
TFEPB_HD void affine_eval(const ParIn<T>& par, T t, T& out, T& ld) {
    const T shift = par[0], log_scale = par[1];
    if (!INVERSE) {
        out = t * Math<T>::exp(log_scale) + shift;
        ld = log_scale;
    } else {
        out = (t - shift) * Math<T>::exp(-log_scale);
        ld = -log_scale;
    }
}

template <typename T>
TFEPB_HD void affine_vjp(const ParIn<T>& par, T x, T gy, T gl, T& gx, const ParOut<T>& gpar) {
    const T sc = Math<T>::exp(par[1]);
    gx = gy * sc;
    gpar.set(0, gy);
    gpar.set(1, gy * x * sc + gl);
}

// ---------------------------------------------------------------------------------------------
// volume-preserving shift: y = x + b, periodic features wrapped as (x + b) % period + lower
// (affine.py:366-456; the reference does NOT subtract `lower` before the modulo -- reproduced)
// ---------------------------------------------------------------------------------------------
template <typename T, bool INVERSE>
TFEPB_HD void shift_eval(const ParIn<T>& par, T t, T period, T lower, T& out) {
    T v = INVERSE ? t - par[0] : t + par[0];
    if (period > T(0)) v = py_remainder(v, period) + lower;
    out = v;
}

// ---------------------------------------------------------------------------------------------
// sum-of-squares polynomial (L = 1): y = a0 + c1 x + c2 x^2 + c3 x^3
// ---------------------------------------------------------------------------------------------
template <typename T>
TFEPB_HD void sos_coefficients(const ParIn<T>& par, int n_poly, T& c1, T& c2, T& c3) {
    c1 = c2 = c3 = T(0);
    for (int k = 0; k < n_poly; ++k) {
        const T k0 = par[1 + 2 * k], k1 = par[2 + 2 * k];
        c1 += k0 * k0;
        c2 += k0 * k1;
        c3 += k1 * k1;
    }
    c3 = c3 / T(3);
}

template <typename T>
TFEPB_HD void sos_eval(const ParIn<T>& par, int n_poly, T x, T& y, T& ld) {
    T c1, c2, c3;
    sos_coefficients(par, n_poly, c1, c2, c3);
    // y with the same association as the reference loop (sos.py:216-226)
    const T x2 = x * x;
    const T t2 = c2 * x, t3 = c3 * x2;
    const T poly = (c1 + t2) + t3;
    y = poly * x + par[0];
    // dy/dx = c1 + 2 c2 x + 3 c3 x^2 is evaluated in its sum-of-squares form sum_k (a_k0 + a_k1 x)^2:
    // identical in exact arithmetic, but never negative (the expanded form can cancel to <= 0 in fp32
    // when the polynomials nearly vanish, which turns the log-det into NaN).
    T dydx = T(0);
    for (int k = 0; k < n_poly; ++k) {
        const T q = par[1 + 2 * k] + par[2 + 2 * k] * x;
        dydx += q * q;
    }
    ld = Math<T>::log(dydx);
}

// Reference backward (sos.py:237-268): the log-det cotangent is NOT propagated.
template <typename T>
TFEPB_HD void sos_vjp(const ParIn<T>& par, int n_poly, T x, T gy, T& gx, const ParOut<T>& gpar) {
    T c1, c2, c3;
    sos_coefficients(par, n_poly, c1, c2, c3);
    const T x2 = x * x, x3 = x2 * x;
    T dydx = T(0);
    for (int k = 0; k < n_poly; ++k) {
        const T q = par[1 + 2 * k] + par[2 + 2 * k] * x;
        dydx += q * q;
    }
    gx = dydx * gy;
    gpar.set(0, gy);
    for (int k = 0; k < n_poly; ++k) {
        const T k0 = par[1 + 2 * k], k1 = par[2 + 2 * k];
        gpar.set(1 + 2 * k, (k1 * x2 + T(2) * k0 * x) * gy);
        gpar.set(2 + 2 * k, (T(2) / T(3) * k1 * x3 + k0 * x2) * gy);
    }
}

// ---------------------------------------------------------------------------------------------
// Moebius on the sphere of radius |x| (vector blocks of `d` consecutive features)
//
// log|det J| of the reference's explicit d x d Jacobian (moebius.py:463-476) has the closed form
//   (d - 1) log c   with c = (|x|^2 - |w|^2) / |x - w|^2            (general radius)
//    d      log c                                                  (unit_sphere=True: no projection)
// because J maps the tangent space at x conformally (factor c, times a reflection) onto the tangent
// space at y and J x = y with |y| = |x|.  Verified against slogdet to 2e-15.
// ---------------------------------------------------------------------------------------------
template <typename T>
TFEPB_HD T moebius_eval(const T* x, int64_t sx, const T* v, int64_t sv, T sign, int d, T max_radius,
                        bool unit_sphere, T* y, int64_t sy) {
    T n2 = T(0), r2 = T(0);
    for (int i = 0; i < d; ++i) {
        const T vi = v[i * sv], xi = x[i * sx];
        n2 += vi * vi;
        r2 += xi * xi;
    }
    const T n = Math<T>::sqrt(n2);
    T resc = max_radius / (T(1) + n);
    T r = T(1);
    if (!unit_sphere) {
        r = Math<T>::sqrt(r2);
        resc = r * resc;
    }
    const T wn = resc * n;
    const T num = unit_sphere ? (T(1) - wn * wn) : (r * r - wn * wn);
    T d2 = T(0);
    for (int i = 0; i < d; ++i) {
        const T di = x[i * sx] - resc * (sign * v[i * sv]);
        d2 += di * di;
    }
    const T dn = Math<T>::sqrt(d2);
    const T c = num / (dn * dn);
    for (int i = 0; i < d; ++i) {
        const T wi = resc * (sign * v[i * sv]);
        y[i * sy] = c * (x[i * sx] - wi) - wi;
    }
    return T(unit_sphere ? d : d - 1) * Math<T>::log(c);
}

// VJP of the forward map (sign = +1).  gy: cotangent of y (d values), gl: cotangent of the log-det.
template <typename T>
TFEPB_HD void moebius_vjp(const T* x, int64_t sx, const T* v, int64_t sv, int d, T max_radius, bool unit_sphere,
                          const T* gy, int64_t sgy, T gl, T* gx, int64_t sgx, T* gv, int64_t sgv) {
    T n2 = T(0), r2 = T(0);
    for (int i = 0; i < d; ++i) {
        n2 += v[i * sv] * v[i * sv];
        r2 += x[i * sx] * x[i * sx];
    }
    const T n = Math<T>::sqrt(n2);
    const T alpha = max_radius / (T(1) + n);
    const T beta = alpha * n;
    const T r = unit_sphere ? T(1) : Math<T>::sqrt(r2);
    const T R2 = unit_sphere ? T(1) : r2;
    const T rho = r * alpha;
    const T num = R2 * (T(1) - beta * beta);
    T D2 = T(0), gy_delta = T(0), gy_v = T(0);
    for (int i = 0; i < d; ++i) {
        const T di = x[i * sx] - rho * v[i * sv];
        D2 += di * di;
        gy_delta += gy[i * sgy] * di;
        gy_v += gy[i * sgy] * v[i * sv];
    }
    const T c = num / D2;
    const T ldf = T(unit_sphere ? d : d - 1);
    const T gc = gy_delta + gl * ldf / c;          // cotangent of c
    const T gnum = gc / D2;
    const T gD2 = -gc * c / D2;
    // delta_bar_i = c gy_i + 2 gD2 delta_i ; w_bar_i = -gy_i - delta_bar_i
    T wbar_v = T(0);                               // sum_i w_bar_i v_i  (cotangent of rho)
    for (int i = 0; i < d; ++i) {
        const T di = x[i * sx] - rho * v[i * sv];
        const T dbar = c * gy[i * sgy] + T(2) * gD2 * di;
        wbar_v += (-gy[i * sgy] - dbar) * v[i * sv];
    }
    const T grho = wbar_v;
    const T galpha = grho * r;
    const T gbeta = T(-2) * R2 * beta * gnum;
    const T gn = (gbeta - galpha) * alpha / (T(1) + n);
    T gR2 = T(0);
    if (!unit_sphere) gR2 = gnum * (T(1) - beta * beta) + (r > T(0) ? grho * alpha / (T(2) * r) : T(0));
    const T gn_over_n = n > T(0) ? gn / n : T(0);
    for (int i = 0; i < d; ++i) {
        const T di = x[i * sx] - rho * v[i * sv];
        const T dbar = c * gy[i * sgy] + T(2) * gD2 * di;
        const T wbar = -gy[i * sgy] - dbar;
        gx[i * sgx] = dbar + T(2) * gR2 * x[i * sx];
        gv[i * sgv] = rho * wbar + gn_over_n * v[i * sv];
    }
    (void)gy_v;
}

// ---------------------------------------------------------------------------------------------
// Symmetrized Moebius (reference moebius.py:481-629): y = |x| s / |s| with s = f(x; w) + f(x; -w), f the Moebius map on
// the sphere of radius |x|.  The log-det has the closed form of the reference (moebius.py:607-629)
//   log[(1 - r2) (1 + r2)^(d-1) / (4 (r2 - t^2) + (1 - r2)^2)^(d/2)],  r2 = |w_u|^2, t = (x / |x|) . w_u,
//   w_u = max_radius / (1 + |w|) w  (the parameter vector inside the unit ball).
// d <= 16.
// ---------------------------------------------------------------------------------------------
template <typename T>
TFEPB_HD T symmoebius_logdet(T r2, T t, int d) {
    const T q = r2 - t * t;
    const T om = T(1) - r2;
    const T E = T(4) * q + om * om;
    return Math<T>::log(om) + T(d - 1) * Math<T>::log(T(1) + r2) - T(0.5) * T(d) * Math<T>::log(E);
}

template <typename T>
TFEPB_HD T symmoebius_eval(const T* x, int64_t sx, const T* v, int64_t sv, int d, T max_radius, T* y, int64_t sy) {
    T f1[16], f2[16];
    moebius_eval<T>(x, sx, v, sv, T(1), d, max_radius, false, f1, 1);
    moebius_eval<T>(x, sx, v, sv, T(-1), d, max_radius, false, f2, 1);
    T s2 = T(0), r2x = T(0), n2 = T(0), xv = T(0);
    for (int i = 0; i < d; ++i) {
        f1[i] += f2[i];
        s2 += f1[i] * f1[i];
        r2x += x[i * sx] * x[i * sx];
        n2 += v[i * sv] * v[i * sv];
        xv += x[i * sx] * v[i * sv];
    }
    const T r = Math<T>::sqrt(r2x);
    const T scale = r / Math<T>::sqrt(s2);
    for (int i = 0; i < d; ++i) y[i * sy] = scale * f1[i];
    const T n = Math<T>::sqrt(n2);
    const T alpha = max_radius / (T(1) + n);
    const T rho = alpha * n;
    return symmoebius_logdet<T>(rho * rho, alpha * xv / r, d);
}

// analytic inverse (moebius.py:553-600): in the plane spanned by w and x the map acts on the angle only
template <typename T>
TFEPB_HD T symmoebius_inverse(const T* yv, int64_t sy, const T* v, int64_t sv, int d, T max_radius, T* x, int64_t sx) {
    T r2y = T(0), n2 = T(0);
    for (int i = 0; i < d; ++i) {
        r2y += yv[i * sy] * yv[i * sy];
        n2 += v[i * sv] * v[i * sv];
    }
    const T r = Math<T>::sqrt(r2y), n = Math<T>::sqrt(n2);
    const T alpha = max_radius / (T(1) + n);
    const T rho = alpha * n;
    // da = w / |w| ; a = y_unit . da ; db = (y_unit - a da) / b
    T a = T(0);
    for (int i = 0; i < d; ++i) a += (yv[i * sy] / r) * (v[i * sv] / n);
    T db[16], b2 = T(0);
    for (int i = 0; i < d; ++i) {
        db[i] = yv[i * sy] / r - a * (v[i * sv] / n);
        b2 += db[i] * db[i];
    }
    const T b = Math<T>::sqrt(b2);
    const T r2 = rho * rho;
    const T a_inv = -a * (r2 + T(1)) / Math<T>::sqrt(T(1) + r2 * r2 + r2 * (T(4) * a * a - T(2)));
    const T b_inv = -Math<T>::sqrt(T(1) - a_inv * a_inv);
    T t = T(0);                                        // x_unit_inv . w_u
    for (int i = 0; i < d; ++i) {
        const T u = -(a_inv * (v[i * sv] / n) + b_inv * (db[i] / b));
        t += u * alpha * v[i * sv];
        x[i * sx] = r * u;
    }
    return -symmoebius_logdet<T>(r2, t, d);
}

// VJP of the forward map: through the two Moebius maps, the projection back on the sphere and the closed-form log-det
template <typename T>
TFEPB_HD void symmoebius_vjp(const T* x, int64_t sx, const T* v, int64_t sv, int d, T max_radius,
                             const T* gy, int64_t sgy, T gl, T* gx, int64_t sgx, T* gv, int64_t sgv) {
    T s[16], f2[16], xs[16], vs[16], vneg[16];
    for (int i = 0; i < d; ++i) {
        xs[i] = x[i * sx];
        vs[i] = v[i * sv];
        vneg[i] = -vs[i];
    }
    moebius_eval<T>(xs, 1, vs, 1, T(1), d, max_radius, false, s, 1);
    moebius_eval<T>(xs, 1, vs, 1, T(-1), d, max_radius, false, f2, 1);
    T s2 = T(0), r2x = T(0), n2 = T(0), xv = T(0), sg = T(0);
    for (int i = 0; i < d; ++i) {
        s[i] += f2[i];
        s2 += s[i] * s[i];
        r2x += xs[i] * xs[i];
        n2 += vs[i] * vs[i];
        xv += xs[i] * vs[i];
    }
    const T sn = Math<T>::sqrt(s2), r = Math<T>::sqrt(r2x), n = Math<T>::sqrt(n2);
    for (int i = 0; i < d; ++i) sg += (s[i] / sn) * gy[i * sgy];          // s_hat . gy = cotangent of r
    T gs[16];
    for (int i = 0; i < d; ++i) gs[i] = (r / sn) * (gy[i * sgy] - (s[i] / sn) * sg);
    T gx1[16], gv1[16], gx2[16], gv2[16];
    moebius_vjp<T>(xs, 1, vs, 1, d, max_radius, false, gs, 1, T(0), gx1, 1, gv1, 1);
    moebius_vjp<T>(xs, 1, vneg, 1, d, max_radius, false, gs, 1, T(0), gx2, 1, gv2, 1);
    // log-det: ld(r2, t), r2 = rho^2, rho = alpha n, t = alpha (x . v) / r
    const T alpha = max_radius / (T(1) + n);
    const T rho = alpha * n, r2 = rho * rho;
    const T t = alpha * xv / r;
    const T om = T(1) - r2, E = T(4) * (r2 - t * t) + om * om;
    const T dld_dr2 = -T(1) / om + T(d - 1) / (T(1) + r2) - T(0.5) * T(d) * (T(4) - T(2) * om) / E;
    const T dld_dt = T(4) * T(d) * t / E;
    const T drho_dn = alpha / (T(1) + n), dalpha_dn = -alpha / (T(1) + n);
    const T gn = gl * (dld_dr2 * T(2) * rho * drho_dn + dld_dt * dalpha_dn * xv / r);     // cotangent of n = |v|
    const T gt = gl * dld_dt;
    for (int i = 0; i < d; ++i) {
        gx[i * sgx] = gx1[i] + gx2[i] + sg * xs[i] / r + gt * alpha * (vs[i] / r - xv * xs[i] / (r * r * r));
        gv[i * sgv] = gv1[i] - gv2[i] + gt * alpha * xs[i] / r + (n > T(0) ? gn * vs[i] / n : T(0));
    }
}

// ---------------------------------------------------------------------------------------------
// rational-quadratic neural spline
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SplineFeat {
    int K;
    int circular, idslopes, learn_lo, learn_hi;
    T x0, xf, y0, yf, min_bin, min_slope;
    TFEPB_HD int n_params() const {
        return 3 * K + 1 + (learn_lo ? 1 : 0) + (learn_hi ? 1 : 0) - (idslopes ? (circular ? 1 : 2) : 0);
    }
    // parameter index feeding the slope at knot j (0..K), or -1 for a fixed (identity) slope;
    // spline.py:353-380
    TFEPB_HD int slope_param(int j) const {
        if (idslopes) return (j == 0 || j == K) ? -1 : 2 * K + j - 1;
        if (circular) return 2 * K + (j == K ? 0 : j);
        return 2 * K + j;
    }
};

template <typename T, int MAXK>
struct SplineState {
    T ew[MAXK], eh[MAXK];   // softmax probabilities
    T w[MAXK], h[MAXK];     // normalised widths / heights
    T cw[MAXK], ch[MAXK];   // prefix sums (accumulated in double, each prefix rounded to T like torch.cumsum on CPU)
    T x0, y0, Rw, Rh, base_w, base_h, scale, offset, dx, min_interval;
    int P;
};

template <typename T, int MAXK>
TFEPB_HD void spline_setup(const SplineFeat<T>& c, const ParIn<T>& par, SplineState<T, MAXK>& st) {
    const int K = c.K;
    st.P = c.n_params();
    st.min_interval = T(K) * c.min_bin;
    st.base_w = c.xf - c.x0 - st.min_interval;
    st.base_h = c.yf - c.y0 - st.min_interval;
    st.scale = T(1);
    st.Rw = st.base_w;
    st.Rh = st.base_h;
    if (c.learn_lo || c.learn_hi) {
        st.scale = Math<T>::exp(par[st.P - 1]);
        st.Rw = st.base_w * st.scale;
        st.Rh = st.base_h * st.scale;
    }
    T mw = par[0], mh = par[K];
#pragma unroll
    for (int k = 1; k < MAXK; ++k)
        if (k < K) {
            const T a = par[k], b = par[K + k];
            mw = a > mw ? a : mw;
            mh = b > mh ? b : mh;
        }
    T sw = T(0), sh = T(0);
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k < K) {
            st.ew[k] = Math<T>::exp(par[k] - mw);
            st.eh[k] = Math<T>::exp(par[K + k] - mh);
            sw += st.ew[k];
            sh += st.eh[k];
        }
    double aw = 0.0, ah = 0.0;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k < K) {
            st.ew[k] = st.ew[k] / sw;
            st.eh[k] = st.eh[k] / sh;
            st.w[k] = st.ew[k] * st.Rw + c.min_bin;
            st.h[k] = st.eh[k] * st.Rh + c.min_bin;
            aw += (double)st.w[k];
            ah += (double)st.h[k];
            st.cw[k] = (T)aw;
            st.ch[k] = (T)ah;
        }
    st.x0 = c.x0;
    st.y0 = c.y0;
    if (c.learn_lo && c.learn_hi) {
        const T sh2 = par[st.P - 2];
        st.x0 = c.x0 + sh2;
        st.y0 = c.y0 + sh2;
    } else if (c.learn_lo) {
        st.x0 = c.xf - st.Rw - st.min_interval;
        st.y0 = c.yf - st.Rh - st.min_interval;
    }
    st.offset = Math<T>::log(Math<T>::exp(T(1) - c.min_slope) - T(1));
    T total = T(0);
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k == K - 1) total = st.cw[k];
    st.dx = total * T(1000);
}

template <typename T, int MAXK>
TFEPB_HD T spline_slope(const SplineFeat<T>& c, const ParIn<T>& par, const SplineState<T, MAXK>& st, int j) {
    const int i = c.slope_param(j);
    const T raw = i < 0 ? T(0) : par[i];
    return softplus(raw + st.offset) + c.min_slope;
}

template <typename T, int MAXK>
TFEPB_HD T pick(const T (&a)[MAXK], int i) {
    T v = T(0);
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k == i) v = a[k];
    return v;
}

// log(dy/dx) of the rational-quadratic segment; spline.py:546-564.
template <typename T>
TFEPB_HD T rq_log_derivative(T s, T dk, T dk1, T e, T u, T e2) {
    const T ome = T(1) - e;
    const T num = (s * s) * (dk1 * e2 + T(2) * s * u + dk * (ome * ome));
    const T dq = s + (dk1 + dk - T(2) * s) * u;
    return Math<T>::log(num / (dq * dq));
}

// Bin of t among the K+3 knots (0 = left tail, 1..K, K+1 = right tail); spline.py:622-625.
template <typename T, int MAXK>
TFEPB_HD int spline_bin(const SplineFeat<T>& c, const SplineState<T, MAXK>& st, T t, bool inverse, T d0, T dK) {
    const int K = c.K;
    const T lo = inverse ? st.y0 : st.x0;
    const T tail0 = inverse ? d0 * st.dx : st.dx;
    const T tailK = inverse ? dK * st.dx : st.dx;
    int cnt = (t > lo - tail0) + (t > lo);
    T last = lo;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k < K) {
            last = lo + (inverse ? st.ch[k] : st.cw[k]);
            cnt += (t > last);
        }
    cnt += (t > last + tailK);
    const int bin = cnt - 1;
    return bin < 0 ? 0 : bin;      // t at or below the far-left knot: outside the reference's contract
}

struct SplineSel {
    int bin;
};

template <typename T, int MAXK, bool INVERSE>
TFEPB_HD void spline_eval(const SplineFeat<T>& c, const ParIn<T>& par, T t, T& out, T& ld, int& bin_out) {
    SplineState<T, MAXK> st;
    spline_setup<T, MAXK>(c, par, st);
    const int K = c.K;
    const T shift = c.circular ? par[st.P - 1] : T(0);
    if (!INVERSE && c.circular) t = py_remainder(t - st.x0 + shift, c.xf - st.x0) + st.x0;
    const T d0 = spline_slope<T, MAXK>(c, par, st, 0);
    const T dK = spline_slope<T, MAXK>(c, par, st, K);
    const int bin = spline_bin<T, MAXK>(c, st, t, INVERSE, d0, dK);
    bin_out = bin;

    T ws, hs, xk, yk, dk, dk1;
    if (bin == 0) {
        ws = st.dx; hs = d0 * st.dx;
        xk = st.x0 - st.dx; yk = st.y0 - hs;
        dk = dk1 = d0;
    } else if (bin == K + 1) {
        ws = st.dx; hs = dK * st.dx;
        xk = st.x0 + pick<T, MAXK>(st.cw, K - 1); yk = st.y0 + pick<T, MAXK>(st.ch, K - 1);
        dk = dk1 = dK;
    } else {
        const int b = bin - 1;
        ws = pick<T, MAXK>(st.w, b); hs = pick<T, MAXK>(st.h, b);
        xk = b == 0 ? st.x0 : st.x0 + pick<T, MAXK>(st.cw, b - 1);
        yk = b == 0 ? st.y0 : st.y0 + pick<T, MAXK>(st.ch, b - 1);
        dk = b == 0 ? d0 : spline_slope<T, MAXK>(c, par, st, b);
        dk1 = b + 1 == K ? dK : spline_slope<T, MAXK>(c, par, st, b + 1);
    }
    // Far tails: the reference evaluates the generic formula on a fake bin of width 1000 W
    // (spline.py:596-614), which is exactly linear in exact arithmetic but loses ~3 digits in fp32
    // (SURVEY.md Appendix C-11).  The analytically identical linear map is evaluated instead.
    if (bin == 0 || bin == K + 1) {
        const T xe = bin == 0 ? st.x0 : xk;      // edge knot of the spline domain
        const T ye = bin == 0 ? st.y0 : yk;
        if (!INVERSE) {
            out = ye + dk * (t - xe);
            ld = Math<T>::log(dk);
        } else {
            T x = xe + (t - ye) / dk;
            ld = -Math<T>::log(dk);
            if (c.circular) x = py_remainder(x - st.x0 - shift, c.xf - st.x0) + st.x0;
            out = x;
        }
        return;
    }
    const T s = hs / ws;
    if (!INVERSE) {
        const T e = (t - xk) / ws;
        const T u = e * (T(1) - e), e2 = e * e;
        const T num = hs * (s * e2 + dk * u);
        const T den = s + (dk1 + dk - T(2) * s) * u;
        out = yk + num / den;
        ld = rq_log_derivative(s, dk, dk1, e, u, e2);
    } else {
        const T yr = t - yk;
        const T q = dk1 + dk - T(2) * s;
        const T a = hs * (s - dk) + yr * q;
        const T b = hs * dk - yr * q;
        const T cc = -s * yr;
        const T e = T(2) * cc / (-b - Math<T>::sqrt(b * b - T(4) * a * cc));
        T x = e * ws + xk;
        ld = -rq_log_derivative(s, dk, dk1, e, e * (T(1) - e), e * e);
        if (c.circular) x = py_remainder(x - st.x0 - shift, c.xf - st.x0) + st.x0;
        out = x;
    }
}

// VJP of the forward direction.  Bin index and the modulo wrap are piecewise constant; the far
// tails are analytically linear (y = y_edge + slope (x - x_edge)), which is what the reference's
// generic formula evaluates to in exact arithmetic.
template <typename T, int MAXK>
TFEPB_HD void spline_vjp(const SplineFeat<T>& c, const ParIn<T>& par, T x, T gy, T gl, T& gx, const ParOut<T>& gpar) {
    SplineState<T, MAXK> st;
    spline_setup<T, MAXK>(c, par, st);
    const int K = c.K;
    const T shift = c.circular ? par[st.P - 1] : T(0);
    T t = x;
    if (c.circular) t = py_remainder(x - st.x0 + shift, c.xf - st.x0) + st.x0;
    const T d0 = spline_slope<T, MAXK>(c, par, st, 0);
    const T dK = spline_slope<T, MAXK>(c, par, st, K);
    const int bin = spline_bin<T, MAXK>(c, st, t, false, d0, dK);

    T gw[MAXK], gh[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) gw[k] = gh[k] = T(0);
    T gt = T(0), gx0 = T(0), gy0 = T(0);
    int ja = -1, jb = -1;          // knots whose slopes receive gradient
    T gda = T(0), gdb = T(0);

    if (bin == 0) {
        gt = gy * d0;
        ja = 0; gda = gy * (t - st.x0) + gl / d0;
        gx0 = -gy * d0; gy0 = gy;
    } else if (bin == K + 1) {
        const T xK = st.x0 + pick<T, MAXK>(st.cw, K - 1);
        gt = gy * dK;
        ja = K; gda = gy * (t - xK) + gl / dK;
        gx0 = -gy * dK; gy0 = gy;
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k < K) { gw[k] = -gy * dK; gh[k] = gy; }
    } else {
        const int b = bin - 1;
        const T ws = pick<T, MAXK>(st.w, b), hs = pick<T, MAXK>(st.h, b);
        const T xk = b == 0 ? st.x0 : st.x0 + pick<T, MAXK>(st.cw, b - 1);
        const T dk = b == 0 ? d0 : spline_slope<T, MAXK>(c, par, st, b);
        const T dk1 = b + 1 == K ? dK : spline_slope<T, MAXK>(c, par, st, b + 1);
        const T s = hs / ws;
        const T e = (t - xk) / ws;
        const T ome = T(1) - e, u = e * ome, e2 = e * e;
        const T q = dk1 + dk - T(2) * s;
        const T Pn = s * e2 + dk * u;
        const T Q = s + q * u;
        const T Nn = dk1 * e2 + T(2) * s * u + dk * ome * ome;
        const T Pe = T(2) * s * e + dk * (T(1) - T(2) * e);
        const T Qe = q * (T(1) - T(2) * e);
        const T iQ = T(1) / Q, iQ2 = iQ * iQ, iN = T(1) / Nn;
        const T y_e = hs * (Pe * Q - Pn * Qe) * iQ2;
        const T y_s = hs * (e2 * Q - Pn * (T(1) - T(2) * u)) * iQ2;
        const T y_dk = hs * u * (Q - Pn) * iQ2;
        const T y_dk1 = -hs * Pn * u * iQ2;
        const T l_e = (T(2) * dk1 * e + T(2) * s * (T(1) - T(2) * e) - T(2) * dk * ome) * iN - T(2) * Qe * iQ;
        const T l_s = T(2) / s + T(2) * u * iN - T(2) * (T(1) - T(2) * u) * iQ;
        const T l_dk = ome * ome * iN - T(2) * u * iQ;
        const T l_dk1 = e2 * iN - T(2) * u * iQ;
        const T ge = gy * y_e + gl * l_e;
        const T gs = gy * y_s + gl * l_s;
        gt = ge / ws;
        const T gxk = -ge / ws;
        const T gws = -ge * e / ws - gs * s / ws;
        const T ghs = gy * Pn * iQ + gs / ws;
        ja = b; gda = gy * y_dk + gl * l_dk;
        jb = b + 1; gdb = gy * y_dk1 + gl * l_dk1;
        gx0 = gxk; gy0 = gy;
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k < K) {
                gw[k] = (k < b ? gxk : T(0)) + (k == b ? gws : T(0));
                gh[k] = (k < b ? gy : T(0)) + (k == b ? ghs : T(0));
            }
    }

    // widths / heights -> softmax logits, domain scale
    T gRw = T(0), gRh = T(0), dotw = T(0), doth = T(0);
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k < K) {
            gRw += gw[k] * st.ew[k];
            gRh += gh[k] * st.eh[k];
        }
    dotw = gRw * st.Rw;     // sum_j (gw_j Rw) ew_j
    doth = gRh * st.Rh;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
        if (k < K) {
            gpar.set(k, st.ew[k] * (gw[k] * st.Rw - dotw));
            gpar.set(K + k, st.eh[k] * (gh[k] * st.Rh - doth));
        }
    // slopes (the parameters are read BEFORE the first cotangent of their range is written: gpar may alias par, as in the
    // fused epilogue of tc_gemm_sm100.cu where both live in the same shared-memory column)
    {
        // accumulate in parameter space (circular splines tie knot K to knot 0)
        const int ia = ja >= 0 ? c.slope_param(ja) : -1;
        const int ib = jb >= 0 ? c.slope_param(jb) : -1;
        T ga = T(0), gb = T(0);
        if (ia >= 0) ga = gda * (par[ia] + st.offset > T(20) ? T(1) : sigmoid(par[ia] + st.offset));
        if (ib >= 0) gb = gdb * (par[ib] + st.offset > T(20) ? T(1) : sigmoid(par[ib] + st.offset));
        for (int i = 2 * K; i < st.P; ++i) gpar.set(i, T(0));
        if (ia >= 0 && ia == ib) {
            gpar.set(ia, ga + gb);
        } else {
            if (ia >= 0) gpar.set(ia, ga);
            if (ib >= 0) gpar.set(ib, gb);
        }
    }
    // learnable domain / circular shift
    if (c.learn_lo && c.learn_hi) {
        gpar.set(st.P - 2, gx0 + gy0);
    } else if (c.learn_lo) {
        gRw -= gx0;
        gRh -= gy0;
    }
    if (c.learn_lo || c.learn_hi) gpar.set(st.P - 1, (gRw * st.base_w + gRh * st.base_h) * st.scale);
    if (c.circular) gpar.set(st.P - 1, gt);
    gx = gt;
}

}  // namespace tfepb
