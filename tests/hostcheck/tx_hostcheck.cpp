// TEST-ONLY: compiles the device math headers (tfep_b200/csrc/tx_math.cuh) for the host with g++ so
// that the `-m "not gpu"` suite can check the very formulas the CUDA kernels run against the
// oracle, without a GPU.  Nothing in the product library links or loads this file.
#include <stdint.h>

#include "../../tfep_b200/csrc/tx_math.cuh"

using namespace tfepb;

namespace {

template <typename T>
struct Io {
    int B, F, P;
    const T* x; const T* par; T* y; T* ld;
    const T* gy; const T* gl; T* gx; T* gpar;
    ParIn<T> pin(int b, int f) const { return ParIn<T>{par + (int64_t)b * P * F + f, F}; }
    ParOut<T> pout(int b, int f) const { return ParOut<T>{gpar + (int64_t)b * P * F + f, F}; }
};

template <typename T>
void affine(const Io<T>& io, int inverse, int backward) {
    for (int b = 0; b < io.B; ++b) {
        T acc = 0;
        for (int f = 0; f < io.F; ++f) {
            const int64_t i = (int64_t)b * io.F + f;
            if (backward) {
                affine_vjp<T>(io.pin(b, f), io.x[i], io.gy[i], io.gl[b], io.gx[i], io.pout(b, f));
            } else {
                T out, ld;
                if (inverse) affine_eval<T, true>(io.pin(b, f), io.x[i], out, ld);
                else affine_eval<T, false>(io.pin(b, f), io.x[i], out, ld);
                io.y[i] = out;
                acc += ld;
            }
        }
        if (!backward) io.ld[b] = acc;
    }
}

template <typename T>
void shift(const Io<T>& io, const T* period, const T* lower, int inverse, int backward) {
    for (int b = 0; b < io.B; ++b) {
        for (int f = 0; f < io.F; ++f) {
            const int64_t i = (int64_t)b * io.F + f;
            if (backward) {
                io.gx[i] = io.gy[i];
                io.pout(b, f).set(0, io.gy[i]);
            } else {
                T out;
                if (inverse) shift_eval<T, true>(io.pin(b, f), io.x[i], period[f], lower[f], out);
                else shift_eval<T, false>(io.pin(b, f), io.x[i], period[f], lower[f], out);
                io.y[i] = out;
            }
        }
        if (!backward) io.ld[b] = 0;
    }
}

template <typename T>
void sos(const Io<T>& io, int n_poly, int backward) {
    for (int b = 0; b < io.B; ++b) {
        T acc = 0;
        for (int f = 0; f < io.F; ++f) {
            const int64_t i = (int64_t)b * io.F + f;
            if (backward) {
                sos_vjp<T>(io.pin(b, f), n_poly, io.x[i], io.gy[i], io.gx[i], io.pout(b, f));
            } else {
                T out, ld;
                sos_eval<T>(io.pin(b, f), n_poly, io.x[i], out, ld);
                io.y[i] = out;
                acc += ld;
            }
        }
        if (!backward) io.ld[b] = acc;
    }
}

template <typename T>
void moebius(const Io<T>& io, int d, double max_radius, int unit, int inverse, int backward) {
    for (int b = 0; b < io.B; ++b) {
        T acc = 0;
        for (int u = 0; u < io.F / d; ++u) {
            const int64_t i = (int64_t)b * io.F + u * d;
            if (unit == 2) {          // symmetrized variant
                if (backward)
                    symmoebius_vjp<T>(io.x + i, 1, io.par + i, 1, d, (T)max_radius, io.gy + i, 1, io.gl[b], io.gx + i, 1,
                                      io.gpar + i, 1);
                else
                    acc += inverse ? symmoebius_inverse<T>(io.x + i, 1, io.par + i, 1, d, (T)max_radius, io.y + i, 1)
                                   : symmoebius_eval<T>(io.x + i, 1, io.par + i, 1, d, (T)max_radius, io.y + i, 1);
            } else if (backward)
                moebius_vjp<T>(io.x + i, 1, io.par + i, 1, d, (T)max_radius, unit != 0, io.gy + i, 1, io.gl[b],
                               io.gx + i, 1, io.gpar + i, 1);
            else
                acc += moebius_eval<T>(io.x + i, 1, io.par + i, 1, inverse ? T(-1) : T(1), d, (T)max_radius, unit != 0,
                                       io.y + i, 1);
        }
        if (!backward) io.ld[b] = acc;
    }
}

template <typename T, int MAXK>
void spline(const Io<T>& io, int K, int circular, int idslopes, int learn_lo, int learn_hi, const T* x0, const T* xf,
            const T* y0, const T* yf, double min_bin, double min_slope, int inverse, int backward, int* bins) {
    for (int b = 0; b < io.B; ++b) {
        T acc = 0;
        for (int f = 0; f < io.F; ++f) {
            SplineFeat<T> c;
            c.K = K; c.circular = circular; c.idslopes = idslopes; c.learn_lo = learn_lo; c.learn_hi = learn_hi;
            c.x0 = x0[f]; c.xf = xf[f]; c.y0 = y0[f]; c.yf = yf[f];
            c.min_bin = (T)min_bin; c.min_slope = (T)min_slope;
            const int64_t i = (int64_t)b * io.F + f;
            if (backward) {
                spline_vjp<T, MAXK>(c, io.pin(b, f), io.x[i], io.gy[i], io.gl[b], io.gx[i], io.pout(b, f));
            } else {
                T out, ld;
                int bin;
                if (inverse) spline_eval<T, MAXK, true>(c, io.pin(b, f), io.x[i], out, ld, bin);
                else spline_eval<T, MAXK, false>(c, io.pin(b, f), io.x[i], out, ld, bin);
                io.y[i] = out;
                acc += ld;
                if (bins) bins[i] = bin;
            }
        }
        if (!backward) io.ld[b] = acc;
    }
}

template <typename T>
Io<T> mk(int B, int F, int P, const void* x, const void* par, void* y, void* ld, const void* gy, const void* gl,
         void* gx, void* gpar) {
    return Io<T>{B, F, P, (const T*)x, (const T*)par, (T*)y, (T*)ld, (const T*)gy, (const T*)gl, (T*)gx, (T*)gpar};
}

}  // namespace

#define IO_ARGS int f64, int B, int F, int P, const void* x, const void* par, void* y, void* ld, const void* gy, \
                const void* gl, void* gx, void* gpar
#define IO(T) mk<T>(B, F, P, x, par, y, ld, gy, gl, gx, gpar)

extern "C" void hc_affine(IO_ARGS, int inverse, int backward) {
    if (f64) affine<double>(IO(double), inverse, backward); else affine<float>(IO(float), inverse, backward);
}

extern "C" void hc_shift(IO_ARGS, const void* period, const void* lower, int inverse, int backward) {
    if (f64) shift<double>(IO(double), (const double*)period, (const double*)lower, inverse, backward);
    else shift<float>(IO(float), (const float*)period, (const float*)lower, inverse, backward);
}

extern "C" void hc_sos(IO_ARGS, int n_poly, int backward) {
    if (f64) sos<double>(IO(double), n_poly, backward); else sos<float>(IO(float), n_poly, backward);
}

extern "C" void hc_moebius(IO_ARGS, int d, double max_radius, int unit, int inverse, int backward) {
    if (f64) moebius<double>(IO(double), d, max_radius, unit, inverse, backward);
    else moebius<float>(IO(float), d, max_radius, unit, inverse, backward);
}

extern "C" void hc_spline(IO_ARGS, int K, int circular, int idslopes, int learn_lo, int learn_hi, const void* x0,
                          const void* xf, const void* y0, const void* yf, double min_bin, double min_slope, int inverse,
                          int backward, int* bins) {
    if (f64)
        spline<double, 8>(IO(double), K, circular, idslopes, learn_lo, learn_hi, (const double*)x0, (const double*)xf,
                          (const double*)y0, (const double*)yf, min_bin, min_slope, inverse, backward, bins);
    else
        spline<float, 8>(IO(float), K, circular, idslopes, learn_lo, learn_hi, (const float*)x0, (const float*)xf,
                         (const float*)y0, (const float*)yf, min_bin, min_slope, inverse, backward, bins);
}
