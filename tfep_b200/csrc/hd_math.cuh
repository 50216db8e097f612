// Scalar math shared by device kernels and the host-compiled math check (tests/hostcheck).
// Compiles as plain C++ (g++) and as CUDA (nvcc); no runtime dependencies.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TFEPB_HD __host__ __device__ __forceinline__
#else
#define TFEPB_HD inline
#endif

namespace tfepb {

template <typename T> struct Math;
template <> struct Math<float> {
    static TFEPB_HD float exp(float x) { return ::expf(x); }
    static TFEPB_HD float log(float x) { return ::logf(x); }
    static TFEPB_HD float log1p(float x) { return ::log1pf(x); }
    static TFEPB_HD float expm1(float x) { return ::expm1f(x); }
    static TFEPB_HD float sqrt(float x) { return ::sqrtf(x); }
    static TFEPB_HD float fmod(float x, float y) { return ::fmodf(x, y); }
    static TFEPB_HD float abs(float x) { return ::fabsf(x); }
};
template <> struct Math<double> {
    static TFEPB_HD double exp(double x) { return ::exp(x); }
    static TFEPB_HD double log(double x) { return ::log(x); }
    static TFEPB_HD double log1p(double x) { return ::log1p(x); }
    static TFEPB_HD double expm1(double x) { return ::expm1(x); }
    static TFEPB_HD double sqrt(double x) { return ::sqrt(x); }
    static TFEPB_HD double fmod(double x, double y) { return ::fmod(x, y); }
    static TFEPB_HD double abs(double x) { return ::fabs(x); }
};

// torch.remainder semantics (result takes the sign of the divisor): nn/transformers/spline.py:238,259.
template <typename T>
TFEPB_HD T py_remainder(T a, T b) {
    T m = Math<T>::fmod(a, b);
    if (m != T(0) && ((b < T(0)) != (m < T(0)))) m += b;
    return m;
}

// nn.ELU (alpha = 1): nn/conditioners/made.py:320.
template <typename T>
TFEPB_HD T elu(T x) { return x > T(0) ? x : Math<T>::expm1(x); }

// torch softplus (beta = 1, threshold = 20): nn/transformers/spline.py:415.
template <typename T>
TFEPB_HD T softplus(T x) { return x > T(20) ? x : Math<T>::log1p(Math<T>::exp(x)); }

template <typename T>
TFEPB_HD T sigmoid(T x) { return T(1) / (T(1) + Math<T>::exp(-x)); }

}  // namespace tfepb
