// ho_math.cuh -- register-resident complex arithmetic for the Berreman kernels (sm_100a).
//
// Everything here is per-thread scalar code on the FP64 (or FP32) pipe: a batch of tiny
// independent 4x4 complex problems has no GEMM shape, so no tensor cores / TMEM.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ho {

template <typename T> struct cx { T re, im; };

template <typename T> __device__ __forceinline__ cx<T> mk(T re, T im) { cx<T> z; z.re = re; z.im = im; return z; }
template <typename T> __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.re + b.re, a.im + b.im); }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.re - b.re, a.im - b.im); }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a) { return mk<T>(-a.re, -a.im); }
template <typename T> __device__ __forceinline__ cx<T> operator*(cx<T> a, cx<T> b) {
  return mk<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> __device__ __forceinline__ cx<T> operator*(T s, cx<T> a) { return mk<T>(s * a.re, s * a.im); }
template <typename T> __device__ __forceinline__ cx<T> operator*(cx<T> a, T s) { return mk<T>(s * a.re, s * a.im); }
template <typename T> __device__ __forceinline__ cx<T> operator+(cx<T> a, T s) { return mk<T>(a.re + s, a.im); }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a, T s) { return mk<T>(a.re - s, a.im); }
template <typename T> __device__ __forceinline__ cx<T> conj(cx<T> a) { return mk<T>(a.re, -a.im); }
template <typename T> __device__ __forceinline__ T norm2(cx<T> a) { return a.re * a.re + a.im * a.im; }
// a + b*c
template <typename T> __device__ __forceinline__ cx<T> fma(cx<T> b, cx<T> c, cx<T> a) {
  return mk<T>(a.re + b.re * c.re - b.im * c.im, a.im + b.re * c.im + b.im * c.re);
}
// a - b*c  (four FMAs: the negations are operand modifiers; `a - b * c` would be a product plus a
// subtraction, six instructions, because the compiler may not re-associate)
template <typename T> __device__ __forceinline__ cx<T> fnma(cx<T> b, cx<T> c, cx<T> a) {
  return mk<T>(a.re - b.re * c.re + b.im * c.im, a.im - b.re * c.im - b.im * c.re);
}
// a + s*c and a - s*c with a real s (two FMAs)
template <typename T> __device__ __forceinline__ cx<T> fma(T s, cx<T> c, cx<T> a) { return mk<T>(a.re + s * c.re, a.im + s * c.im); }
template <typename T> __device__ __forceinline__ cx<T> fnma(T s, cx<T> c, cx<T> a) { return mk<T>(a.re - s * c.re, a.im - s * c.im); }
// conj(a)*b
template <typename T> __device__ __forceinline__ cx<T> cmulc(cx<T> a, cx<T> b) {
  return mk<T>(a.re * b.re + a.im * b.im, a.re * b.im - a.im * b.re);
}
template <typename T> __device__ __forceinline__ cx<T> mul_i(cx<T> a) { return mk<T>(-a.im, a.re); }  // i*a
// Branch-free FP64 reciprocal and (sqrt, rsqrt): MUFU seed (rcp/rsqrt.approx.ftz.f64, ~2^-20) plus
// two Newton rounds on the FP64 pipe.  The library versions (1.0/x, sqrt, rsqrt) are IEEE-exact but
// carry a slow-path call that splits every use into separate basic blocks; these serial chains
// (~5 dependent DFMAs each, ~30 per point) are where a 4-warp scheduler starves, and without the
// branches the compiler interleaves neighbouring ones.  Accuracy ~1 ulp (parity bar is 1e-10).
// Special values: 0 and inf give NaN instead of inf/0 -- callers that can see an exact zero
// select around it; inf only occurs after a transfer-backend overflow, whose results are
// non-finite by contract.  Denormal inputs flush to zero.
__device__ __forceinline__ double ho_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = ::fma(-x, y, 1.0);
  y = ::fma(y, e, y);
  e = ::fma(-x, y, 1.0);
  return ::fma(y, e, y);
}
__device__ __forceinline__ float ho_rcp(float x) { return 1.f / x; }
// s = sqrt(x), rs = 1/sqrt(x) by the coupled (Goldschmidt) iteration
__device__ __forceinline__ void ho_sqrt_rsqrt(double x, double& s, double& rs) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = ::fma(-h, g, 0.5);
  g = ::fma(g, r, g); h = ::fma(h, r, h);
  r = ::fma(-h, g, 0.5);
  g = ::fma(g, r, g); h = ::fma(h, r, h);
  s = g; rs = h + h;
}
__device__ __forceinline__ void ho_sqrt_rsqrt(float x, float& s, float& rs) { rs = rsqrtf(x); s = x * rs; }

__device__ __forceinline__ cx<double> recip(cx<double> a) {
  double inv = ho_rcp(norm2(a));
  return mk<double>(a.re * inv, -a.im * inv);
}
// FP32: |a|^2 overflows already at |a| ~ 1.8e19 (the prism's 1/cos(theta) reaches 1.6e16 at the
// grazing end points), so scale by the larger component first.
__device__ __forceinline__ cx<float> recip(cx<float> a) {
  float s = fmaxf(fabsf(a.re), fabsf(a.im));
  if (!(s > 0.f)) return mk<float>(1.f / a.re, 0.f);
  float is = 1.f / s;
  float re = a.re * is, im = a.im * is;
  float inv = is / (re * re + im * im);
  return mk<float>(re * inv, -im * inv);
}
template <typename T> __device__ __forceinline__ cx<T> operator/(cx<T> a, cx<T> b) { return a * recip(b); }
template <typename T> __device__ __forceinline__ cx<T> sel(bool c, cx<T> a, cx<T> b) { return c ? a : b; }

__device__ __forceinline__ double ho_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float ho_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double ho_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float ho_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double ho_cbrt(double x) { return cbrt(x); }
__device__ __forceinline__ float ho_cbrt(float x) { return cbrtf(x); }
__device__ __forceinline__ double ho_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float ho_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double ho_exp(double x) { return exp(x); }
__device__ __forceinline__ float ho_exp(float x) { return expf(x); }
__device__ __forceinline__ void ho_sincos(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ void ho_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ double ho_abs(double x) { return fabs(x); }
__device__ __forceinline__ float ho_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double ho_copysign(double x, double y) { return copysign(x, y); }
__device__ __forceinline__ float ho_copysign(float x, float y) { return copysignf(x, y); }

template <typename T> __device__ __forceinline__ T cabs(cx<T> a) { return ho_sqrt(norm2(a)); }

// principal square root, branch-free: sqrt|z| and then (t, 1/t) of t^2 = (|z| + |re|)/2 from one
// coupled iteration each -- no division, no second square root.
template <typename T> __device__ __forceinline__ cx<T> csqrt(cx<T> z) {
  T n2 = norm2(z);
  T r, unused, t, it;
  ho_sqrt_rsqrt(n2, r, unused);
  T x = T(0.5) * (r + ho_abs(z.re));
  ho_sqrt_rsqrt(x, t, it);
  T u = (T(0.5) * z.im) * it;
  cx<T> w = (z.re >= T(0)) ? mk<T>(t, u) : mk<T>(ho_abs(u), ho_copysign(t, z.im));
  return (n2 > T(0)) ? w : mk<T>(T(0), T(0));
}

// principal cube root (float: library polar form)
__device__ __forceinline__ cx<float> ccbrt(cx<float> z) {
  float r = cabs(z);
  float th = atan2f(z.im, z.re) * (1.0f / 3.0f);
  float s, c;
  sincosf(th, &s, &c);
  float cr = cbrtf(r);
  return mk<float>(cr * c, cr * s);
}
// principal cube root (double): polar form evaluated in FP32 on the power-of-8 scaled argument
// (FP32/MUFU pipes, off the FP64 critical path), then two Newton rounds u <- (2u + z/u^2)/3 in
// FP64 (2e-7 -> 1e-13 -> rounding).  Branch-free; replaces atan2 + sincos + cbrt of the FP64
// library (115 FP64 instructions and three slow-path calls).  The Newton rounds stay on the
// principal branch because the seed is within 1e-6 of it.
__device__ __forceinline__ cx<double> ccbrt(cx<double> z) {
  const double big = fmax(fabs(z.re), fabs(z.im));
  // exponent of the larger component, rounded down to a multiple of 3
  const int ex = ((__double2hiint(big) >> 20) & 0x7ff) - 1023;
  const int e3 = (ex >= 0 ? ex : ex - 2) / 3;
  const double down = __hiloint2double((1023 - 3 * e3) << 20, 0);  // 2^(-3 e3)
  const double up = __hiloint2double((1023 + e3) << 20, 0);        // 2^(e3)
  const float fr = (float)(z.re * down), fi = (float)(z.im * down);  // |.| in [1, 8)
  const float r = sqrtf(fr * fr + fi * fi);
  const float th = atan2f(fi, fr) * (1.0f / 3.0f);
  float sn, cs;
  __sincosf(th, &sn, &cs);
  const float cr = cbrtf(r);
  cx<double> u = mk<double>((double)(cr * cs) * up, (double)(cr * sn) * up);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    cx<double> t = z * recip(u * u);
    u = mk<double>((2.0 * u.re + t.re) * (1.0 / 3.0), (2.0 * u.im + t.im) * (1.0 / 3.0));
  }
  return (big > 0.0) ? u : mk<double>(0.0, 0.0);
}

template <typename T> __device__ __forceinline__ cx<T> cexp(cx<T> z) {
  T e = ho_exp(z.re);
  T s, c;
  ho_sincos(z.im, &s, &c);
  return mk<T>(e * c, e * s);
}

// (exp(x) - 1) / x by its series, for |x| < 0.1: nine terms leave 0.1^9 / 10! = 3e-16 (above that
// the callers use the two exponentials directly, whose cancellation error eps / |x| is <= 1e-15)
template <typename T> __device__ __forceinline__ cx<T> expm1c_small(cx<T> x) {
  cx<T> acc = mk<T>(T(1), T(0));
#pragma unroll
  for (int n = 10; n >= 2; --n) acc = fma(x * (T(1) / T(n)), acc, mk<T>(T(1), T(0)));
  return acc;
}

}  // namespace ho
