// ho_berreman.cuh -- per-point physics of the fused reflection kernel.
//
// One thread owns one batch point and keeps everything in registers:
//   rotate eps/mu about z  ->  structured Berreman matrix  ->  characteristic quartic
//   ->  forward/backward spectral split  ->  spectral projector  ->  running 4x2 "plane"
// swept right-to-left (exit -> prism), then the 2x2 reflection block and the Mueller matrix.
// The algorithm never forms eigenvectors; see DESIGN.md section 4 and tools/algo_model.py
// (the NumPy statement-for-statement model this file was written against).
#pragma once
#include "ho_math.cuh"

namespace ho {

// Berreman matrix for symmetric eps, mu (reference waves.py:221-245):
//   [[a, b, c, d], [0, e, f, -c], [g, h, e, -b], [j, -g, 0, a]]
// NM = true: mu is a scalar multiple of the identity, so c = e = 0 (and f = -mu); the members
// stay in the struct but are never read, which lets the compiler drop them.
template <typename T, bool NM = false> struct Delta { cx<T> a, b, c, d, e, f, g, h, j; };
template <typename T> struct Vec4 { cx<T> v0, v1, v2, v3; };
template <typename T> struct Sym6 { cx<T> xx, xy, xz, yy, yz, zz; };
// sums / products of the forward and backward root pairs
template <typename T> struct Spectrum { cx<T> sf, pf, sb, pb; };
// z rows of the rotated tensors, for the longitudinal fields of a mode (reference waves.py:382-408):
// Ez = -(k Hy + e20 Ex + e21 Ey) / e22,  Hz = (k Ey - m20 Hx - m21 Hy) / m22
template <typename T> struct ZRow { cx<T> e20, e21, ie22, m20, m21, im22; };

template <typename T> __device__ __forceinline__ Vec4<T> axpy(cx<T> s, const Vec4<T>& x, const Vec4<T>& y) {
  Vec4<T> r; r.v0 = fma(s, x.v0, y.v0); r.v1 = fma(s, x.v1, y.v1); r.v2 = fma(s, x.v2, y.v2); r.v3 = fma(s, x.v3, y.v3); return r;
}
template <typename T> __device__ __forceinline__ Vec4<T> scale(cx<T> s, const Vec4<T>& x) {
  Vec4<T> r; r.v0 = s * x.v0; r.v1 = s * x.v1; r.v2 = s * x.v2; r.v3 = s * x.v3; return r;
}
template <typename T> __device__ __forceinline__ Vec4<T> sub(const Vec4<T>& x, const Vec4<T>& y) {
  Vec4<T> r; r.v0 = x.v0 - y.v0; r.v1 = x.v1 - y.v1; r.v2 = x.v2 - y.v2; r.v3 = x.v3 - y.v3; return r;
}
template <typename T> __device__ __forceinline__ Vec4<T> add(const Vec4<T>& x, const Vec4<T>& y) {
  Vec4<T> r; r.v0 = x.v0 + y.v0; r.v1 = x.v1 + y.v1; r.v2 = x.v2 + y.v2; r.v3 = x.v3 + y.v3; return r;
}
template <typename T> __device__ __forceinline__ T norm2(const Vec4<T>& x) {
  return norm2(x.v0) + norm2(x.v1) + norm2(x.v2) + norm2(x.v3);
}
template <typename T> __device__ __forceinline__ cx<T> dotc(const Vec4<T>& x, const Vec4<T>& y) {  // x^H y
  return cmulc(x.v0, y.v0) + cmulc(x.v1, y.v1) + cmulc(x.v2, y.v2) + cmulc(x.v3, y.v3);
}
template <typename T> __device__ __forceinline__ Vec4<T> sel(bool c, const Vec4<T>& a, const Vec4<T>& b) {
  Vec4<T> r; r.v0 = sel(c, a.v0, b.v0); r.v1 = sel(c, a.v1, b.v1); r.v2 = sel(c, a.v2, b.v2); r.v3 = sel(c, a.v3, b.v3); return r;
}

// Rz(beta) S Rz(beta)^T for a packed symmetric tensor (reference layers.py:55-66, 411-417)
template <typename T> __device__ __forceinline__ Sym6<T> rotate_z(const Sym6<T>& s, T cb, T sb) {
  T c2 = cb * cb, s2 = sb * sb, cs = cb * sb;
  Sym6<T> r;
  r.xx = fma(s2, s.yy, fnma(T(2) * cs, s.xy, c2 * s.xx));
  r.xy = fma(c2 - s2, s.xy, cs * (s.xx - s.yy));
  r.xz = fnma(sb, s.yz, cb * s.xz);
  r.yy = fma(c2, s.yy, fma(T(2) * cs, s.xy, s2 * s.xx));
  r.yz = fma(cb, s.yz, sb * s.xz);
  r.zz = s.zz;
  return r;
}

// scalar-mu flavour: m = mu * I
template <typename T> __device__ __forceinline__ Delta<T, true> berreman_nm(T k, const Sym6<T>& e, cx<T> mu, ZRow<T>& z) {
  cx<T> ie = recip(e.zz), im = recip(mu);
  z.e20 = e.xz; z.e21 = e.yz; z.ie22 = ie; z.m20 = mk<T>(T(0), T(0)); z.m21 = z.m20; z.im22 = im;
  T k2 = k * k;
  Delta<T, true> D;
  D.a = -(k * (e.xz * ie));
  D.b = -(k * (e.yz * ie));
  D.d = fnma(k2, ie, mu);
  D.f = -mu;
  D.g = fma(e.yz * e.xz, ie, -e.xy);
  D.h = fma(e.yz * e.yz, ie, fnma(T(1), e.yy, k2 * im));
  D.j = fnma(e.xz * e.xz, ie, e.xx);
  D.c = mk<T>(T(0), T(0));
  D.e = mk<T>(T(0), T(0));
  return D;
}

template <typename T> __device__ __forceinline__ Delta<T> berreman(T k, const Sym6<T>& e, const Sym6<T>& m, ZRow<T>& z) {
  cx<T> ie = recip(e.zz), im = recip(m.zz);
  z.e20 = e.xz; z.e21 = e.yz; z.ie22 = ie; z.m20 = m.xz; z.m21 = m.yz; z.im22 = im;
  T k2 = k * k;
  Delta<T> D;
  D.a = -(k * (e.xz * ie));
  D.b = k * (m.yz * im - e.yz * ie);
  D.c = m.xy - m.yz * m.xz * im;
  D.d = m.yy - m.yz * m.yz * im - k2 * ie;
  D.e = -(k * (m.xz * im));
  D.f = m.xz * m.xz * im - m.xx;
  D.g = e.yz * e.xz * ie - e.xy;
  D.h = k2 * im - e.yy + e.yz * e.yz * ie;
  D.j = e.xx - e.xz * e.xz * ie;
  return D;
}

template <typename T> __device__ __forceinline__ Vec4<T> matvec(const Delta<T, true>& D, const Vec4<T>& v) {
  Vec4<T> r;
  r.v0 = fma(D.d, v.v3, fma(D.b, v.v1, D.a * v.v0));
  r.v1 = D.f * v.v2;
  r.v2 = fnma(D.b, v.v3, fma(D.h, v.v1, D.g * v.v0));
  r.v3 = fnma(D.g, v.v1, fma(D.a, v.v3, D.j * v.v0));
  return r;
}

template <typename T> __device__ __forceinline__ Vec4<T> matvec(const Delta<T>& D, const Vec4<T>& v) {
  Vec4<T> r;
  r.v0 = fma(D.d, v.v3, fma(D.c, v.v2, fma(D.b, v.v1, D.a * v.v0)));
  r.v1 = fnma(D.c, v.v3, fma(D.f, v.v2, D.e * v.v1));
  r.v2 = fnma(D.b, v.v3, fma(D.e, v.v2, fma(D.h, v.v1, D.g * v.v0)));
  r.v3 = fnma(D.g, v.v1, fma(D.a, v.v3, D.j * v.v0));
  return r;
}

// det(lambda I - D) = lambda^4 + c3 lambda^3 + c2 lambda^2 + c1 lambda + c0, in two halves: a uniaxial
// film only needs c3 and c2 (its quartic factorises), everything else all four
template <typename T> __device__ __forceinline__ void charpoly_hi(const Delta<T, true>& D, cx<T>& c3, cx<T>& c2) {
  c3 = T(-2) * D.a;
  c2 = fnma(D.f, D.h, fnma(D.d, D.j, D.a * D.a));
}
template <typename T> __device__ __forceinline__ void charpoly_lo(const Delta<T, true>& D, cx<T>& c1, cx<T>& c0) {
  cx<T> a2 = D.a * D.a, ab = D.a * D.b;
  c1 = T(2) * (D.f * fnma(D.b, D.g, D.a * D.h));
  c0 = D.f * fma(D.d, fma(D.h, D.j, D.g * D.g), fma(D.b * D.b, D.j, fnma(a2, D.h, T(2) * (ab * D.g))));
}
template <typename T> __device__ __forceinline__ void charpoly_hi(const Delta<T>& D, cx<T>& c3, cx<T>& c2) {
  cx<T> a2 = D.a * D.a, e2 = D.e * D.e, cg = D.c * D.g, dj = D.d * D.j, fh = D.f * D.h, ae = D.a * D.e;
  c3 = T(-2) * (D.a + D.e);
  c2 = a2 + T(4) * ae - T(2) * cg - dj + e2 - fh;
}
template <typename T> __device__ __forceinline__ void charpoly_lo(const Delta<T>& D, cx<T>& c1, cx<T>& c0) {
  cx<T> a2 = D.a * D.a, e2 = D.e * D.e, cg = D.c * D.g, dj = D.d * D.j, fh = D.f * D.h;
  cx<T> bcj = D.b * D.c * D.j, bf = D.b * D.f, ae = D.a * D.e;
  c1 = T(-2) * (a2 * D.e - D.a * cg + D.a * e2 - D.a * fh - bcj + bf * D.g - D.e * cg - D.e * dj);
  c0 = a2 * e2 - a2 * fh + T(2) * (D.a * bf * D.g) - T(2) * (ae * cg) + D.b * bf * D.j - T(2) * (D.e * bcj)
       + cg * cg + D.c * D.c * D.h * D.j - e2 * dj + D.d * D.f * D.g * D.g + dj * fh;
}
template <typename T, bool NM> __device__ __forceinline__ void charpoly(const Delta<T, NM>& D, cx<T>& c3, cx<T>& c2, cx<T>& c1, cx<T>& c0) {
  charpoly_hi(D, c3, c2);
  charpoly_lo(D, c1, c0);
}

// Roots of y^2 + bq y + cq without cancellation.  A pair closer than ~sqrt(tau) relative to |bq|
// is collapsed onto its mean: the split of a near-double root is rounding noise (~sqrt(eps)) and
// could otherwise put its two copies on opposite sides of the forward/backward classification
// when the loss is tiny (seen in FP32 on SiC at 1700 cm^-1).
template <typename T> __device__ __forceinline__ void quad_roots(cx<T> bq, cx<T> cq, cx<T>& y0, cx<T>& y1) {
  const T tau = sizeof(T) == 8 ? T(1e-12) : T(1e-5);
  cx<T> sq = csqrt(fma(bq, bq, T(-4) * cq));
  if (norm2(sq) <= tau * norm2(bq)) sq = mk<T>(T(0), T(0));
  bool plus = cmulc(bq, sq).re >= T(0);
  cx<T> t = T(-0.5) * (plus ? bq + sq : bq - sq);
  y0 = t;
  y1 = (norm2(t) > T(0)) ? cq / t : mk<T>(T(0), T(0));
}

// Ferrari / Cardano closed form: start values for the spectral split
template <typename T> __device__ __forceinline__ void ferrari(cx<T> c3, cx<T> c2, cx<T> c1, cx<T> c0, cx<T> lam[4]) {
  cx<T> q3 = T(0.25) * c3;
  cx<T> q32 = q3 * q3;
  cx<T> p = c2 - T(6) * q32;
  cx<T> q = c1 - T(2) * (q3 * c2) + T(8) * (q32 * q3);
  cx<T> r = c0 - q3 * c1 + q32 * c2 - T(3) * (q32 * q32);
  cx<T> p2 = p * p;
  cx<T> P = T(-1.0 / 12.0) * p2 - r;
  cx<T> Q = T(-1.0 / 108.0) * (p2 * p) + T(1.0 / 3.0) * (p * r) - T(0.125) * (q * q);
  cx<T> sq = csqrt(T(0.25) * (Q * Q) + T(1.0 / 27.0) * (P * P * P));
  cx<T> ua = T(-0.5) * Q + sq, ub = T(-0.5) * Q - sq;
  cx<T> u = ccbrt(norm2(ua) >= norm2(ub) ? ua : ub);
  const cx<T> w1 = mk<T>(T(-0.5), T(0.86602540378443864676));
  const cx<T> w2 = mk<T>(T(-0.5), T(-0.86602540378443864676));
  cx<T> third_p = T(1.0 / 3.0) * p;
  cx<T> m = mk<T>(T(0), T(0));
  T best = T(-1);
  // P / (3 u w^k) = (P / 3u) conj(w^k): one reciprocal serves the three cube roots
  const bool u_ok = norm2(u) > T(0);
  cx<T> pu = u_ok ? T(1.0 / 3.0) * (P * recip(u)) : mk<T>(T(0), T(0));
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    cx<T> uk = (kk == 0) ? u : (kk == 1 ? u * w1 : u * w2);
    cx<T> tk = uk - ((kk == 0) ? pu : (kk == 1 ? pu * w2 : pu * w1));
    cx<T> mk_ = tk - third_p;
    T ak = norm2(mk_);
    if (ak > best) { best = ak; m = mk_; }
  }
  cx<T> alpha = csqrt(T(2) * m);
  cx<T> qa = (norm2(alpha) > T(0)) ? T(0.5) * (q / alpha) : mk<T>(T(0), T(0));
  cx<T> base = T(0.5) * p + m;
  cx<T> y0, y1, y2, y3;
  quad_roots(alpha, base - qa, y0, y1);
  quad_roots(-alpha, base + qa, y2, y3);
  lam[0] = y0 - q3; lam[1] = y1 - q3; lam[2] = y2 - q3; lam[3] = y3 - q3;
}

// Start values for the spectral split: always the FP64 closed form, also for the complex64
// kernel.  An FP32 Ferrari was tried (it would run on the FP32 pipe next to the FP64 work) and
// rejected: when both root pairs are nearly double (near-normal incidence on a c-cut uniaxial
// film) the resolvent discriminant cancels catastrophically in FP32 and the start roots come out
// ~8 % off with the wrong sign of Im -> wrong forward/backward classification.
__device__ __forceinline__ void start_roots(cx<double> c3, cx<double> c2, cx<double> c1, cx<double> c0, cx<double> lam[4]) {
  ferrari(c3, c2, c1, c0, lam);
}
__device__ __forceinline__ void start_roots(cx<float> c3, cx<float> c2, cx<float> c1, cx<float> c0, cx<float> lam[4]) {
  cx<double> l64[4];
  start_roots(mk<double>(c3.re, c3.im), mk<double>(c2.re, c2.im), mk<double>(c1.re, c1.im), mk<double>(c0.re, c0.im), l64);
#pragma unroll
  for (int i = 0; i < 4; ++i) lam[i] = mk<float>((float)l64[i].re, (float)l64[i].im);
}

// Forward ("transmitted") pair selection, reference waves.py:273-289: all |Im| > 1e-9 -> two
// largest Im; none -> two largest Re; mixed -> Im > tol, or |Im| <= tol and Re > 0 (ranked).
template <typename T> __device__ __forceinline__ void classify(const cx<T> lam[4], bool fwd[4]) {
  const T tol = T(1e-9);
  int n_c = 0;
  bool cplx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { cplx[i] = ho_abs(lam[i].im) > tol; n_c += cplx[i] ? 1 : 0; }
  int k1[4];
  T k2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int score = cplx[i] ? (lam[i].im > T(0) ? 2 : -2) : (lam[i].re > T(0) ? 1 : -1);
    k1[i] = (n_c == 4 || n_c == 0) ? 0 : score;
    k2[i] = (n_c == 0) ? lam[i].re : lam[i].im;
  }
  // rank = number of roots that sort before this one (lexicographic k1, k2, Re; index breaks
  // ties, so the order is strict and each unordered pair needs one evaluation)
  int rank[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int jx = i + 1; jx < 4; ++jx) {
      // does jx sort before i?  (on a full tie the lower index i wins)
      bool j_first = (k1[jx] > k1[i]) ||
                     (k1[jx] == k1[i] && ((k2[jx] > k2[i]) || (k2[jx] == k2[i] && lam[jx].re > lam[i].re)));
      rank[i] += j_first ? 1 : 0;
      rank[jx] += j_first ? 0 : 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) fwd[i] = rank[i] < 2;
}

// Bairstow: Newton on (u, v) so that lambda^2 + u lambda + v divides the quartic; the
// complementary factor lambda^2 + b1 lambda + b0 follows by deflation.  Per-point early exits:
// the Newton step is skipped once the residual of the division is at rounding level, and the
// loop ends once a step is negligible (accurate start values need exactly one step).
template <typename T> __device__ __forceinline__ bool bairstow(cx<T> c3, cx<T> c2, cx<T> c1, cx<T> c0, cx<T>& u, cx<T>& v, cx<T>& b1, cx<T>& b0) {
  const T res_tol2 = sizeof(T) == 8 ? T(1e-28) : T(1e-12);
  const T step_tol2 = sizeof(T) == 8 ? T(1e-22) : T(1e-7);
  const int max_steps = 3;
  bool converged = false;
#pragma unroll 1
  for (int it = 0;; ++it) {
    b1 = c3 - u;
    b0 = fnma(u, b1, c2 - v);
    cx<T> t1 = u * b0, t2 = v * b1, t3 = v * b0;
    cx<T> r1 = c1 - t1 - t2;
    cx<T> r0 = c0 - t3;
    if (it > 0 && norm2(r1) <= res_tol2 * (norm2(t1) + norm2(t2)) && norm2(r0) <= res_tol2 * norm2(t3)) { converged = true; break; }
    if (it == max_steps) break;
    cx<T> umb = u - b1;
    cx<T> j22 = v - b0;
    cx<T> j11 = fnma(u, umb, j22);
    cx<T> j21 = -(v * umb);
    cx<T> det = fnma(umb, j21, j11 * j22);
    if (!(norm2(det) > T(0))) break;
    cx<T> idet = recip(det);
    cx<T> du = fnma(r1, j22, r0 * umb) * idet;
    cx<T> dv = fnma(r0, j11, r1 * j21) * idet;
    u = u + du;
    v = v + dv;
    if (norm2(du) + norm2(dv) <= step_tol2 * (norm2(u) + norm2(v))) {
      converged = true;
      b1 = c3 - u;
      b0 = fnma(u, b1, c2 - v);
      break;
    }
  }
  return converged;
}

// square root of y oriented forward: the reference's mode rule (waves.py:273-289) on the pair +-mu
template <typename T> __device__ __forceinline__ cx<T> forward_sqrt(cx<T> y) {
  cx<T> m = csqrt(y);
  bool fwd = (ho_abs(m.im) > T(1e-9)) ? (m.im > T(0)) : (m.re > T(0));
  return fwd ? m : -m;
}

// charpoly -> start values -> Bairstow polish of the forward quadratic factor.
// Start values: when the odd coefficients are negligible (z is a principal axis of the rotated
// tensors up to the reference's hidden 1e-8 angle offsets -- every un-tilted crystal) the quartic
// is a perturbed biquadratic and the forward pair comes from three square roots; otherwise
// Ferrari + classification.  A biquadratic start that fails to converge falls back to Ferrari.
// FAST == 1 (fast-path kernel): only the biquadratic start is compiled; a point that would need
// Ferrari sets `bail` and is finished later by the general kernel.  FAST == 2 (lean kernel for
// tilted stacks): the general start values, as FAST == 0.
// (Uniaxial crystals do not come here: split_uniaxial.)
template <typename T, int FAST = 0> __device__ __forceinline__ Spectrum<T> split_spectrum(cx<T> c3, cx<T> c2, cx<T> c1, cx<T> c0, bool& bail) {
  const T tau4 = T(1e-20);
  T n2 = norm2(c2), n3 = norm2(c3), n1 = norm2(c1);
  bool biq = (n3 * n3 <= tau4 * n2) && (n1 * n1 <= tau4 * (n2 * n2 * n2));
  cx<T> u, v, b1, b0;
  if constexpr (FAST == 1) {
    cx<T> y0, y1;
    quad_roots(c2, c0, y0, y1);
    cx<T> m0 = forward_sqrt(y0), m1 = forward_sqrt(y1);
    u = -(m0 + m1);
    v = m0 * m1;
    bool ok = bairstow(c3, c2, c1, c0, u, v, b1, b0);
    bail = bail || !biq || !ok;
  } else {
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (biq) {
        cx<T> y0, y1;
        quad_roots(c2, c0, y0, y1);
        cx<T> m0 = forward_sqrt(y0), m1 = forward_sqrt(y1);
        u = -(m0 + m1);
        v = m0 * m1;
      } else {
        cx<T> lam[4];
        start_roots(c3, c2, c1, c0, lam);
        bool fwd[4];
        classify(lam, fwd);
        cx<T> sf = mk<T>(T(0), T(0)), pf = mk<T>(T(1), T(0));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (fwd[i]) { sf = sf + lam[i]; pf = pf * lam[i]; }
        }
        u = -sf;
        v = pf;
      }
      bool ok = bairstow(c3, c2, c1, c0, u, v, b1, b0);
      if (ok || !biq) break;
      biq = false;
    }
  }
  Spectrum<T> s;
  s.sf = -u; s.pf = v; s.sb = -b1; s.pb = b0;
  return s;
}

// Spectral split of a uniaxial crystal (any orientation; q_o = kx^2 - mu eps_o from
// ho_layer_t.eps_ordinary): the quartic factorises exactly into the ordinary and the extraordinary
// quadratic, (x^2 + q_o)(x^2 + c3 x + q_e) with q_e = c2 - q_o, so the four roots are two square
// roots away -- no Ferrari, no polish, and c1, c0 are not needed at all (the host vouches for the
// hint: plan._ordinary_value).  Forward pair by the reference's rule (waves.py:273-289) on the four
// roots; the fast build uses the rule per root and reports a root pair that is not one forward +
// one backward (`false`: the caller falls back to the general split).
template <typename T, int FAST = 0> __device__ __forceinline__ bool split_uniaxial(cx<T> c3, cx<T> c2, cx<T> qo, Spectrum<T>& s) {
  cx<T> lam[4];
  lam[0] = csqrt(-qo);
  lam[1] = -lam[0];
  quad_roots(c3, c2 - qo, lam[2], lam[3]);
  bool fwd[4];
  if constexpr (FAST == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) fwd[i] = (ho_abs(lam[i].im) > T(1e-9)) ? (lam[i].im > T(0)) : (lam[i].re > T(0));
    if (fwd[0] == fwd[1] || fwd[2] == fwd[3]) return false;
  } else {
    classify(lam, fwd);
  }
  const cx<T> one = mk<T>(T(1), T(0)), zero = mk<T>(T(0), T(0));
  s.sf = zero; s.pf = one; s.sb = zero; s.pb = one;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (fwd[i]) { s.sf = s.sf + lam[i]; s.pf = s.pf * lam[i]; }
    else        { s.sb = s.sb + lam[i]; s.pb = s.pb * lam[i]; }
  }
  return true;
}

// h(x) = h0 + h1 x with h q_b = 1 (mod q_f): the forward spectral projector is q_b(D) h(D)
template <typename T> __device__ __forceinline__ void projector_coeffs(const Spectrum<T>& s, cx<T>& h0, cx<T>& h1) {
  cx<T> U = s.sf - s.sb, V = s.pb - s.pf;
  cx<T> R = fma(V, V, fma(U * V, s.sf, (U * U) * s.pf));
  cx<T> iR = recip(R);
  h0 = fma(U, s.sf, V) * iR;
  h1 = -(U * iR);
}

// z = P_f y, also returns dy = D y
template <typename T, bool NM> __device__ __forceinline__ Vec4<T> project_forward(const Delta<T, NM>& D, const Spectrum<T>& s, cx<T> h0, cx<T> h1,
                                                                        const Vec4<T>& y, Vec4<T>& dy) {
  dy = matvec(D, y);
  Vec4<T> w = axpy(h1, dy, scale(h0, y));
  Vec4<T> t1 = matvec(D, w);
  Vec4<T> t2 = matvec(D, t1);
  return axpy(s.pb, w, axpy(-s.sb, t1, t2));
}

// Root pair (sum s, product p) ordered by |exp(c lambda)|: exp(c K) = phi_s I + dd (K - lam_s I)
// growth = max |Re(c lambda)| of the two modes
template <typename T> __device__ __forceinline__ void pair_modes(cx<T> s, cx<T> p, cx<T> c, cx<T>& lam_s, cx<T>& phi_s, cx<T>& dd, bool& separated, T& growth) {
  cx<T> m = T(0.5) * s;
  cx<T> h = csqrt(fma(m, m, -p));
  if ((c * h).re < T(0)) h = -h;
  lam_s = m - h;
  cx<T> lam_g = m + h;
  cx<T> cs = c * lam_s;
  growth = fmax(ho_abs(cs.re), ho_abs((c * lam_g).re));
  phi_s = cexp(cs);
  cx<T> x = T(2) * (c * h);
  separated = x.re > T(1);
  if (norm2(x) < T(0.01)) {
    dd = c * phi_s * expm1c_small(x);
  } else {
    dd = (cexp(c * lam_g) - phi_s) / (lam_g - lam_s);
  }
}

// Project the (theoretically rank-one) 4x2 matrix [r0 r1] onto rank one.
template <typename T> __device__ __forceinline__ void rank1_clean(Vec4<T>& r0, Vec4<T>& r1) {
  T n0 = norm2(r0), n1 = norm2(r1);
  bool first = n0 >= n1;
  Vec4<T> v = sel(first, r0, r1), w = sel(first, r1, r0);
  T nv = first ? n0 : n1;
  cx<T> beta = mk<T>(T(0), T(0));
  if (nv > T(0)) beta = ho_rcp(nv) * dotc(v, w);
  Vec4<T> wc = scale(beta, v);
  r0 = sel(first, v, wc);
  r1 = sel(first, wc, v);
}

// rank-one structure of [r0 r1]: the larger column v and beta with  other column = beta v
template <typename T> __device__ __forceinline__ void rank1_coeff(const Vec4<T>& r0, const Vec4<T>& r1, bool& first, cx<T>& beta) {
  T n0 = norm2(r0), n1 = norm2(r1);
  first = n0 >= n1;
  cx<T> d = dotc(r0, r1);      // r0^H r1
  if (!first) d = conj(d);     // r1^H r0
  T nv = first ? n0 : n1;
  beta = (nv > T(0)) ? ho_rcp(nv) * d : mk<T>(T(0), T(0));
}

// exp(c D) applied to two columns lying in one invariant subspace, given the pair's modes
template <typename T> __device__ __forceinline__ void block_exp_apply(cx<T> lam_s, cx<T> phi_s, cx<T> dd, bool separated,
                                                                     const Vec4<T>& z0, const Vec4<T>& z1,
                                                                     const Vec4<T>& dz0, const Vec4<T>& dz1, Vec4<T>& o0, Vec4<T>& o1) {
  Vec4<T> r0 = axpy(-lam_s, z0, dz0), r1 = axpy(-lam_s, z1, dz1);
  if (separated) rank1_clean(r0, r1);
  o0 = axpy(dd, r0, scale(phi_s, z0));
  o1 = axpy(dd, r1, scale(phi_s, z1));
}

// cosh(z) and sinh(z)/z as functions of w = z^2: both are entire in w, so neither a complex square
// root nor a branch choice is needed.  Below |w| = 1 they are evaluated by their Taylor series in w
// (Horner with real coefficients 1/(2k)!, 1/(2k+1)!: 9 steps each, truncation < 4e-19, no
// cancellation in e^z - e^-z, no 0/0) -- cheaper than one complex exponential plus two reciprocals,
// and for an optically thin film (|c mu| = k0 d |mu| << 1: BASELINE C4's 0.1 um hBN) it is the path
// nearly every point takes, so warps do not diverge on the switch.
// 1/(2k)! and 1/(2k+1)!, k = 9 .. 0 (literal arguments: the unrolled callers fold them to immediates)
__host__ __device__ constexpr double cosh_coef(int i) {
  return i == 0 ? 1.0 / 6402373705728000.0 : i == 1 ? 1.0 / 20922789888000.0 : i == 2 ? 1.0 / 87178291200.0 : i == 3 ? 1.0 / 479001600.0
       : i == 4 ? 1.0 / 3628800.0 : i == 5 ? 1.0 / 40320.0 : i == 6 ? 1.0 / 720.0 : i == 7 ? 1.0 / 24.0 : i == 8 ? 0.5 : 1.0;
}
__host__ __device__ constexpr double sinhc_coef(int i) {
  return i == 0 ? 1.0 / 121645100408832000.0 : i == 1 ? 1.0 / 355687428096000.0 : i == 2 ? 1.0 / 1307674368000.0 : i == 3 ? 1.0 / 6227020800.0
       : i == 4 ? 1.0 / 39916800.0 : i == 5 ? 1.0 / 362880.0 : i == 6 ? 1.0 / 5040.0 : i == 7 ? 1.0 / 120.0 : i == 8 ? 1.0 / 6.0 : 1.0;
}
constexpr int kCoshTerms = 10;
// one Horner step with a real constant term: w * a + k
template <typename T> __device__ __forceinline__ cx<T> horner_step(cx<T> w, cx<T> a, double k) {
  return mk<T>(::fma(w.re, a.re, ::fma(-w.im, a.im, T(k))), ::fma(w.re, a.im, w.im * a.re));
}
// first step: the leading coefficient is real
template <typename T> __device__ __forceinline__ cx<T> horner_first(cx<T> w, double k1, double k0) {
  return mk<T>(::fma(T(k1), w.re, T(k0)), T(k1) * w.im);
}
template <typename T> __device__ __forceinline__ void cosh_sinhc_series(cx<T> w, cx<T>& ch, cx<T>& shc) {
  ch = horner_first(w, cosh_coef(0), cosh_coef(1));
  shc = horner_first(w, sinhc_coef(0), sinhc_coef(1));
#pragma unroll
  for (int i = 2; i < kCoshTerms; ++i) { ch = horner_step(w, ch, cosh_coef(i)); shc = horner_step(w, shc, sinhc_coef(i)); }
}
// the same for two arguments at once: four independent Horner chains for the FP64 pipe to interleave.
// FIRST: index of the leading coefficient -- 0: all ten terms (|w| < 1); 3: seven terms, enough below
// |w| = 0.1 (truncation 0.1^7 / 14! = 1e-18), where most points of an optically thin film are.
template <typename T, int FIRST> __device__ __forceinline__ void cosh_sinhc_series2(cx<T> w0, cx<T> w1, cx<T>& ch0, cx<T>& shc0, cx<T>& ch1, cx<T>& shc1) {
  ch0 = horner_first(w0, cosh_coef(FIRST), cosh_coef(FIRST + 1));   ch1 = horner_first(w1, cosh_coef(FIRST), cosh_coef(FIRST + 1));
  shc0 = horner_first(w0, sinhc_coef(FIRST), sinhc_coef(FIRST + 1)); shc1 = horner_first(w1, sinhc_coef(FIRST), sinhc_coef(FIRST + 1));
#pragma unroll
  for (int i = FIRST + 2; i < kCoshTerms; ++i) {
    ch0 = horner_step(w0, ch0, cosh_coef(i)); ch1 = horner_step(w1, ch1, cosh_coef(i));
    shc0 = horner_step(w0, shc0, sinhc_coef(i)); shc1 = horner_step(w1, shc1, sinhc_coef(i));
  }
}
template <typename T> __device__ __forceinline__ void cosh_sinhc_w(cx<T> w, cx<T>& ch, cx<T>& shc) {
  if (norm2(w) < T(1)) {
    cosh_sinhc_series(w, ch, shc);
  } else {
    cx<T> z = csqrt(w);
    cx<T> e = cexp(z), ei = recip(e);
    ch = T(0.5) * (e + ei);
    shc = (T(0.5) * (e - ei)) * recip(z);
  }
}

// both root pairs of a film: four interleaved series when both arguments are small (the usual case)
template <typename T> __device__ __forceinline__ void cosh_sinhc_pair(cx<T> w0, cx<T> w1, cx<T>& A0, cx<T>& S0, cx<T>& A1, cx<T>& S1) {
  const T n0 = norm2(w0), n1 = norm2(w1);
  if (sizeof(T) == 8 && n0 < T(0.01) && n1 < T(0.01)) {
    cosh_sinhc_series2<T, 3>(w0, w1, A0, S0, A1, S1);
  } else if (n0 < T(1) && n1 < T(1)) {
    cosh_sinhc_series2<T, 0>(w0, w1, A0, S0, A1, S1);
  } else {   // rare (|c mu| >= 1): one rolled instance of the general evaluation keeps the code small
    cx<T> ww[2] = {w0, w1}, AA[2], SS[2];
#pragma unroll 1
    for (int i = 0; i < 2; ++i) cosh_sinhc_w(ww[i], AA[i], SS[i]);
    A0 = AA[0]; S0 = SS[0]; A1 = AA[1]; S1 = SS[1];
  }
}

// Thin-film switch of both recursions: every mode of the pair m +- h has |Re(c lambda)| below the
// limit.  The cubic mixes the four exponentials and costs log10 of their ratio in digits: e^6 (2.6
// of 16 digits) in complex128, e^3 (1.3 of 7) in the opt-in complex64 kernels.
// Given Re(c m) and w = (c h)^2 the test |Re(c m)| + |Re sqrt(w)| < limit is evaluated without the
// square root: |Re sqrt(w)| = sqrt((|w| + Re w) / 2) < L  <=>  |w| < 2 L^2 - Re w.
template <typename T> __device__ __forceinline__ bool thin_pair(T re_cm, cx<T> w) {
  const T limit = sizeof(T) == 8 ? T(3) : T(1.5);
  const T lim = limit - ho_abs(re_cm);
  const T r = ::fma(T(2) * lim, lim, -w.re);
  return lim > T(0) && r > T(0) && norm2(w) < r * r;
}
// Linear interpolants of exp(c x) on the two root pairs m +- h of a thin film, centred form
// alpha(x) = ph + dd (x - m) with ph = e^{cm} cosh(ch), dd = e^{cm} c sinh(ch)/(ch): entire and even in
// h -- no ordering of the two modes, no square root, one exponential per pair (plus a series while
// |ch| < 1) instead of two.  cm = c m, w = (c h)^2 = c^2 (m^2 - p).
// SCALE_FREE (no power / transmission outputs): the running plane is only defined up to a common
// scalar, so the factor e^{c m_b} is dropped from both interpolants and ONE complex exponential,
// e^{c (m_f - m_b)}, is left.
template <typename T, bool SCALE_FREE> __device__ __forceinline__ void thin_pairs(cx<T> cmf, cx<T> wf, cx<T> cmb, cx<T> wb, cx<T> c,
                                                                            cx<T>& phf, cx<T>& ddf, cx<T>& phb, cx<T>& ddb) {
  cx<T> Af, Sf, Ab, Sb;
  cosh_sinhc_pair(wf, wb, Af, Sf, Ab, Sb);
  if constexpr (SCALE_FREE) {
    cx<T> e = cexp(cmf - cmb);
    phf = e * Af; ddf = e * (c * Sf);
    phb = Ab;     ddb = c * Sb;
  } else {
    cx<T> ef = cexp(cmf), eb = cexp(cmb);
    phf = ef * Af; ddf = ef * (c * Sf);
    phb = eb * Ab; ddb = eb * (c * Sb);
  }
}

// Coefficients of E(x) = alpha_b(x) + (alpha_f(x) - alpha_b(x)) q_b(x) h(x) mod q_f(x) q_b(x), with
// alpha(x) = (phi_s - dd lam_s) + dd x the linear interpolant of exp(c x) on each root pair
template <typename T> __device__ __forceinline__ void cubic_coeffs(const Spectrum<T>& s, cx<T> h0, cx<T> h1, cx<T> lf, cx<T> phf, cx<T> ddf,
                                                                  cx<T> lb, cx<T> phb, cx<T> ddb, cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
  cx<T> ab0 = fnma(ddb, lb, phb);
  cx<T> d0 = fnma(ddf, lf, phf) - ab0, d1 = ddf - ddb;
  cx<T> g0 = d0 * h0, g1 = fma(d1, h0, d0 * h1), g2 = d1 * h1;
  // q_f q_b = x^4 + c3 x^3 + c2 x^2 + c1 x + c0
  cx<T> c3 = -(s.sf + s.sb), c2 = fma(s.sf, s.sb, s.pf + s.pb), c1 = -fma(s.sf, s.pb, s.sb * s.pf), c0 = s.pf * s.pb;
  a3 = fnma(g2, s.sb + c3, g1);
  a2 = fma(g2, s.pb - c2, fnma(s.sb, g1, g0));
  a1 = fnma(g2, c1, fma(s.pb, g1, fnma(s.sb, g0, ddb)));
  a0 = fnma(g2, c0, fma(s.pb, g0, ab0));
}

// y <- (a0 + a1 D + a2 D^2 + a3 D^3) y: three mat-vecs, one accumulator
template <typename T, bool NM> __device__ __forceinline__ void cubic_apply(const Delta<T, NM>& D, cx<T> a0, cx<T> a1, cx<T> a2, cx<T> a3, Vec4<T>& y) {
  Vec4<T> v = matvec(D, y);
  Vec4<T> o = axpy(a1, v, scale(a0, y));
  v = matvec(D, v);
  o = axpy(a2, v, o);
  v = matvec(D, v);
  y = axpy(a3, v, o);
}

// exp(c D) Y, unnormalised (transfer-matrix semantics: overflows for thick evanescent layers).
// Thin films (all four |Re(c lambda)| < 3, i.e. mode exponentials within e^6 of each other) take
// the exponential as ONE cubic polynomial in D,
//   E(x) = alpha_b(x) + (alpha_f(x) - alpha_b(x)) q_b(x) h(x)   mod   q_f(x) q_b(x),
// alpha_{f,b} the linear interpolants of exp(c x) on each root pair: 6 mat-vecs instead of 8 and
// no projected intermediates.  Mixing all four exponentials in one polynomial costs log10 of their
// ratio in digits (<= 2.6 here); thick films -- where that form loses the sub-dominant growing
// mode, SURVEY 7.0-6a -- keep the invariant-subspace block form.
// Coefficients of that cubic from the spectral split; false when the film is not thin.
template <typename T, bool SCALE_FREE> __device__ __forceinline__ bool film_thin_coeffs(const Spectrum<T>& s, cx<T> c, cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
  cx<T> lf = T(0.5) * s.sf, lb = T(0.5) * s.sb, phf, ddf, phb, ddb;   // centres of the two root pairs
  const cx<T> cc = c * c, cmf = c * lf, cmb = c * lb;
  const cx<T> wf = cc * fma(lf, lf, -s.pf), wb = cc * fma(lb, lb, -s.pb);   // (c x half-distance)^2
  if (!(thin_pair(cmf.re, wf) && thin_pair(cmb.re, wb))) return false;
  cx<T> h0, h1;
  projector_coeffs(s, h0, h1);
  thin_pairs<T, SCALE_FREE>(cmf, wf, cmb, wb, c, phf, ddf, phb, ddb);
  cubic_coeffs(s, h0, h1, lf, phf, ddf, lb, phb, ddb, a0, a1, a2, a3);
  return true;
}

// The same for the plane an isotropic exit (or isotropic crystal exit) leaves behind,
//   y0 = (0, p, q, 0),  y1 = (r, 0, 0, s)   (the s and the p wave: exit_iso_plane, iso_crystal_exit_plane):
// the first mat-vec and the first accumulation only touch the non-zero entries -- 4 + 6 instead of
// 10 + 10 complex products (non-magnetic Delta) -- and, for y1 with a scalar mu, the second mat-vec
// skips the row that is still zero.  Every film that sits directly on such an exit takes this.
template <typename T, bool NM> __device__ __forceinline__ void cubic_apply_sparse(const Delta<T, NM>& D, cx<T> a0, cx<T> a1, cx<T> a2, cx<T> a3, Vec4<T>& y0, Vec4<T>& y1) {
  {  // y0 = (0, p, q, 0)
    const cx<T> p = y0.v1, q = y0.v2;
    Vec4<T> v;
    if constexpr (NM) {
      v.v0 = D.b * p; v.v1 = D.f * q; v.v2 = D.h * p; v.v3 = -(D.g * p);
    } else {
      v.v0 = fma(D.c, q, D.b * p); v.v1 = fma(D.f, q, D.e * p); v.v2 = fma(D.e, q, D.h * p); v.v3 = -(D.g * p);
    }
    Vec4<T> o;
    o.v0 = a1 * v.v0; o.v1 = fma(a1, v.v1, a0 * p); o.v2 = fma(a1, v.v2, a0 * q); o.v3 = a1 * v.v3;
    v = matvec(D, v);
    o = axpy(a2, v, o);
    v = matvec(D, v);
    y0 = axpy(a3, v, o);
  }
  {  // y1 = (r, 0, 0, s)
    const cx<T> r = y1.v0, t = y1.v3;
    Vec4<T> v;
    v.v0 = fma(D.d, t, D.a * r);
    v.v2 = fnma(D.b, t, D.g * r);
    v.v3 = fma(D.a, t, D.j * r);
    Vec4<T> o;
    o.v0 = fma(a1, v.v0, a0 * r); o.v2 = a1 * v.v2; o.v3 = fma(a1, v.v3, a0 * t);
    if constexpr (NM) {
      // v.v1 = 0: second mat-vec without the column that multiplies it
      Vec4<T> u;
      u.v0 = fma(D.d, v.v3, D.a * v.v0);
      u.v1 = D.f * v.v2;
      u.v2 = fnma(D.b, v.v3, D.g * v.v0);
      u.v3 = fma(D.a, v.v3, D.j * v.v0);
      o.v0 = fma(a2, u.v0, o.v0); o.v1 = a2 * u.v1; o.v2 = fma(a2, u.v2, o.v2); o.v3 = fma(a2, u.v3, o.v3);
      v = matvec(D, u);
    } else {
      v.v1 = -(D.c * t);
      o.v1 = a1 * v.v1;
      v = matvec(D, v);
      o = axpy(a2, v, o);
      v = matvec(D, v);
    }
    y1 = axpy(a3, v, o);
  }
}

// The cubic applied to the running plane; stable recursion: then re-based to unit (Ex, Ey) rows,
// which keeps the basis well conditioned so that the <= 2.6 digits the cubic costs do not compound
// over layers (W = the re-basing matrix, for the power bookkeeping).
template <typename T, bool NM, bool STABLE> __device__ __forceinline__ void film_cubic(const Delta<T, NM>& D, cx<T> a0, cx<T> a1, cx<T> a2, cx<T> a3,
                                                                                Vec4<T>& y0, Vec4<T>& y1, cx<T> w[4], bool sparse = false) {
  if (sparse) {
    cubic_apply_sparse(D, a0, a1, a2, a3, y0, y1);
  } else {
    cubic_apply(D, a0, a1, a2, a3, y0);
    cubic_apply(D, a0, a1, a2, a3, y1);
  }
  if constexpr (STABLE) {
    inv2(y0.v0, y1.v0, y0.v1, y1.v1, w[0], w[1], w[2], w[3]);
    Vec4<T> t0 = axpy(w[2], y1, scale(w[0], y0));
    y1 = axpy(w[3], y1, scale(w[1], y0));
    y0 = t0;
  } else {
    w[0] = mk<T>(T(1), T(0)); w[1] = mk<T>(T(0), T(0)); w[2] = w[1]; w[3] = w[0];
  }
}

// Thick film, transfer recursion: the invariant-subspace block form.
template <typename T, bool NM> __device__ __forceinline__ void film_transfer_thick(const Delta<T, NM>& D, const Spectrum<T>& s, cx<T> c, Vec4<T>& y0, Vec4<T>& y1, cx<T> w[4]) {
  cx<T> h0, h1;
  projector_coeffs(s, h0, h1);
  cx<T> lf, phf, ddf, lb, phb, ddb;
  T gf, gb;
  w[0] = mk<T>(T(1), T(0)); w[1] = mk<T>(T(0), T(0)); w[2] = w[1]; w[3] = w[0];
  bool sepf, sepb;
  pair_modes(s.sf, s.pf, c, lf, phf, ddf, sepf, gf);
  pair_modes(s.sb, s.pb, c, lb, phb, ddb, sepb, gb);
  Vec4<T> dy0, dy1;
  Vec4<T> zf0 = project_forward(D, s, h0, h1, y0, dy0);
  Vec4<T> zf1 = project_forward(D, s, h0, h1, y1, dy1);
  Vec4<T> dzf0 = matvec(D, zf0), dzf1 = matvec(D, zf1);
  Vec4<T> b0, b1;
  block_exp_apply(lb, phb, ddb, sepb, sub(y0, zf0), sub(y1, zf1), sub(dy0, dzf0), sub(dy1, dzf1), b0, b1);
  Vec4<T> r0 = axpy(-lf, zf0, dzf0), r1 = axpy(-lf, zf1, dzf1);
  if (!sepf) {
    y0 = add(axpy(ddf, r0, scale(phf, zf0)), b0);
    y1 = add(axpy(ddf, r1, scale(phf, zf1)), b1);
    return;
  }
  // The two forward modes grow at different rates: [r0 r1] is rank one (a multiple of the faster
  // eigenvector), r_other = beta r_dom.  Adding phi_s z_k to the huge dd r_k would drown the
  // slower mode in rounding error (ratio eps phi_g / phi_s), so the plane is re-based by the
  // column operation  other <- other - beta dom, whose dominant parts cancel analytically:
  //   dom   = phi_s z_dom + dd r_dom + b_dom,     other = phi_s (z_oth - beta z_dom) + (b_oth - beta b_dom).
  // Same plane (r_ij unchanged), and w records the operation for the coordinate walk (t_ij, power).
  bool first;
  cx<T> beta;
  rank1_coeff(r0, r1, first, beta);
  if (first) {
    y0 = add(axpy(ddf, r0, scale(phf, zf0)), b0);
    y1 = add(scale(phf, axpy(-beta, zf0, zf1)), axpy(-beta, b0, b1));
    w[1] = -beta;
  } else {
    y1 = add(axpy(ddf, r1, scale(phf, zf1)), b1);
    y0 = add(scale(phf, axpy(-beta, zf1, zf0)), axpy(-beta, b1, b0));
    w[2] = -beta;
  }
}

// Film of an un-tilted crystal (z a principal axis up to the reference's hidden 1e-8 tilt, any
// in-plane rotation): det(x I - D) is a biquadratic x^4 + c2 x^2 + c0 plus odd coefficients of
// the order of the tilt, so the roots come in pairs delta_i +- mu_i with delta_i ~ tilt.
//   1. start from the biquadratic limit, q_0 = x^2 - y0 (y0, y1 the roots of y^2 + c2 y + c0).
//      The Bairstow Jacobian there is -(y0 - y1) I up to O(tilt), so the polish is the fixed-point
//      step (u, v) += (r1, r0) / (y0 - y1): linear convergence at rate ~ tilt |mu| / |y0 - y1|,
//      three steps, the last one checked.
//   2. on each pair exp(c x) is interpolated by
//        alpha_i(x) = e^{c delta_i} [cosh(c mu_i) + (x - delta_i) sinh(c mu_i) / mu_i],
//      entire and even in mu_i: no forward/backward classification, one exponential per pair.
//   3. the two interpolants are merged into one cubic by the projector formula of cubic_coeffs.
// Uniaxial film (`uni`, q_o = kx^2 - mu eps_o; any orientation): the two pairs are known in closed
// form -- the ordinary pair +-mu_o, mu_o^2 = -q_o, and the extraordinary pair delta_e +- mu_e of
// x^2 + c3 x + q_e, q_e = c2 - q_o -- so step 1 disappears (c1 and c0 are not even computed), and
// e^{c delta_e} is a true exponential when the crystal is tilted (delta_e = -c3/2 not small).
// Returns false when the point has to take the general path: odd coefficients not small, y0 ~ y1
// (the merge would lose more than 3 digits: near-normal incidence on a c-cut uniaxial film),
// polish not converged, or thick film (a |Re(c mu)| >= 3).
// Step 1 for a general un-tilted crystal: the two pair factors (as a Spectrum: "f" = pair 0, "b" = pair 1).
template <typename T> __device__ __forceinline__ bool pair_factors_untilted(cx<T> c3, cx<T> c2, cx<T> c1, cx<T> c0, Spectrum<T>& s) {
  const T odd_tol = sizeof(T) == 8 ? T(1e-12) : T(1e-6);
  const T step_tol = sizeof(T) == 8 ? T(1e-20) : T(1e-9);
  T n2 = norm2(c2), n3 = norm2(c3), n1 = norm2(c1);
  if (!((n3 * n3 <= odd_tol * n2) && (n1 * n1 <= odd_tol * (n2 * n2 * n2)))) return false;
  cx<T> y0, y1;
  quad_roots(c2, c0, y0, y1);
  cx<T> dy = y0 - y1;
  if (!(norm2(dy) >= T(1e-6) * (norm2(y0) + norm2(y1)))) return false;
  cx<T> idy = recip(dy);
  cx<T> u = mk<T>(T(0), T(0)), v = -y0, b1, b0, du, dv;
  // two steps; a third only where the second still moved (rate ~ tilt |mu| / |y0 - y1|: step k has
  // size rate^k, the error after it rate^(k+1))
  bool done = false;
#pragma unroll 1
  for (int it = 0; it < 3 && !done; ++it) {
    b1 = c3 - u;
    b0 = fnma(u, b1, c2 - v);
    du = fnma(v, b1, fnma(u, b0, c1)) * idy;
    dv = fnma(v, b0, c0) * idy;
    u = u + du;
    v = v + dv;
    T nv = norm2(v), ndu = norm2(du);
    done = it >= 1 && norm2(dv) <= step_tol * nv && ndu * ndu <= step_tol * step_tol * nv;  // |du|^2 <= tol |v|
  }
  if (!done) return false;
  b1 = c3 - u;
  b0 = fnma(u, b1, c2 - v);
  s.sf = -u; s.pf = v; s.sb = -b1; s.pb = b0;
  return true;
}
// Step 1 for a uniaxial crystal: the ordinary and the extraordinary factor, exact (the host vouches
// for the hint: plan._ordinary_value checks the rotated table against it).  Only c3 and c2 are needed.
template <typename T> __device__ __forceinline__ void pair_factors_uniaxial(cx<T> c3, cx<T> c2, cx<T> qo, Spectrum<T>& s) {
  s.sf = mk<T>(T(0), T(0)); s.pf = qo; s.sb = -c3; s.pb = c2 - qo;
}
// Steps 2 and 3.  `exp_delta`: pair 1 may sit off the origin (tilted uniaxial film: delta_e = -c3/2
// not small), e^{c delta_e} is then a true exponential.
template <typename T> __device__ __forceinline__ bool pair_symmetric_film_coeffs(const Spectrum<T>& s, cx<T> c, bool exp_delta,
                                                                                cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
  cx<T> d0 = T(0.5) * s.sf, d1 = T(0.5) * s.sb;   // pair centres delta_i
  const cx<T> cc = c * c;
  cx<T> m0 = fma(d0, d0, -s.pf), m1 = fma(d1, d1, -s.pb);              // mu_i^2
  // the merge of the two interpolants divides by the resultant ~ (mu_0^2 - mu_1^2)^2: y0 ~ y1 loses digits
  if (!(norm2(m0 - m1) >= T(1e-6) * (norm2(m0) + norm2(m1)))) return false;
  cx<T> w0 = cc * m0, w1 = cc * m1;   // (c mu_i)^2
  cx<T> x0 = c * d0, x1 = c * d1;
  if (!(thin_pair(x0.re, w0) && thin_pair(x1.re, w1) && norm2(x0) < T(1e-10))) return false;
  cx<T> A0, S0, A1, S1;
  cosh_sinhc_pair(w0, w1, A0, S0, A1, S1);
  const cx<T> one = mk<T>(T(1), T(0));
  cx<T> e0 = fma(T(0.5) * x0, x0, one + x0), e1;  // e^{c delta}, |c delta| < 1e-5
  if (norm2(x1) < T(1e-10)) e1 = fma(T(0.5) * x1, x1, one + x1);
  else if (exp_delta) e1 = cexp(x1);
  else return false;
  cx<T> h0, h1;
  projector_coeffs(s, h0, h1);
  cubic_coeffs(s, h0, h1, d0, e0 * A0, e0 * (c * S0), d1, e1 * A1, e1 * (c * S1), a0, a1, a2, a3);
  return true;
}

// 2x2 inverse of [[m00, m01], [m10, m11]]
template <typename T> __device__ __forceinline__ void inv2(cx<T> m00, cx<T> m01, cx<T> m10, cx<T> m11, cx<T>& n00, cx<T>& n01, cx<T>& n10, cx<T>& n11) {
  cx<T> idet = recip(fnma(m01, m10, m00 * m11));
  n00 = m11 * idet; n01 = -(m01 * idet); n10 = -(m10 * idet); n11 = m00 * idet;
}

// A I + B (D - lam I) applied to x, given dx = D x
template <typename T> __device__ __forceinline__ Vec4<T> lin_apply(cx<T> phi, cx<T> dd, cx<T> lam, const Vec4<T>& x, const Vec4<T>& dx) {
  return axpy(dd, axpy(-lam, x, dx), scale(phi, x));
}

// Stable propagation of the plane span(Y) through a film: only decaying exponentials.
// Renormalises so that rows (Ex, Ey) of the forward part are the identity.
// w[0..3] = (w00, w01, w10, w11): the 2x2 map W with Y_new = exp(cD) Y_old W, i.e. coordinates at
// the top of the film -> coordinates at its bottom (used by the power bookkeeping only).
template <typename T, bool NM> __device__ __forceinline__ void film_stable_thick(const Delta<T, NM>& D, const Spectrum<T>& s, cx<T> c, Vec4<T>& y0, Vec4<T>& y1, cx<T> w[4]) {
  cx<T> h0, h1;
  projector_coeffs(s, h0, h1);
  cx<T> lf, phf, ddf, lb, phb, ddb;
  T gf, gb;
  bool sepf, sepb;
  pair_modes(s.sf, s.pf, c, lf, phf, ddf, sepf, gf);
  pair_modes(s.sb, s.pb, c, lb, phb, ddb, sepb, gb);
  Vec4<T> dy;
  Vec4<T> zf0 = project_forward(D, s, h0, h1, y0, dy);
  Vec4<T> zf1 = project_forward(D, s, h0, h1, y1, dy);
  Vec4<T> zb0 = sub(y0, zf0), zb1 = sub(y1, zf1);
  cx<T> n00, n01, n10, n11;
  inv2(zf0.v0, zf1.v0, zf0.v1, zf1.v1, n00, n01, n10, n11);
  Vec4<T> xf0 = axpy(n10, zf1, scale(n00, zf0)), xf1 = axpy(n11, zf1, scale(n01, zf0));
  Vec4<T> xb0 = axpy(n10, zb1, scale(n00, zb0)), xb1 = axpy(n11, zb1, scale(n01, zb0));
  cx<T> lam, phi, dd;
  bool sep;
  // g = L exp(-c K_f) X_f  (2x2): forward modes run back down the film, decaying
  T growth;
  pair_modes(s.sf, s.pf, -c, lam, phi, dd, sep, growth);
  Vec4<T> g0 = lin_apply(phi, dd, lam, xf0, matvec(D, xf0));
  Vec4<T> g1 = lin_apply(phi, dd, lam, xf1, matvec(D, xf1));
  // exp(+c K_b) X_b: backward modes up the film, decaying (modes of the pair already known)
  Vec4<T> e0 = lin_apply(phb, ddb, lb, xb0, matvec(D, xb0));
  Vec4<T> e1 = lin_apply(phb, ddb, lb, xb1, matvec(D, xb1));
  y0 = axpy(g0.v1, e1, axpy(g0.v0, e0, xf0));
  y1 = axpy(g1.v1, e1, axpy(g1.v0, e0, xf1));
  w[0] = fma(n01, g0.v1, n00 * g0.v0); w[1] = fma(n01, g1.v1, n00 * g1.v0);
  w[2] = fma(n11, g0.v1, n10 * g0.v0); w[3] = fma(n11, g1.v1, n10 * g1.v0);
}

// exp(c D) Y with the exact scale, thin or thick (the materialising kernels of ho_modes.cuh)
template <typename T, bool NM> __device__ __forceinline__ void film_transfer(const Delta<T, NM>& D, const Spectrum<T>& s, cx<T> c, Vec4<T>& y0, Vec4<T>& y1, cx<T> w[4]) {
  cx<T> a0, a1, a2, a3;
  if (film_thin_coeffs<T, false>(s, c, a0, a1, a2, a3)) film_cubic<T, NM, false>(D, a0, a1, a2, a3, y0, y1, w);
  else film_transfer_thick(D, s, c, y0, y1, w);
}

// Forward invariant subspace of a semi-infinite crystal: q_b(D) [e0 e1], q_b(x) = x^2 - s_b x + p_b.
// q_b(D) annihilates the backward subspace and is invertible on the forward one, so its columns
// span the same plane as P_f [e0 e1] (r_*, the Mueller matrix and the flux ratios only depend on
// the plane); D e_k is just column k of D, so each column costs one mat-vec.
template <typename T, bool NM> __device__ __forceinline__ void substrate_plane(const Delta<T, NM>& D, const Spectrum<T>& s, Vec4<T>& y0, Vec4<T>& y1) {
  const cx<T> zero = mk<T>(T(0), T(0));
  Vec4<T> col;
  // column 0 of D is (a, 0, g, j)
  col.v0 = D.a; col.v1 = zero; col.v2 = D.g; col.v3 = D.j;
  y0 = axpy(-s.sb, col, matvec(D, col));
  y0.v0 = y0.v0 + s.pb;
  // column 1 of D is (b, e, h, -g)
  col.v0 = D.b; col.v1 = NM ? zero : D.e; col.v2 = D.h; col.v3 = -D.g;
  y1 = axpy(-s.sb, col, matvec(D, col));
  y1.v1 = y1.v1 + s.pb;
}

// ---- isotropic layers: D^2 = kz^2 I, closed forms (reference-equivalent) ----
template <typename T> __device__ __forceinline__ Vec4<T> iso_matvec(cx<T> d, cx<T> f, cx<T> h, cx<T> j, const Vec4<T>& v) {
  Vec4<T> r; r.v0 = d * v.v3; r.v1 = f * v.v2; r.v2 = h * v.v1; r.v3 = j * v.v0; return r;
}

template <typename T> __device__ __forceinline__ void iso_film(T k, cx<T> eps, cx<T> mu, cx<T> c, bool stable, Vec4<T>& y0, Vec4<T>& y1, cx<T> w[4]) {
  T k2 = k * k;
  cx<T> d = fnma(k2, recip(eps), mu), f = -mu, h = fma(k2, recip(mu), -eps), j = eps;
  cx<T> kz2 = eps * mu - k2;
  Vec4<T> dy0 = iso_matvec(d, f, h, j, y0), dy1 = iso_matvec(d, f, h, j, y1);
  if (!stable) {
    // exp(c D) = cosh(c kz) I + c sinhc(c kz) D  (even in kz: no branch choice needed)
    cx<T> A, S;
    cosh_sinhc_w((c * c) * kz2, A, S);
    cx<T> B = c * S;
    y0 = axpy(B, dy0, scale(A, y0));
    y1 = axpy(B, dy1, scale(A, y1));
    w[0] = mk<T>(T(1), T(0)); w[1] = mk<T>(T(0), T(0)); w[2] = w[1]; w[3] = w[0];
    return;
  }
  // stable: forward branch of kz by the reference's mode rule (waves.py:273-289)
  cx<T> kz = csqrt(kz2);
  bool fwd = (ho_abs(kz.im) > T(1e-9)) ? (kz.im > T(0)) : (kz.re > T(0));
  if (!fwd) kz = -kz;
  cx<T> ikz = recip(kz);
  cx<T> ph2 = cexp(T(-2) * (c * kz));  // exp(2 i tau kz): |.| <= 1
  Vec4<T> zf0 = scale(mk<T>(T(0.5), T(0)), axpy(ikz, dy0, y0)), zf1 = scale(mk<T>(T(0.5), T(0)), axpy(ikz, dy1, y1));
  Vec4<T> zb0 = sub(y0, zf0), zb1 = sub(y1, zf1);
  cx<T> n00, n01, n10, n11;
  inv2(zf0.v0, zf1.v0, zf0.v1, zf1.v1, n00, n01, n10, n11);
  y0 = axpy(ph2, axpy(n10, zb1, scale(n00, zb0)), axpy(n10, zf1, scale(n00, zf0)));
  y1 = axpy(ph2, axpy(n11, zb1, scale(n01, zb0)), axpy(n11, zf1, scale(n01, zf0)));
  cx<T> ph = cexp(-(c * kz));  // exp(i tau kz): |.| <= 1
  w[0] = n00 * ph; w[1] = n01 * ph; w[2] = n10 * ph; w[3] = n11 * ph;
}

// Forward plane of a semi-infinite ISOTROPIC crystal (scalar complex eps(f), mu(f)): the modes of
// SURVEY A.13, s = (0, -mu, kz, 0) and p = (mu - k^2/eps, 0, 0, kz), kz the forward root of
// eps mu - k^2 by the reference's rule (waves.py:273-289).  Same plane as the general path's
// q_b(D) [e0 e1] for that medium (double forward root), without the quartic.
template <typename T> __device__ __forceinline__ void iso_crystal_exit_plane(T k, cx<T> eps, cx<T> mu, Vec4<T>& y0, Vec4<T>& y1) {
  const cx<T> zero = mk<T>(T(0), T(0));
  T k2 = k * k;
  cx<T> kz = forward_sqrt(eps * mu - k2);
  y0.v0 = zero; y0.v1 = -mu; y0.v2 = kz; y0.v3 = zero;
  y1.v0 = fnma(k2, recip(eps), mu); y1.v1 = zero; y1.v2 = zero; y1.v3 = kz;
}

// columns 0 and 2 of the isotropic exit matrix (reference layers.py:190-238)
template <typename T> __device__ __forceinline__ void exit_iso_plane(cx<T> cf, T n_exit, Vec4<T>& y0, Vec4<T>& y1) {
  const cx<T> one = mk<T>(T(1), T(0)), zero = mk<T>(T(0), T(0));
  y0.v0 = zero; y0.v1 = one; y0.v2 = -(n_exit * cf); y0.v3 = zero;
  y1.v0 = cf; y1.v1 = zero; y1.v2 = zero; y1.v3 = mk<T>(n_exit, T(0));
}

// prism rows + 2x2 minors (reference layers.py:109-152, structure.py:285-304); the common 1/2 cancels
template <typename T> __device__ __forceinline__ void reflect_from_plane(const Vec4<T>& y0, const Vec4<T>& y1, T inv_c, T inv_nc, T inv_n,
                                                                        cx<T>& r_pp, cx<T>& r_ps, cx<T>& r_sp, cx<T>& r_ss) {
  cx<T> g00 = fnma(inv_nc, y0.v2, y0.v1), g10 = fma(inv_nc, y0.v2, y0.v1);
  cx<T> n0 = inv_n * y0.v3, n1 = inv_n * y1.v3;
  cx<T> g20 = fma(inv_c, y0.v0, n0), g30 = fnma(inv_c, y0.v0, n0);
  cx<T> g02 = fnma(inv_nc, y1.v2, y1.v1), g12 = fma(inv_nc, y1.v2, y1.v1);
  cx<T> g22 = fma(inv_c, y1.v0, n1), g32 = fnma(inv_c, y1.v0, n1);
  cx<T> ib = recip(fnma(g02, g20, g00 * g22));
  r_pp = fnma(g30, g02, g00 * g32) * ib;
  r_ps = fnma(g10, g02, g00 * g12) * ib;
  r_sp = fnma(g32, g20, g30 * g22) * ib;
  r_ss = fnma(g12, g20, g10 * g22) * ib;
}

// ---- power bookkeeping (reference fields.py:45-51, 120-282) ----
// Hermitian 2x2 form of S_z = 1/2 Re(Ex Hy* - Ey Hx*) on the plane Y c:
//   S_z = h[0] |c0|^2 + h[1] |c1|^2 + Re(c0 conj(c1) (h[2] + i h[3]))
template <typename T> __device__ __forceinline__ void flux_form(const Vec4<T>& y0, const Vec4<T>& y1, T h[4]) {
  cx<T> q00 = cmulc(y0.v3, y0.v0) - cmulc(y0.v2, y0.v1);
  cx<T> q11 = cmulc(y1.v3, y1.v0) - cmulc(y1.v2, y1.v1);
  cx<T> q01 = cmulc(y1.v3, y0.v0) - cmulc(y1.v2, y0.v1);
  cx<T> q10 = cmulc(y0.v3, y1.v0) - cmulc(y0.v2, y1.v1);
  h[0] = T(0.5) * q00.re; h[1] = T(0.5) * q11.re;
  h[2] = T(0.5) * (q01.re + q10.re); h[3] = T(0.5) * (q01.im - q10.im);
}
template <typename T> __device__ __forceinline__ T flux_eval(const T h[4], cx<T> c0, cx<T> c1) {
  cx<T> z = cmulc(c1, c0);
  return h[0] * norm2(c0) + h[1] * norm2(c1) + (z.re * h[2] - z.im * h[3]);
}
// Coordinates c of the field on the top plane for unit incident s / p amplitude: solves
// [[G00, G02], [G20, G22]] c = (a_s, a_p), G = M_prism Y with the prism's 1/2 kept (fields.py:120-145)
template <typename T> __device__ __forceinline__ void incident_coords(const Vec4<T>& y0, const Vec4<T>& y1, T inv_c, T inv_nc, T inv_n,
                                                                     cx<T>& cp0, cx<T>& cp1, cx<T>& cs0, cx<T>& cs1) {
  cx<T> g00 = T(0.5) * (y0.v1 - inv_nc * y0.v2), g20 = T(0.5) * (inv_c * y0.v0 + inv_n * y0.v3);
  cx<T> g02 = T(0.5) * (y1.v1 - inv_nc * y1.v2), g22 = T(0.5) * (inv_c * y1.v0 + inv_n * y1.v3);
  cx<T> idet = recip(g00 * g22 - g02 * g20);
  cp0 = -(g02 * idet); cp1 = g00 * idet;   // (a_s, a_p) = (0, 1)
  cs0 = g22 * idet;    cs1 = -(g20 * idet);  // (a_s, a_p) = (1, 0)
}

// ---- transmission coefficients: the reference's mode basis of a crystal exit ----
// LAPACK zgeev convention (what numpy.linalg.eig returns, reference waves.py:270): unit 2-norm,
// the component of largest modulus (first on ties) real positive.
template <typename T> __device__ __forceinline__ Vec4<T> lapack_normalise(const Vec4<T>& v) {
  T n0 = norm2(v.v0), n1 = norm2(v.v1), n2 = norm2(v.v2), n3 = norm2(v.v3);
  T inv = ho_rsqrt(n0 + n1 + n2 + n3);
  cx<T> big = v.v0; T nb = n0;
  if (n1 > nb) { big = v.v1; nb = n1; }
  if (n2 > nb) { big = v.v2; nb = n2; }
  if (n3 > nb) { big = v.v3; nb = n3; }
  cx<T> ph = (inv * ho_rsqrt(nb)) * conj(big);  // unit phase / norm
  return scale(ph, v);
}

// p-likeness measures of a mode (reference waves.py:382-408, 490-497): complex Poynting E x H
// (no conjugate) and the electric field
template <typename T> __device__ __forceinline__ void mode_character(const Vec4<T>& v, const ZRow<T>& z, T k, T& cp, T& ce) {
  cx<T> ez = -((k * v.v3 + z.e20 * v.v0 + z.e21 * v.v1) * z.ie22);
  cx<T> hz = (k * v.v1 - z.m20 * v.v2 - z.m21 * v.v3) * z.im22;
  T px = norm2(v.v1 * hz - ez * v.v3), py = norm2(ez * v.v2 - v.v0 * hz);
  T ex = norm2(v.v0), ey = norm2(v.v1);
  cp = px / (px + py);
  ce = ex / (ex + ey);
}

// binv (row-major 2x2): coordinates in the kernel's forward-plane basis [y0 y1] -> amplitudes of
// the reference's exit modes (t0, t1) = eigenvectors in LAPACK normalisation, ordered by
// sort_poynting_indices (waves.py:474-510).  Degenerate forward pair (isotropic crystal): the
// reference's basis is whatever LAPACK returns for a 2-dimensional eigenspace -- not reproducible;
// the kernel basis normalised to unit E rows is used instead (binv = L [y0 y1]).
template <typename T, bool NM> __device__ __forceinline__ void exit_mode_basis(const Delta<T, NM>& D, const Spectrum<T>& s, const ZRow<T>& z, T k,
                                                                          const Vec4<T>& y0, const Vec4<T>& y1, cx<T> binv[4]) {
  cx<T> l0, l1;
  quad_roots(-s.sf, s.pf, l0, l1);
  cx<T> n00, n01, n10, n11;
  if (norm2(l0 - l1) <= T(1e-20) * (norm2(l0) + norm2(l1))) {
    binv[0] = y0.v0; binv[1] = y1.v0; binv[2] = y0.v1; binv[3] = y1.v1;
    return;
  }
  Vec4<T> dy0 = matvec(D, y0), dy1 = matvec(D, y1);
  // eigenvector of l0: (D - l1) y, of l1: (D - l0) y; take the better-conditioned seed column
  Vec4<T> a0 = axpy(-l1, y0, dy0), a1 = axpy(-l1, y1, dy1);
  Vec4<T> b0 = axpy(-l0, y0, dy0), b1 = axpy(-l0, y1, dy1);
  Vec4<T> u = lapack_normalise(sel(norm2(a0) >= norm2(a1), a0, a1));
  Vec4<T> v = lapack_normalise(sel(norm2(b0) >= norm2(b1), b0, b1));
  T cpu_, ceu, cpv, cev;
  mode_character(u, z, k, cpu_, ceu);
  mode_character(v, z, k, cpv, cev);
  // order: larger C_P first when they differ by more than 1e-6, else smaller C_E first; the
  // reference's incoming order is Im(lambda) descending (ties keep that order)
  bool u_first_in = l0.im >= l1.im;
  bool swap;
  if (ho_abs(cpu_ - cpv) > T(1e-6)) swap = cpv > cpu_;
  else swap = (cev < ceu) || (cev == ceu && !u_first_in);
  Vec4<T> t0 = sel(swap, v, u), t1 = sel(swap, u, v);
  // B = (L Y)^-1 L [t0 t1], binv = B^-1
  inv2(y0.v0, y1.v0, y0.v1, y1.v1, n00, n01, n10, n11);
  cx<T> b00 = fma(n01, t0.v1, n00 * t0.v0), b01 = fma(n01, t1.v1, n00 * t1.v0);
  cx<T> b10 = fma(n11, t0.v1, n10 * t0.v0), b11 = fma(n11, t1.v1, n10 * t1.v0);
  inv2(b00, b01, b10, b11, binv[0], binv[1], binv[2], binv[3]);
}

// Mueller matrix M = Re(A F A^-1), F = J (x) J* in the reference's layout (mueller.py:266-307),
// with the products written out: a = r_pp, b = r_ps, c = r_sp, d = r_ss,
//   row 0: (|a|^2+|b|^2+|c|^2+|d|^2)/2, (|a|^2-|b|^2+|c|^2-|d|^2)/2,  Re(ab*+cd*),  Im(ab*+cd*)
//   row 1: (|a|^2+|b|^2-|c|^2-|d|^2)/2, (|a|^2-|b|^2-|c|^2+|d|^2)/2,  Re(ab*-cd*),  Im(ab*-cd*)
//   row 2:  Re(ac*+bd*),                 Re(ac*-bd*),                  Re(ad*+bc*),  Im(ad*-bc*)
//   row 3: -Im(ac*+bd*),                -Im(ac*-bd*),                 -Im(ad*+bc*),  Re(ad*-bc*)
// (six complex products instead of the sixteen entries of F).
template <typename T> __device__ __forceinline__ void mueller_from_jones(cx<T> pp, cx<T> ps, cx<T> sp, cx<T> ss, T M[16]) {
  const T na = norm2(pp), nb = norm2(ps), nc = norm2(sp), nd = norm2(ss);
  const cx<T> ab = cmulc(ps, pp), cd = cmulc(ss, sp);   // a b*, c d*   (cmulc(x, y) = conj(x) y)
  const cx<T> ac = cmulc(sp, pp), bd = cmulc(ss, ps);   // a c*, b d*
  const cx<T> ad = cmulc(ss, pp), bc = cmulc(sp, ps);   // a d*, b c*
  const T h = T(0.5);
  M[0] = h * ((na + nb) + (nc + nd)); M[1] = h * ((na - nb) + (nc - nd));
  M[2] = ab.re + cd.re;               M[3] = ab.im + cd.im;
  M[4] = h * ((na + nb) - (nc + nd)); M[5] = h * ((na - nb) - (nc - nd));
  M[6] = ab.re - cd.re;               M[7] = ab.im - cd.im;
  M[8] = ac.re + bd.re;               M[9] = ac.re - bd.re;
  M[10] = ad.re + bc.re;              M[11] = ad.im - bc.im;
  M[12] = -(ac.im + bd.im);           M[13] = bd.im - ac.im;
  M[14] = -(ad.im + bc.im);           M[15] = ad.re - bc.re;
}

}  // namespace ho
