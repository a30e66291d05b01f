// Polarimetry epilogue: consumers of the Jones planes (r_ij or t_ij) that stay resident in HBM.
//
// Replaces, per batch point, the NumPy post-processing of the reference's
//   Mueller.calculate_stokes_parameters / get_*            mueller.py:479-631
//   Mueller.calculate_transmission_mueller_matrix         mueller.py:327-356
//   Mueller.decompose (Lu-Chipman)                        mueller.py:368-444
//   Jones.calculate_jones_vector / get_* / get_ellipse    jones.py:182-226
//   Jones.eigenpolarizations / find_exceptional_points    jones.py:228-347
//   Jones.ellipsometric_parameters                        jones.py:290-320
// One pass: 64 B of Jones coefficients read per point, only the requested planes written.
// HBM-bound (a few hundred flops per point); every output is a plane of n values so that the
// stores of a warp are contiguous, except the 4x4 matrices ([n][16], 128 contiguous bytes/thread).
#pragma once

#include "ho_b200.h"
#include "ho_berreman.cuh"

namespace ho {

__device__ __forceinline__ double deg(double rad) { return rad * 57.29577951308232; }

// numpy/LAPACK eigenvector convention: unit 2-norm, largest component real (positive)
__device__ __forceinline__ void unit_phase2(cx<double>& x, cx<double>& y) {
  double nx = norm2(x), ny = norm2(y);
  double inv = ho_rcp(ho_sqrt(nx + ny));
  cx<double> big = nx >= ny ? x : y;
  double ib = ho_rcp(ho_sqrt(nx >= ny ? nx : ny));
  cx<double> ph = mk<double>(big.re * ib * inv, -big.im * ib * inv);  // conj(big)/|big| / norm
  x = x * ph;
  y = y * ph;
}

// eigenvalues of a symmetric 3x3 (a00 a01 a02 / a11 a12 / a22), any order (trigonometric form)
__device__ __forceinline__ void sym3_eigenvalues(double a00, double a01, double a02, double a11, double a12, double a22,
                                                 double& e0, double& e1, double& e2) {
  double q = (a00 + a11 + a22) * (1.0 / 3.0);
  double p1 = a01 * a01 + a02 * a02 + a12 * a12;
  double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
  double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
  double p = ho_sqrt(p2 * (1.0 / 6.0));
  if (!(p > 0.0)) { e0 = e1 = e2 = q; return; }
  double ip = 1.0 / p;
  b00 *= ip; b11 *= ip; b22 *= ip;
  double c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
  double det = b00 * (b11 * b22 - c12 * c12) - c01 * (c01 * b22 - c12 * c02) + c02 * (c01 * c12 - b11 * c02);
  double r = fmin(1.0, fmax(-1.0, 0.5 * det));
  double phi = acos(r) * (1.0 / 3.0);
  e2 = q + 2.0 * p * cos(phi);
  e0 = q + 2.0 * p * cos(phi + 2.0943951023931953);
  e1 = 3.0 * q - e0 - e2;
}

// inverse of a general 3x3 (row-major) by the adjugate; returns det
__device__ __forceinline__ double inv3(const double m[9], double o[9]) {
  double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return det;
}

__device__ __forceinline__ void mul3(const double a[9], const double b[9], double o[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

// 16 contiguous doubles, streaming (written once, never re-read by the kernel)
__device__ __forceinline__ void store16(double* dst, const double m[16]) {
  double2* d2 = reinterpret_cast<double2*>(dst);
#pragma unroll
  for (int k = 0; k < 8; ++k) __stcs(d2 + k, make_double2(m[2 * k], m[2 * k + 1]));
}

// The same for a FULL warp holding 32 consecutive points: the 4 KB block is staged in shared memory
// (rows of 17 doubles: conflict-free) and written as eight fully coalesced 512-byte warp stores.
// The direct form above makes every warp store touch 32 different 128-byte lines with 16 bytes each
// -- invisible while a kernel is FP64-bound, but it capped the HBM-bound epilogue at 0.60 of peak.
constexpr int kStageRow = 17;
__device__ __forceinline__ void store16_warp(double* warp_dst, const double m[16], double* stage, int lane) {
#pragma unroll
  for (int k = 0; k < 16; ++k) stage[lane * kStageRow + k] = m[k];
  __syncwarp();
  double2* d2 = reinterpret_cast<double2*>(warp_dst);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int flat = j * 32 + lane, p = flat >> 3, d = flat & 7;
    __stcs(d2 + flat, make_double2(stage[p * kStageRow + 2 * d], stage[p * kStageRow + 2 * d + 1]));
  }
  __syncwarp();
}

// [[1, top^T], [left, sub]] -> 16 contiguous doubles
__device__ __forceinline__ void store_assembled(double* dst, const double top[3], const double left[3], const double sub[9]) {
  const double m[16] = {1.0, top[0], top[1], top[2], left[0], sub[0], sub[1], sub[2],
                        left[1], sub[3], sub[4], sub[5], left[2], sub[6], sub[7], sub[8]};
  store16(dst, m);
}

// Lu-Chipman polar decomposition M = M_delta M_R M_D of one Mueller matrix (mueller.py:368-444)
__device__ __forceinline__ void lu_chipman(const double M[16], double scalars[5], double* mats, int64_t n, int64_t i) {
  const double m00 = M[0];
  const double safe = fabs(m00) > 1e-12 ? m00 : 1.0;
  const double is = 1.0 / safe;
  double N[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) N[k] = M[k] * is;
  const double d[3] = {N[1], N[2], N[3]};
  const double pv[3] = {N[4], N[8], N[12]};
  const double D = ho_sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  const double P = ho_sqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
  const bool has_d = D > 1e-12;
  const double idn = 1.0 / (has_d ? D : 1.0);
  const double dh[3] = {d[0] * idn, d[1] * idn, d[2] * idn};
  const double root = ho_sqrt(fmax(1.0 - D * D, 0.0));
  double mD[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      mD[3 * r + c] = has_d ? ((r == c ? root : 0.0) + (1.0 - root) * dh[r] * dh[c]) : (r == c ? 1.0 : 0.0);
  // M' = N M_D^-1 with M_D^-1 = [[1, -d^T], [-d, m_D]] / (1 - D^2)   (exact block inverse)
  const double ia2 = 1.0 / (1.0 - D * D);
  double mp[9], pdel[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double* row = N + 4 * (r + 1);
    pdel[r] = (row[0] - (row[1] * d[0] + row[2] * d[1] + row[3] * d[2])) * ia2;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      mp[3 * r + c] = (-row[0] * d[c] + row[1] * mD[c] + row[2] * mD[3 + c] + row[3] * mD[6 + c]) * ia2;
  }
  // depolariser from the eigenvalues of G = m' m'^T
  double G[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) G[3 * r + c] = mp[3 * r] * mp[3 * c] + mp[3 * r + 1] * mp[3 * c + 1] + mp[3 * r + 2] * mp[3 * c + 2];
  double e0, e1, e2;
  sym3_eigenvalues(G[0], G[1], G[2], G[4], G[5], G[8], e0, e1, e2);
  const double s0 = ho_sqrt(fmax(e0, 0.0)), s1 = ho_sqrt(fmax(e1, 0.0)), s2 = ho_sqrt(fmax(e2, 0.0));
  const double c1 = s0 + s1 + s2, c2 = s0 * s1 + s1 * s2 + s2 * s0, c3 = s0 * s1 * s2;
  const double detmp = mp[0] * (mp[4] * mp[8] - mp[5] * mp[7]) - mp[1] * (mp[3] * mp[8] - mp[5] * mp[6]) + mp[2] * (mp[3] * mp[7] - mp[4] * mp[6]);
  const double sign = detmp < 0.0 ? -1.0 : 1.0;
  double A[9], B[9], Ai[9], mdep[9], mdi[9], mret[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const double eye = (k % 4 == 0) ? 1.0 : 0.0;
    A[k] = G[k] + c2 * eye;
    B[k] = c1 * G[k] + c3 * eye;
  }
  inv3(A, Ai);
  mul3(Ai, B, mdep);
#pragma unroll
  for (int k = 0; k < 9; ++k) mdep[k] *= sign;
  inv3(mdep, mdi);
  mul3(mdi, mp, mret);
  const double tr_ret = mret[0] + mret[4] + mret[8];
  scalars[0] = D;
  scalars[1] = P;
  scalars[2] = acos(fmin(1.0, fmax(-1.0, 0.5 * (tr_ret - 1.0))));
  scalars[3] = 1.0 - fabs(mdep[0] + mdep[4] + mdep[8]) * (1.0 / 3.0);
  scalars[4] = m00;
  if (mats) {
    const double zero3[3] = {0.0, 0.0, 0.0};
    store_assembled(mats + 16 * i, d, d, mD);
    store_assembled(mats + 16 * (n + i), zero3, zero3, mret);
    store_assembled(mats + 16 * (2 * n + i), zero3, pdel, mdep);
  }
}

}  // namespace ho

namespace {

constexpr int kPolarThreads = 256;

__device__ __forceinline__ void st_plain(ho_c128* p, ho::cx<double> v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.re, v.im)); }

// DECOMP: the Lu-Chipman branch is compiled in (128 registers, 2 CTAs/SM); the light build
// without it fits 3 CTAs/SM, which the streaming requests need to cover the HBM latency.
template <bool DECOMP>
__global__ void __launch_bounds__(kPolarThreads, DECOMP ? 2 : 3)
polarimetry_kernel(const __grid_constant__ ho_polarimetry_t q, int64_t n) {
  using namespace ho;
  typedef cx<double> C;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  __shared__ double stage[kPolarThreads / 32][32 * kStageRow];
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const C pp = ldc<double>(q.j_pp + i), ps = ldc<double>(q.j_ps + i);
    const C sp = ldc<double>(q.j_sp + i), ss = ldc<double>(q.j_ss + i);
    const bool want_m = q.stokes || q.polarisation || q.mueller || (DECOMP && (q.decomposition || q.decomposition_matrices));
    if (want_m) {
      double M[16];
      mueller_from_jones(pp, ps, sp, ss, M);
      if (q.mueller) {
        if (__activemask() == 0xffffffffu) store16_warp(q.mueller + 16 * (i - lane), M, stage[threadIdx.x >> 5], lane);
        else store16(q.mueller + 16 * i, M);   // ragged last warp
      }
      if (q.stokes || q.polarisation) {
        double v[4], s[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          v[r] = M[4 * r] * q.s_pre[0] + M[4 * r + 1] * q.s_pre[1] + M[4 * r + 2] * q.s_pre[2] + M[4 * r + 3] * q.s_pre[3];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          s[r] = q.m_post[4 * r] * v[0] + q.m_post[4 * r + 1] * v[1] + q.m_post[4 * r + 2] * v[2] + q.m_post[4 * r + 3] * v[3];
        if (q.stokes) {
#pragma unroll
          for (int r = 0; r < 4; ++r) q.stokes[r * n + i] = s[r];
        }
        if (q.polarisation) {
          const double lin = ho_sqrt(s[1] * s[1] + s[2] * s[2]);
          const double dop = ho_sqrt(s[1] * s[1] + s[2] * s[2] + s[3] * s[3]) / fmax(s[0], 1e-10);
          q.polarisation[i] = fmin(fmax(dop, 0.0), 1.0);
          q.polarisation[n + i] = 0.5 * atan2(s[3], lin);
          q.polarisation[2 * n + i] = 0.5 * atan2(s[2], s[1]);
        }
      }
      if (DECOMP && (q.decomposition || q.decomposition_matrices)) {
        double sc[5];
        lu_chipman(M, sc, q.decomposition_matrices, n, i);
        if (q.decomposition) {
#pragma unroll
          for (int r = 0; r < 5; ++r) q.decomposition[r * n + i] = sc[r];
        }
      }
    }
    if (q.jones_vector || q.jones_stokes) {
      const C e0 = mk<double>(q.e_pre[0].re, q.e_pre[0].im), e1 = mk<double>(q.e_pre[1].re, q.e_pre[1].im);
      const C u0 = fma(ps, e1, pp * e0), u1 = fma(ss, e1, sp * e0);
      const C a = mk<double>(q.j_post[0].re, q.j_post[0].im), b = mk<double>(q.j_post[1].re, q.j_post[1].im);
      const C c = mk<double>(q.j_post[2].re, q.j_post[2].im), d = mk<double>(q.j_post[3].re, q.j_post[3].im);
      const C ep = fma(b, u1, a * u0), es = fma(d, u1, c * u0);
      if (q.jones_vector) {
        st_plain(q.jones_vector + i, ep);
        st_plain(q.jones_vector + n + i, es);
      }
      if (q.jones_stokes) {
        const double np_ = norm2(ep), ns = norm2(es);
        const C x = cmulc(es, ep);  // conj(es) ep = ep es*
        const double s1 = np_ - ns, s2 = 2.0 * x.re, s3 = -2.0 * x.im;
        q.jones_stokes[i] = np_ + ns;
        q.jones_stokes[n + i] = s1;
        q.jones_stokes[2 * n + i] = s2;
        q.jones_stokes[3 * n + i] = s3;
        q.jones_stokes[4 * n + i] = 0.5 * atan2(s2, s1);
        q.jones_stokes[5 * n + i] = 0.5 * atan2(s3, ho_sqrt(s1 * s1 + s2 * s2));
      }
    }
    if (q.ellipsometry) {
      // an exactly zero numerator gives +0 + 0i like numpy's division (Delta = 0, not +-180 from a -0)
      const C zero = mk<double>(0.0, 0.0);
      const C r0 = (norm2(pp) == 0.0 && norm2(ss) > 0.0) ? zero : pp / ss;
      const C r1 = (norm2(ps) == 0.0 && norm2(ss) > 0.0) ? zero : ps / ss;
      const C r2 = (norm2(sp) == 0.0 && norm2(pp) > 0.0) ? zero : sp / pp;
      q.ellipsometry[i] = deg(atan(cabs(r0)));
      q.ellipsometry[n + i] = deg(atan2(r0.im, r0.re));
      q.ellipsometry[2 * n + i] = deg(atan(cabs(r1)));
      q.ellipsometry[3 * n + i] = deg(atan2(r1.im, r1.re));
      q.ellipsometry[4 * n + i] = deg(atan(cabs(r2)));
      q.ellipsometry[5 * n + i] = deg(atan2(r2.im, r2.re));
    }
    if (q.eigenvalues || q.eigenvectors || q.discriminant || q.overlap) {
      const C h = 0.5 * (pp + ss), g = 0.5 * (pp - ss);
      const C disc = fma(ps, sp, g * g);
      const C sq = csqrt(disc);
      const C l0 = h + sq, l1 = h - sq;
      if (q.discriminant) st_plain(q.discriminant + i, disc);
      if (q.eigenvalues) {
        st_plain(q.eigenvalues + i, l0);
        st_plain(q.eigenvalues + n + i, l1);
      }
      if (q.eigenvectors || q.overlap) {
        // eigenvector of lambda: the larger of the two columns of adj(J - lambda I)
        C v[2][2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const C lam = k == 0 ? l0 : l1;
          C ax = ps, ay = lam - pp;       // (J - lam) annihilates (ps, lam - pp)
          C bx = lam - ss, by = sp;       // ... and (lam - ss, sp)
          const bool first = norm2(ax) + norm2(ay) >= norm2(bx) + norm2(by);
          C x = first ? ax : bx, y = first ? ay : by;
          // J = lambda I up to rounding (|J - lambda I| < 1e-15 |J|): any basis; numpy returns the identity
          if (!(norm2(x) + norm2(y) > 1e-30 * (norm2(pp) + norm2(ps) + norm2(sp) + norm2(ss)))) {
            x = mk<double>(k == 0 ? 1.0 : 0.0, 0.0);
            y = mk<double>(k == 0 ? 0.0 : 1.0, 0.0);
          }
          unit_phase2(x, y);
          v[k][0] = x; v[k][1] = y;
        }
        if (q.eigenvectors) {
          st_plain(q.eigenvectors + i, v[0][0]);          // (row 0, col 0)
          st_plain(q.eigenvectors + n + i, v[1][0]);      // (row 0, col 1)
          st_plain(q.eigenvectors + 2 * n + i, v[0][1]);  // (row 1, col 0)
          st_plain(q.eigenvectors + 3 * n + i, v[1][1]);  // (row 1, col 1)
        }
        if (q.overlap) q.overlap[i] = cabs(cmulc(v[0][0], v[1][0]) + cmulc(v[0][1], v[1][1]));
      }
    }
  }
}

// eps(f) of a built-in dispersion model, rotated by Q = Ry*Rx, packed symmetric
__global__ void __launch_bounds__(128)
material_table_kernel(const __grid_constant__ ho_material_t m, const double* __restrict__ freq, int64_t F, ho_c128* __restrict__ out) {
  using namespace ho;
  typedef cx<double> C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < F; i += (int64_t)gridDim.x * blockDim.x) {
    const double f = freq[i], f2 = f * f;
    C e[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) e[r][c] = mk<double>(0.0, 0.0);
    if (m.kind == HO_MAT_TOLO) {
      // eps_inf * prod_k (w_lo^2 - f^2 - i f g_lo) / (w_to^2 - f^2 - i f g_to)   (materials.py:176-213)
#pragma unroll
      for (int ax = 0; ax < 3; ++ax) {
        C prod = mk<double>(1.0, 0.0);
        for (int k = 0; k < m.n_osc[ax]; ++k) {
          const C num = mk<double>(m.w_lo[ax][k] * m.w_lo[ax][k] - f2, -f * m.g_lo[ax][k]);
          const C den = mk<double>(m.w_to[ax][k] * m.w_to[ax][k] - f2, -f * m.g_to[ax][k]);
          prod = prod * (num / den);
        }
        e[ax][ax] = m.eps_inf[ax] * prod;
      }
    } else {
      // monoclinic: Bu modes in the x-y plane at angles alpha, Au modes along z (materials.py:614-651)
      C xx = mk<double>(m.eps_inf[0], 0.0), yy = mk<double>(m.eps_inf[1], 0.0), xy = mk<double>(m.eps_inf[2], 0.0);
      C zz = mk<double>(m.eps_inf[3], 0.0);
      for (int k = 0; k < m.n_osc[0]; ++k) {
        const C rho = (m.amp[0][k] * m.amp[0][k]) * recip(mk<double>(m.w_to[0][k] * m.w_to[0][k] - f2, -f * m.g_to[0][k]));
        double sa, ca;
        sincos(m.alpha[k], &sa, &ca);
        xx = xx + (ca * ca) * rho; xy = xy + (sa * ca) * rho; yy = yy + (sa * sa) * rho;
      }
      for (int k = 0; k < m.n_osc[1]; ++k)
        zz = zz + (m.amp[1][k] * m.amp[1][k]) * recip(mk<double>(m.w_to[1][k] * m.w_to[1][k] - f2, -f * m.g_to[1][k]));
      e[0][0] = xx; e[0][1] = xy; e[1][0] = xy; e[1][1] = yy; e[2][2] = zz;
    }
    // (Q e) Q^T, Q real
    C t[3][3], r[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) t[a][b] = m.rot[3 * a] * e[0][b] + m.rot[3 * a + 1] * e[1][b] + m.rot[3 * a + 2] * e[2][b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) r[a][b] = t[a][0] * m.rot[3 * b] + t[a][1] * m.rot[3 * b + 1] + t[a][2] * m.rot[3 * b + 2];
    ho_c128* o = out + 6 * i;
    st_plain(o + 0, r[0][0]); st_plain(o + 1, r[0][1]); st_plain(o + 2, r[0][2]);
    st_plain(o + 3, r[1][1]); st_plain(o + 4, r[1][2]); st_plain(o + 5, r[2][2]);
  }
}

}  // namespace
