// ho_modes.cuh -- opt-in materialising kernels.
//
// The reflection kernels keep a 4x2 plane in registers and never form eigenvectors.  The
// reference, however, exposes per-layer objects that its own consumers read:
//   layer.matrix, layer.profile (WaveProfile)        waves.py:35-131, 298-335, 512-575
//   structure.transfer_matrix                        structure.py:259-270
//   -> FieldProfile (fields.py:78-97, 156-169), scattering_coefficients (scattering.py:45-63)
// These two kernels write exactly those arrays, in the reference's canonical [A, B, F, T, ...]
// layout, on request (HBM-bound: 256 B per point and layer):
//   modes_kernel : one Wave-built layer -> 4x4 layer matrix, the four partial waves in LAPACK's
//                  normalisation (unit 2-norm, largest component real positive, waves.py:270)
//                  ordered like wave_sorting + sort_poynting_indices (waves.py:256-296, 474-510),
//                  kz, Ez/Hz (waves.py:382-408) and the time-averaged Poynting vectors.
//   gamma_kernel : Gamma = M_prism M_1 ... M_exit with the exit columns in that same mode basis.
// Included from ho_engine.cu after apply_layer / layer_delta.
#pragma once

namespace {

struct ModesOut {
  ho::cx<double>* matrix;   // [nm][4][4] row-major
  ho::cx<double>* modes;    // [np][4][4] rows Ex, Ey, Hx, Hy; columns t0, t1, r0, r1
  ho::cx<double>* kz;       // [np][4]
  ho::cx<double>* fields;   // [np][6][4] rows Ex, Ey, Ez, Hx, Hy, Hz
  double* poynting;         // [np][3][4] rows Px, Py, Pz (time-averaged)
  int mA, mB, mF, mT;       // batch shape of the matrix (1 = the layer does not depend on that axis)
  int pA, pB, pF;           // batch shape of the profile (never depends on T; on F only if dispersive)
};

struct PartialWave {
  ho::Vec4<double> v;       // Ex, Ey, Hx, Hy
  ho::cx<double> lam, ez, hz;
  double cp, ce;            // p-likeness by complex Poynting / by E (waves.py:490-497)
  double px, py, pz;        // time-averaged Poynting vector
};

__device__ __forceinline__ ho::Vec4<double> unit_vec(int k) {
  using namespace ho;
  const cx<double> one = mk(1.0, 0.0), zero = mk(0.0, 0.0);
  Vec4<double> e;
  e.v0 = k == 0 ? one : zero; e.v1 = k == 1 ? one : zero; e.v2 = k == 2 ? one : zero; e.v3 = k == 3 ? one : zero;
  return e;
}

// Longitudinal fields, Poynting vectors and the two p-likeness measures of one normalised mode
template <bool NM>
__device__ __forceinline__ void finish_wave(PartialWave& m, const ho::ZRow<double>& z, double k) {
  using namespace ho;
  const cx<double> ex = m.v.v0, ey = m.v.v1, hx = m.v.v2, hy = m.v.v3;
  m.ez = -((k * hy + z.e20 * ex + z.e21 * ey) * z.ie22);
  m.hz = (k * ey - z.m20 * hx - z.m21 * hy) * z.im22;
  const double cpx = norm2(ey * m.hz - m.ez * hy), cpy = norm2(m.ez * hx - ex * m.hz);
  const double nex = norm2(ex), ney = norm2(ey);
  m.cp = cpx / (cpx + cpy);
  m.ce = nex / (nex + ney);
  m.px = 0.5 * (cmulc(m.hz, ey).re - cmulc(hy, m.ez).re);   // Re(Ey Hz* - Ez Hy*)
  m.py = 0.5 * (cmulc(hx, m.ez).re - cmulc(m.hz, ex).re);   // Re(Ez Hx* - Ex Hz*)
  m.pz = 0.5 * (cmulc(hy, ex).re - cmulc(hx, ey).re);       // Re(Ex Hy* - Ey Hx*)
}

// Eigenvector of root `lam` of D: any non-vanishing column of (D - a)(D - b)(D - c), the other
// three roots -- the largest of the four is taken.
template <bool NM>
__device__ __forceinline__ ho::Vec4<double> eigenvector(const ho::Delta<double, NM>& D, ho::cx<double> a, ho::cx<double> b, ho::cx<double> c) {
  using namespace ho;
  Vec4<double> best = unit_vec(0);
  double nbest = -1.0;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    Vec4<double> w = unit_vec(k);
    w = axpy(-c, w, matvec(D, w));
    w = axpy(-b, w, matvec(D, w));
    w = axpy(-a, w, matvec(D, w));
    const double n = norm2(w);
    if (n > nbest) { nbest = n; best = w; }
  }
  return best;
}

// Two independent vectors of the 2-d invariant subspace complementary to the roots (a, b): the
// largest column of (D - a)(D - b) and the column with the largest part orthogonal to it.
template <bool NM>
__device__ __forceinline__ void invariant_plane_basis(const ho::Delta<double, NM>& D, ho::cx<double> a, ho::cx<double> b,
                                                      ho::Vec4<double>& v0, ho::Vec4<double>& v1) {
  using namespace ho;
  double nbest = -1.0;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    Vec4<double> q = unit_vec(k);
    q = axpy(-b, q, matvec(D, q));
    q = axpy(-a, q, matvec(D, q));
    const double n = norm2(q);
    if (n > nbest) { nbest = n; v0 = q; }
  }
  const double inv = nbest > 0.0 ? 1.0 / nbest : 0.0;
  double rbest = -1.0;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    Vec4<double> q = unit_vec(k);
    q = axpy(-b, q, matvec(D, q));
    q = axpy(-a, q, matvec(D, q));
    q = axpy(-(inv * dotc(v0, q)), v0, q);
    const double n = norm2(q);
    if (n > rbest) { rbest = n; v1 = q; }
  }
}

// Incoming order of a root pair (wave_sorting, waves.py:273-289): Im descending when the
// eigenvalues are complex (|Im| > 1e-9), else Re descending.  Returns true when l1 comes first.
__device__ __forceinline__ bool second_first(ho::cx<double> l0, ho::cx<double> l1) {
  const bool cplx = fabs(l0.im) > 1e-9 && fabs(l1.im) > 1e-9;
  return cplx ? (l1.im > l0.im) : (l1.re > l0.re);
}

// sort_poynting_indices (waves.py:474-510) on a pair given in incoming order: by C_P descending
// when the two differ by more than 1e-6, else by C_E ascending (stable).  True = swap.
__device__ __forceinline__ bool swap_by_character(const PartialWave& m0, const PartialWave& m1) {
  if (fabs(m1.cp - m0.cp) > 1e-6) return m1.cp > m0.cp;
  return m1.ce < m0.ce;
}

// One layer of the stack: matrix + partial waves.  Thread <-> point of the matrix grid, canonical
// order n = ((a*mB + b)*mF + f)*mT + t.
template <bool NM>
__global__ void __launch_bounds__(128, 2)
modes_kernel(const __grid_constant__ ho_problem_t P, int layer, ModesOut out, int64_t n_points) {
  using namespace ho;
  typedef double T;
  const ho_layer_t& L = P.layers[layer];
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_points; n += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = n;
    const int t = (int)(q % out.mT); q /= out.mT;
    const int f = (int)(q % out.mF); q /= out.mF;
    const int b = (int)(q % out.mB);
    const int a = (int)(q / out.mB);
    const T k = __ldg(P.kx + a), k0 = __ldg(P.k0 + f);
    const bool crystal = L.kind == HO_ANISO_FILM || L.kind == HO_ANISO_EXIT;
    const bool iso = L.kind == HO_ISO_FILM || (crystal && L.eps_scalar && L.mu_scalar);
    const bool finite = L.kind == HO_ISO_FILM || L.kind == HO_ANISO_FILM;

    ZRow<T> z;
    Delta<T, NM> D;
    cx<T> eps = mk(T(L.iso_eps.re), T(L.iso_eps.im)), mu = mk(T(L.iso_mu.re), T(L.iso_mu.im));
    PartialWave w[4];   // incoming order: forward pair, backward pair
    Spectrum<T> s;
    if (iso) {
      if (crystal) {
        eps = ldc<T>(L.eps + 6 * (int64_t)(L.n_eps > 1 ? f : 0));
        mu = ldc<T>(L.mu + 6 * (int64_t)(L.n_mu > 1 ? f : 0));
      }
      // scalar medium: p modes (d, 0, 0, +-kz), s modes (0, -mu, +-kz, 0)
      const cx<T> zero = mk(0.0, 0.0);
      const T k2 = k * k;
      const cx<T> d = mu - k2 * recip(eps);
      const cx<T> kz = forward_sqrt(eps * mu - k2);
      z.e20 = zero; z.e21 = zero; z.ie22 = recip(eps); z.m20 = zero; z.m21 = zero; z.im22 = recip(mu);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const cx<T> lam = i < 2 ? kz : -kz;
        w[i].lam = lam;
        if ((i & 1) == 0) { w[i].v.v0 = d; w[i].v.v1 = zero; w[i].v.v2 = zero; w[i].v.v3 = lam; }
        else              { w[i].v.v0 = zero; w[i].v.v1 = -mu; w[i].v.v2 = lam; w[i].v.v3 = zero; }
      }
    } else {
      D = layer_delta<T, NM>(L, k, f, b, z);
      cx<T> c3, c2, c1, c0;
      charpoly(D, c3, c2, c1, c0);
      bool bail = false;
      s = split_spectrum<T, false>(c3, c2, c1, c0, bail);
      cx<T> l[4];
      quad_roots(-s.sf, s.pf, l[0], l[1]);
      quad_roots(-s.sb, s.pb, l[2], l[3]);
      if (second_first(l[0], l[1])) { cx<T> x = l[0]; l[0] = l[1]; l[1] = x; }
      if (second_first(l[2], l[3])) { cx<T> x = l[2]; l[2] = l[3]; l[3] = x; }
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        w[i].lam = l[i];
        w[i].v = eigenvector<NM>(D, l[(i + 1) & 3], l[(i + 2) & 3], l[(i + 3) & 3]);
      }
      // (near-)degenerate pair, e.g. normal incidence on a c-cut uniaxial film: the two products
      // above coincide; any basis of the 2-d invariant subspace is as good as LAPACK's
#pragma unroll 1
      for (int pair = 0; pair < 2; ++pair) {
        const int i0 = 2 * pair, o0 = 2 - 2 * pair;
        if (norm2(l[i0] - l[i0 + 1]) <= 1e-14 * (norm2(l[i0]) + norm2(l[i0 + 1])))
          invariant_plane_basis<NM>(D, l[o0], l[o0 + 1], w[i0].v, w[i0 + 1].v);
      }
    }
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      w[i].v = lapack_normalise(w[i].v);
      finish_wave<NM>(w[i], z, k);
    }
    if (swap_by_character(w[0], w[1])) { PartialWave x = w[0]; w[0] = w[1]; w[1] = x; }
    if (swap_by_character(w[2], w[3])) { PartialWave x = w[2]; w[2] = w[3]; w[3] = x; }

    // ---- layer matrix ----
    if (out.matrix) {
      Vec4<T> col[4];
      if (!finite) {
        // semi-infinite crystal: [t0, 0, t1, 0]  (waves.py:525-534)
        const cx<T> zero = mk(0.0, 0.0);
        Vec4<T> nul; nul.v0 = zero; nul.v1 = zero; nul.v2 = zero; nul.v3 = zero;
        col[0] = w[0].v; col[1] = nul; col[2] = w[1].v; col[3] = nul;
      } else {
        // exp(-i D k0 d) = V diag(exp(-i kz k0 d)) V^-1  (waves.py:298-335), column by column
        const T dd = __ldg(L.thickness + (L.n_thickness > 1 ? t : 0));
        const cx<T> c = mk(0.0, -(k0 * dd));
#pragma unroll 1
        for (int pair = 0; pair < 2; ++pair) {
          Vec4<T> y0 = unit_vec(2 * pair), y1 = unit_vec(2 * pair + 1);
          cx<T> wm[4];
          if (iso) {
            iso_film(k, eps, mu, c, false, y0, y1, wm);
          } else {
            film_transfer<T, NM>(D, s, c, y0, y1, wm);
            // undo the re-basing of a separated forward pair: the matrix wants exp(cD) e_k itself
            if (wm[1].re != 0.0 || wm[1].im != 0.0) y1 = axpy(-wm[1], y0, y1);
            if (wm[2].re != 0.0 || wm[2].im != 0.0) y0 = axpy(-wm[2], y1, y0);
          }
          col[2 * pair] = y0; col[2 * pair + 1] = y1;
        }
      }
      cx<T>* m = out.matrix + 16 * n;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        m[0 + c] = col[c].v0; m[4 + c] = col[c].v1; m[8 + c] = col[c].v2; m[12 + c] = col[c].v3;
      }
    }
    // ---- profile (independent of t; of f only through a dispersive tensor) ----
    if (t == 0 && (out.pF > 1 || f == 0)) {
      const int64_t np = ((int64_t)a * out.pB + b) * out.pF + (out.pF > 1 ? f : 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (out.modes) {
          cx<T>* v = out.modes + 16 * np;
          v[0 + i] = w[i].v.v0; v[4 + i] = w[i].v.v1; v[8 + i] = w[i].v.v2; v[12 + i] = w[i].v.v3;
        }
        if (out.kz) out.kz[4 * np + i] = w[i].lam;
        if (out.fields) {
          cx<T>* fl = out.fields + 24 * np;
          fl[0 + i] = w[i].v.v0; fl[4 + i] = w[i].v.v1; fl[8 + i] = w[i].ez;
          fl[12 + i] = w[i].v.v2; fl[16 + i] = w[i].v.v3; fl[20 + i] = w[i].hz;
        }
        if (out.poynting) {
          double* pv = out.poynting + 12 * np;
          pv[0 + i] = w[i].px; pv[4 + i] = w[i].py; pv[8 + i] = w[i].pz;
        }
      }
    }
  }
}

// Gamma = M_prism M_1 ... M_exit (structure.py:259-270), transfer semantics (overflows like the
// reference), canonical order n = ((a*B + b)*F + f)*T + t.  Built from the same running plane as
// the reflection kernel: columns (0, 2) are M_prism E [t0 t1] with E the film product and
// [t0 t1] = Y B the exit modes in the reference's basis (exit_mode_basis); the plane is tracked in
// the kernel's own basis (with the re-basing maps W of separated films), so
//   Gamma[:, (0, 2)] = M_prism Y_top (B^-1 W_{L-2} ... W_0)^-1.
// A crystal exit has zero columns 1, 3 (waves.py:525-534); an isotropic exit's backward columns
// are the forward ones with cos(theta_f) -> -cos(theta_f) (layers.py:190-238).
template <bool NM>
__global__ void __launch_bounds__(128, 2)
gamma_kernel(const __grid_constant__ ho_problem_t P, ho::cx<double>* gamma, int64_t n_begin, int64_t n_end) {
  using namespace ho;
  typedef double T;
  const int last = P.n_layers - 1;
  const bool iso_exit = P.layers[last].kind == HO_ISO_EXIT;
  for (int64_t n = n_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_end; n += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = n;
    const int t = (int)(q % P.T); q /= P.T;
    const int f = (int)(q % P.F); q /= P.F;
    const int b = (int)(q % P.B);
    const int a = (int)(q / P.B);
    const T k = __ldg(P.kx + a), k0 = __ldg(P.k0 + f);
    const T inv_c = __ldg(P.prism_inv_cos + a), inv_nc = __ldg(P.prism_inv_ncos + a), inv_n = P.prism_inv_n;
    cx<T>* g = gamma + 16 * (n - n_begin);
    const cx<T> zero = mk(0.0, 0.0);
#pragma unroll 1
    for (int plane = 0; plane < 2; ++plane) {
      if (plane == 1 && !iso_exit) {
#pragma unroll
        for (int r = 0; r < 4; ++r) { g[4 * r + 1] = zero; g[4 * r + 3] = zero; }
        break;
      }
      Vec4<T> y0, y1;
      cx<T> binv[4];
      cx<T> m00 = mk(1.0, 0.0), m01 = zero, m10 = zero, m11 = mk(1.0, 0.0);   // W_{L-2} ... W_l so far
#pragma unroll 1
      for (int l = last; l >= 0; --l) {
        cx<T> w[4];
        w[0] = mk(1.0, 0.0); w[1] = zero; w[2] = zero; w[3] = w[0];
        bool bail = false;
        if (l == last && plane == 1) {
          exit_iso_plane(-ldc<T>(P.layers[l].exit_cos + a), T(P.layers[l].exit_n), y0, y1);
          binv[0] = mk(1.0, 0.0); binv[1] = zero; binv[2] = zero; binv[3] = binv[0];
        } else {
          apply_layer<T, false, NM, true, false>(P.layers[l], k, k0, f, a, b, t, y0, y1, w, binv, bail, false);
        }
        if (l < last) {
          const cx<T> a00 = fma(m01, w[2], m00 * w[0]), a01 = fma(m01, w[3], m00 * w[1]);
          const cx<T> a10 = fma(m11, w[2], m10 * w[0]), a11 = fma(m11, w[3], m10 * w[1]);
          m00 = a00; m01 = a01; m10 = a10; m11 = a11;
        }
      }
      // C = (B^-1 M)^-1
      const cx<T> e00 = fma(binv[1], m10, binv[0] * m00), e01 = fma(binv[1], m11, binv[0] * m01);
      const cx<T> e10 = fma(binv[3], m10, binv[2] * m00), e11 = fma(binv[3], m11, binv[2] * m01);
      cx<T> c00, c01, c10, c11;
      inv2(e00, e01, e10, e11, c00, c01, c10, c11);
      const Vec4<T> u0 = axpy(c10, y1, scale(c00, y0)), u1 = axpy(c11, y1, scale(c01, y0));
      const int j0 = plane == 0 ? 0 : 1, j1 = plane == 0 ? 2 : 3;
      // prism rows (layers.py:109-152)
      g[0 + j0] = 0.5 * (u0.v1 - inv_nc * u0.v2);  g[0 + j1] = 0.5 * (u1.v1 - inv_nc * u1.v2);
      g[4 + j0] = 0.5 * (u0.v1 + inv_nc * u0.v2);  g[4 + j1] = 0.5 * (u1.v1 + inv_nc * u1.v2);
      g[8 + j0] = 0.5 * (inv_c * u0.v0 + inv_n * u0.v3);  g[8 + j1] = 0.5 * (inv_c * u1.v0 + inv_n * u1.v3);
      g[12 + j0] = 0.5 * (inv_n * u0.v3 - inv_c * u0.v0); g[12 + j1] = 0.5 * (inv_n * u1.v3 - inv_c * u1.v0);
    }
  }
}

}  // namespace
