"""NumPy model of the DEVICE algorithm (development + CPU-side algorithm tests).

This is a vectorised, statement-for-statement model of what the CUDA kernels in
``hyperbolic_optics_b200/csrc`` compute per point.  It exists so the eigenvector-free
algorithm can be checked against the oracle on full grids in
the CPU container.  It is NOT the oracle (that is ``oracle/reference_port.py``, a
restatement of the reference's eig-based algorithm) and NOT a product path: nothing in
``hyperbolic_optics_b200`` imports it.

Per anisotropic layer (symmetric eps, mu rotated about z by beta):
  1. Berreman matrix in its 9-parameter structured form
        [[a, b, c, d], [0, e, f, -c], [g, h, e, -b], [j, -g, 0, a]]
  2. characteristic quartic, Ferrari/Cardano roots, forward/backward classification,
     Bairstow polish of the forward quadratic factor q_f; q_b by deflation
  3. spectral projector P_f = q_b(D) h(D), h linear with h q_b = 1 (mod q_f)
  4. substrate: forward plane = P_f [e0 e1];  film: exp(cD) Y = G_f P_f Y + G_b P_b Y with
     2x2-block exponentials written in (mean, half-difference^2) of each root pair.
"""

from __future__ import annotations

import numpy as np

IM_TOL = 1e-9  # reference waves.py:273


# ----------------------------------------------------------------------------- tensors
def rotate_z(t, cb, sb):
    """Rz(beta) T Rz(beta)^T for symmetric T = (xx, xy, xz, yy, yz, zz)."""
    xx, xy, xz, yy, yz, zz = t
    c2, s2, cs = cb * cb, sb * sb, cb * sb
    return (c2 * xx - 2.0 * cs * xy + s2 * yy,
            cs * (xx - yy) + (c2 - s2) * xy,
            cb * xz - sb * yz,
            s2 * xx + 2.0 * cs * xy + c2 * yy,
            sb * xz + cb * yz,
            zz)


def berreman9(k, e, m):
    """Structured Berreman entries (a,b,c,d,e,f,g,h,j) -- reference waves.py:221-245 with
    symmetric tensors."""
    exx, exy, exz, eyy, eyz, ezz = e
    mxx, mxy, mxz, myy, myz, mzz = m
    ie, im = 1.0 / ezz, 1.0 / mzz
    k2 = k * k
    a = -k * exz * ie
    b = k * (myz * im - eyz * ie)
    c = mxy - myz * mxz * im
    d = myy - myz * myz * im - k2 * ie
    ee = -k * mxz * im
    f = mxz * mxz * im - mxx
    g = eyz * exz * ie - exy
    h = k2 * im - eyy + eyz * eyz * ie
    j = exx - exz * exz * ie
    return a, b, c, d, ee, f, g, h, j


def matvec(D, v):
    """Structured D @ v for v = (v0, v1, v2, v3)."""
    a, b, c, d, e, f, g, h, j = D
    v0, v1, v2, v3 = v
    return (a * v0 + b * v1 + c * v2 + d * v3,
            e * v1 + f * v2 - c * v3,
            g * v0 + h * v1 + e * v2 - b * v3,
            j * v0 - g * v1 + a * v3)


def charpoly(D):
    """lambda^4 + c3 lambda^3 + c2 lambda^2 + c1 lambda + c0 = det(lambda I - D)."""
    a, b, c, d, e, f, g, h, j = D
    a2, e2, cg, dj, fh = a * a, e * e, c * g, d * j, f * h
    bcj, bf = b * c * j, b * f
    c3 = -2.0 * (a + e)
    c2 = a2 + 4.0 * a * e - 2.0 * cg - dj + e2 - fh
    c1 = -2.0 * (a2 * e - a * cg + a * e2 - a * fh - bcj + bf * g - e * cg - e * dj)
    c0 = (a2 * e2 - a2 * fh + 2.0 * a * bf * g - 2.0 * a * e * cg + b * bf * j - 2.0 * e * bcj
          + cg * cg + c * c * h * j - e2 * dj + d * f * g * g + dj * fh)
    return c3, c2, c1, c0


# ----------------------------------------------------------------------------- quartic
_OMEGA = complex(np.exp(2j * np.pi / 3.0))  # python scalar: keeps complex64 inputs complex64


def _ccbrt(z):
    r = np.abs(z)
    th = np.angle(z)
    return np.cbrt(r) * np.exp(1j * th / 3.0)


def _quad_roots(bq, cq):
    """Roots of y^2 + bq y + cq (cancellation-free).  A pair closer than ~sqrt(tau) (relative to
    |bq|) is collapsed onto its mean -bq/2: the split of a near-double root is rounding noise
    (~sqrt(eps)), and letting it through can put the two copies on opposite sides of the
    forward/backward classification when the loss is tiny."""
    tau = 1e-12 if np.asarray(bq).real.dtype == np.float64 else 1e-5
    sq = np.sqrt(bq * bq - 4.0 * cq)
    sq = np.where(np.abs(sq) ** 2 <= tau * np.abs(bq) ** 2, 0.0, sq)
    sgn = np.where((np.conj(bq) * sq).real >= 0.0, 1.0, -1.0).astype(sq.real.dtype)
    t = -0.5 * (bq + sgn * sq)
    safe = np.abs(t) > 0.0
    y2 = np.where(safe, cq / np.where(safe, t, 1.0), 0.0)
    return t, y2


def ferrari(c3, c2, c1, c0):
    """Closed-form roots (start values only; polished later)."""
    c3 = c3 + 0j  # dtype-preserving (complex64 stays complex64)
    q3 = 0.25 * c3
    p = c2 - 6.0 * q3 * q3
    q = c1 - 2.0 * q3 * c2 + 8.0 * q3 * q3 * q3
    r = c0 - q3 * c1 + q3 * q3 * c2 - 3.0 * q3 ** 4
    # resolvent m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0 ; m = t - p/3
    P = -p * p / 12.0 - r
    Q = -p * p * p / 108.0 + p * r / 3.0 - q * q / 8.0
    sq = np.sqrt(Q * Q / 4.0 + P * P * P / 27.0)
    u3a, u3b = -Q / 2.0 + sq, -Q / 2.0 - sq
    u3 = np.where(np.abs(u3a) >= np.abs(u3b), u3a, u3b)
    u = _ccbrt(u3)
    best_m, best_abs = None, None
    for kk in range(3):
        uk = u * _OMEGA ** kk
        ok = np.abs(uk) > 0.0
        tk = uk - np.where(ok, P / (3.0 * np.where(ok, uk, 1.0)), 0.0)
        mk = tk - p / 3.0
        ak = np.abs(mk)
        if best_m is None:
            best_m, best_abs = mk, ak
        else:
            take = ak > best_abs
            best_m = np.where(take, mk, best_m)
            best_abs = np.where(take, ak, best_abs)
    m = best_m
    alpha = np.sqrt(2.0 * m)
    ok = np.abs(alpha) > 0.0
    qa = np.where(ok, q / (2.0 * np.where(ok, alpha, 1.0)), 0.0)
    beta1 = 0.5 * p + m - qa
    beta2 = 0.5 * p + m + qa
    y0, y1 = _quad_roots(alpha, beta1)
    y2, y3 = _quad_roots(-alpha, beta2)
    return [y - q3 for y in (y0, y1, y2, y3)]


def newton_polish(roots, c3, c2, c1, c0, iters=2):
    out = []
    for lam in roots:
        for _ in range(iters):
            pv = (((lam + c3) * lam + c2) * lam + c1) * lam + c0
            dv = ((4.0 * lam + 3.0 * c3) * lam + 2.0 * c2) * lam + c1
            ok = np.abs(dv) > 1e-300
            step = np.where(ok, pv / np.where(ok, dv, 1.0), 0.0)
            # guard against wild steps at near-multiple roots
            big = np.abs(step) > 0.5 * (np.abs(lam) + 1e-300)
            lam = lam - np.where(big, 0.0, step)
        out.append(lam)
    return out


def classify_forward(roots):
    """Boolean mask per root: the two 'transmitted' modes (reference waves.py:273-289).

    all |Im| > tol : two largest Im;  none: two largest Re;  mixed: forward iff Im > tol or
    (|Im| <= tol and Re > 0), resolved to exactly two by ranking.
    """
    lam = np.stack(roots, axis=-1)
    im, re = lam.imag, lam.real
    cplx = np.abs(im) > IM_TOL
    n_c = cplx.sum(axis=-1, keepdims=True)
    score = np.where(cplx, np.where(im > 0, 2.0, -2.0), np.where(re > 0, 1.0, -1.0))
    key1 = np.where(n_c == 4, 0.0, np.where(n_c == 0, 0.0, score))
    key2 = np.where(n_c == 0, re, im)
    key3 = re
    # rank = number of roots strictly "better" (lexicographic, index as final tie-break)
    rank = np.zeros(lam.shape, dtype=np.int64)
    for i in range(4):
        for jx in range(4):
            if i == jx:
                continue
            better = (key1[..., jx] > key1[..., i]) | (
                (key1[..., jx] == key1[..., i]) & ((key2[..., jx] > key2[..., i]) | (
                    (key2[..., jx] == key2[..., i]) & ((key3[..., jx] > key3[..., i]) | (
                        (key3[..., jx] == key3[..., i]) & (jx < i))))))
            rank[..., i] += better
    return [rank[..., i] < 2 for i in range(4)], n_c[..., 0]


def bairstow(u, v, c3, c2, c1, c0, iters=3):
    """Newton on (u, v) so that lambda^2 + u lambda + v divides the quartic."""
    for _ in range(iters):
        b1 = c3 - u
        b0 = c2 - u * b1 - v
        r1 = c1 - u * b0 - v * b1
        r0 = c0 - v * b0
        j11 = -b0 - u * (u - b1) + v
        j12 = u - b1
        j21 = -v * (u - b1)
        j22 = v - b0
        det = j11 * j22 - j12 * j21
        ok = np.abs(det) > 1e-300
        idet = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        du = (-r1 * j22 + r0 * j12) * idet
        dv = (-r0 * j11 + r1 * j21) * idet
        u, v = u + du, v + dv
    b1 = c3 - u
    b0 = c2 - u * b1 - v
    return u, v, b1, b0


BIQ_TAU4 = 1e-20   # (1e-5)^4: relative size of the odd coefficients below which the quartic is
                   # treated as a perturbed biquadratic for the *start values*
STEP_TOL2 = 1e-22  # squared relative Newton step below which Bairstow has converged


def _forward_sqrt(y):
    """Square root of y oriented forward (reference waves.py:273-289 applied to the pair +-mu)."""
    mu = np.sqrt(y + 0j)
    fwd = np.where(np.abs(mu.imag) > IM_TOL, mu.imag > 0, mu.real > 0)
    return np.where(fwd, mu, -mu)


def biquadratic_seed(c2, c0):
    """Forward pair (sum, product) of lambda^4 + c2 lambda^2 + c0: lambda^2 = y_{1,2}."""
    y1, y2 = _quad_roots(c2, c0)
    m1, m2 = _forward_sqrt(y1), _forward_sqrt(y2)
    return m1 + m2, m1 * m2


RES_TOL2 = 1e-28   # squared relative residual (1e-14) at which the quadratic factor is accepted


def bairstow_adaptive(u, v, c3, c2, c1, c0, max_iters=3):
    """Bairstow with per-point early exits: a Newton step is skipped once the residual of the
    division is at rounding level, and the loop ends once a step is negligible.
    Returns (u, v, b1, b0, converged, newton_steps)."""
    n2 = lambda z: np.abs(z) ** 2
    done = np.zeros(np.shape(u), dtype=bool)
    its = np.zeros(np.shape(u), dtype=np.int64)
    for it in range(max_iters + 1):
        b1 = c3 - u
        ub0_ = c2 - v
        b0 = ub0_ - u * b1
        t1, t2, t3 = u * b0, v * b1, v * b0
        r1 = c1 - t1 - t2
        r0 = c0 - t3
        if it > 0:
            res_small = (n2(r1) <= RES_TOL2 * (n2(t1) + n2(t2))) & (n2(r0) <= RES_TOL2 * n2(t3))
            done = done | res_small
        if it == max_iters:
            break
        umb = u - b1
        j22 = v - b0
        j11 = j22 - u * umb
        j21 = -v * umb
        det = j11 * j22 - umb * j21
        ok = np.abs(det) > 0
        idet = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        du = (r0 * umb - r1 * j22) * idet
        dv = (r1 * j21 - r0 * j11) * idet
        u = np.where(done, u, u + du)
        v = np.where(done, v, v + dv)
        its += ~done
        small = (n2(du) + n2(dv)) <= STEP_TOL2 * (n2(u) + n2(v))
        done = done | small
    b1 = c3 - u
    b0 = c2 - u * b1 - v
    return u, v, b1, b0, done, its


def split_spectrum(D, polish=0, bairstow_iters=3, stats=None):
    """-> (sf, pf, sb, pb, n_complex): sum/product of forward and backward root pairs.

    Start values: when the odd coefficients are tiny (z is a principal axis of the rotated
    tensors up to the reference's hidden 1e-8 offsets: every un-tilted crystal) the quartic is a
    perturbed biquadratic and the forward pair comes from two square roots; otherwise Ferrari.
    Bairstow then polishes the forward quadratic factor with a per-point early exit; a
    biquadratic start that has not converged falls back to the Ferrari start."""
    c3, c2, c1, c0 = charpoly(D)
    n2 = lambda z: np.abs(z) ** 2
    biq = (n2(c3) ** 2 <= BIQ_TAU4 * n2(c2)) & (n2(c1) ** 2 <= BIQ_TAU4 * n2(c2) ** 3)
    sf_b, pf_b = biquadratic_seed(c2, c0)
    ub, vb, b1b, b0b, conv_b, its_b = bairstow_adaptive(-sf_b, pf_b, c3, c2, c1, c0, bairstow_iters)
    use_b = biq & conv_b
    roots = ferrari(c3, c2, c1, c0)
    if polish:
        roots = newton_polish(roots, c3, c2, c1, c0, polish)
    fwd, n_c = classify_forward(roots)
    sf = sum(np.where(fm, lam, 0.0) for fm, lam in zip(fwd, roots))
    pf = np.prod([np.where(fm, lam, 1.0) for fm, lam in zip(fwd, roots)], axis=0)
    u, v, b1, b0, conv_f, its_f = bairstow_adaptive(-sf, pf, c3, c2, c1, c0, bairstow_iters)
    if stats is not None:
        stats.setdefault("biq", []).append(float(np.mean(use_b)))
        stats.setdefault("biq_rejected", []).append(float(np.mean(biq & ~conv_b)))
        stats.setdefault("its_biq", []).append(np.bincount(its_b[use_b].ravel(), minlength=4) if use_b.any() else None)
        stats.setdefault("its_ferrari", []).append(np.bincount(its_f[~use_b].ravel(), minlength=4) if (~use_b).any() else None)
        stats.setdefault("unconverged_ferrari", []).append(float(np.mean(~conv_f & ~use_b)))
    u, v, b1, b0 = (np.where(use_b, x, y) for x, y in ((ub, u), (vb, v), (b1b, b1), (b0b, b0)))
    return -u, v, -b1, b0, n_c


def projector_coeffs(sf, pf, sb, pb):
    """h(x) = h0 + h1 x with h q_b = 1 (mod q_f)."""
    U = sf - sb
    V = pb - pf
    R = U * U * pf + U * V * sf + V * V
    iR = 1.0 / R
    return (U * sf + V) * iR, -U * iR


def apply_q(D, w, s, p):
    """(D^2 - s D + p I) w, also returns D w."""
    t1 = matvec(D, w)
    t2 = matvec(D, t1)
    return tuple(t2[i] - s * t1[i] + p * w[i] for i in range(4)), t1


def project_forward(D, y, sf, pf, sb, pb):
    """P_f y  (y a 4-tuple), plus D y (reused by callers)."""
    h0, h1 = projector_coeffs(sf, pf, sb, pb)
    dy = matvec(D, y)
    w = tuple(h0 * y[i] + h1 * dy[i] for i in range(4))
    z, _ = apply_q(D, w, sb, pb)
    return z, dy


def pair_exponential(s, p, cfac):
    """(A, B) with exp(cfac * K) = A I + B (K - m I), K having root pair (sum s, product p)."""
    m = 0.5 * s
    h2 = m * m - p
    h = np.sqrt(h2)
    z = cfac * h
    small = np.abs(z) < 0.25
    # general branch from the two individual exponentials (no 0*inf hazards)
    e_p = np.exp(cfac * (m + h))
    e_m = np.exp(cfac * (m - h))
    A_g = 0.5 * (e_p + e_m)
    hs = np.where(small, 1.0, h)
    B_g = 0.5 * (e_p - e_m) / hs
    # confluent branch: series in z^2
    z2 = z * z
    ch = 1.0 + z2 / 2.0 * (1.0 + z2 / 12.0 * (1.0 + z2 / 30.0 * (1.0 + z2 / 56.0 * (1.0 + z2 / 90.0))))
    sh = 1.0 + z2 / 6.0 * (1.0 + z2 / 20.0 * (1.0 + z2 / 42.0 * (1.0 + z2 / 72.0 * (1.0 + z2 / 110.0))))
    em = np.exp(cfac * m)
    A = np.where(small, em * ch, A_g)
    B = np.where(small, em * cfac * sh, B_g)
    return A, B, m


def _expm1c(x):
    """(exp(x) - 1) / x with a series for small |x|."""
    small = np.abs(x) < 0.5
    xs = np.where(small, x, 1.0)
    ser = 1.0 + xs / 2.0 * (1.0 + xs / 3.0 * (1.0 + xs / 4.0 * (1.0 + xs / 5.0 * (1.0 + xs / 6.0 * (
        1.0 + xs / 7.0 * (1.0 + xs / 8.0 * (1.0 + xs / 9.0 * (1.0 + xs / 10.0 * (1.0 + xs / 11.0 * (
            1.0 + xs / 12.0 * (1.0 + xs / 13.0 * (1.0 + xs / 14.0))))))))))))
    xl = np.where(small, 1.0, x)
    return np.where(small, ser, (np.exp(xl) - 1.0) / xl)


def pair_modes(s, p, cfac):
    """Order the root pair by |exp(cfac*lambda)|: returns (lam_small, phi_small, dd, separated)
    with exp(cfac K) = phi_small I + dd (K - lam_small I); ``separated`` = the two exponentials
    differ by more than a factor e."""
    m = 0.5 * s
    h = np.sqrt(m * m - p)
    swap = (cfac * h).real < 0.0
    h = np.where(swap, -h, h)          # now Re(cfac*h) >= 0: lam_big = m + h
    lam_s, lam_g = m - h, m + h
    phi_s, phi_g = np.exp(cfac * lam_s), np.exp(cfac * lam_g)
    x = 2.0 * cfac * h
    small = np.abs(x) < 0.5
    dd_direct = (phi_g - phi_s) / np.where(small, 1.0, lam_g - lam_s)
    dd = np.where(small, cfac * phi_s * _expm1c(np.where(small, x, 0.0)), dd_direct)
    return lam_s, phi_s, dd, x.real > 1.0


def _rank1_clean(r0, r1):
    """Project the (theoretically rank-one) 4x2 matrix [r0 r1] onto rank one: keep the
    larger column as the direction, replace the other by its component along it."""
    n0 = sum(np.abs(x) ** 2 for x in r0)
    n1 = sum(np.abs(x) ** 2 for x in r1)
    first = n0 >= n1
    v = tuple(np.where(first, a, b) for a, b in zip(r0, r1))
    w = tuple(np.where(first, b, a) for a, b in zip(r0, r1))
    nv = np.where(first, n0, n1)
    dot = sum(np.conj(a) * b for a, b in zip(v, w))
    beta = np.where(nv > 0, dot / np.where(nv > 0, nv, 1.0), 0.0)
    wclean = tuple(beta * a for a in v)
    c0 = tuple(np.where(first, a, b) for a, b in zip(v, wclean))
    c1 = tuple(np.where(first, b, a) for a, b in zip(v, wclean))
    return c0, c1


def block_exp_apply(D, Z, s, p, cfac, clean=True):
    """exp(cfac D) applied to the two columns Z (which lie in the invariant subspace whose
    root pair has sum s / product p):  phi_s Z + dd (D - lam_s) Z, the second term forced
    to rank one (it is a multiple of the faster-growing eigenvector)."""
    lam_s, phi_s, dd, separated = pair_modes(s, p, cfac)
    R = []
    for z in Z:
        dz = matvec(D, z)
        R.append(tuple(dz[i] - lam_s * z[i] for i in range(4)))
    if clean:
        # only when the pair's exponentials differ by more than e: then lam_s is a
        # well-resolved simple root and (D - lam_s) Z is rank one up to rounding
        C = _rank1_clean(R[0], R[1])
        R = [tuple(np.where(separated, C[k][i], R[k][i]) for i in range(4)) for k in range(2)]
    return [tuple(phi_s * Z[k][i] + dd * R[k][i] for i in range(4)) for k in range(2)]


MONO_GROWTH = 3.0  # thin-film switch: all four |Re(cfac lambda)| below this -> one cubic polynomial


def film_transfer(D, Y, spec, cfac, clean=True, mono=True, return_w=False):
    """exp(cfac D) Y.  Thin films: the monolithic cubic.  Otherwise the invariant-subspace block
    form; when the two forward modes grow at different rates ("separated") the plane is re-based by
    the column operation  other <- other - beta dom  (r_other = beta r_dom, rank one), whose dominant
    parts cancel analytically -- adding phi_s z to the huge dd r would drown the slower mode."""
    sf, pf, sb, pb = spec
    zf, zb = [], []
    for y in Y:
        z, _ = project_forward(D, y, sf, pf, sb, pb)
        zf.append(z)
        zb.append(tuple(y[i] - z[i] for i in range(4)))
    lam_s, phi_s, dd, sep = pair_modes(sf, pf, cfac)
    R = []
    for z in zf:
        dz = matvec(D, z)
        R.append(tuple(dz[i] - lam_s * z[i] for i in range(4)))
    gb = block_exp_apply(D, zb, sb, pb, cfac, clean)
    plain = [tuple(phi_s * zf[k][i] + dd * R[k][i] + gb[k][i] for i in range(4)) for k in range(2)]
    n0 = sum(np.abs(R[0][i]) ** 2 for i in range(4))
    n1 = sum(np.abs(R[1][i]) ** 2 for i in range(4))
    first = n0 >= n1
    pick = lambda a, b: [np.where(first, a[i], b[i]) for i in range(4)]
    v, w = pick(R[0], R[1]), pick(R[1], R[0])
    nv = np.where(first, n0, n1)
    beta = sum(np.conj(v[i]) * w[i] for i in range(4)) / np.where(nv > 0, nv, 1.0)
    zv, zw = pick(zf[0], zf[1]), pick(zf[1], zf[0])
    bv, bw = pick(gb[0], gb[1]), pick(gb[1], gb[0])
    dom = [phi_s * zv[i] + dd * v[i] + bv[i] for i in range(4)]
    oth = [phi_s * (zw[i] - beta * zv[i]) + (bw[i] - beta * bv[i]) for i in range(4)]
    rebased = [tuple(np.where(first, dom[i], oth[i]) for i in range(4)),
               tuple(np.where(first, oth[i], dom[i]) for i in range(4))]
    block = [tuple(np.where(sep, rebased[k][i], plain[k][i]) for i in range(4)) for k in range(2)]
    thin = (film_growth(spec, cfac) < MONO_GROWTH) if mono else np.zeros(np.shape(sep), dtype=bool)
    if mono:
        poly = film_transfer_mono(D, Y, spec, cfac)
        block = [tuple(np.where(thin, poly[k][i], block[k][i]) for i in range(4)) for k in range(2)]
    if not return_w:
        return block
    # coordinate map of the column operation (identity unless re-based): c_bottom = W c_top
    used = sep & ~thin
    one, zero = np.ones_like(beta), np.zeros_like(beta)
    w01 = np.where(used & first, -beta, zero)
    w10 = np.where(used & ~first, -beta, zero)
    return block, ((one, w01), (w10, one))


def _inv2(m00, m01, m10, m11):
    det = m00 * m11 - m01 * m10
    idet = 1.0 / det
    return m11 * idet, -m01 * idet, -m10 * idet, m00 * idet


def film_stable(D, Y, spec, cfac, rows=(0, 1), return_w=False, mono=True):
    """Stable propagation; thin films (MONO_GROWTH) use the cubic of the transfer recursion followed
    by a re-basing to unit E rows, thick films the decaying-exponential block form."""
    block, w_block = _film_stable_block(D, Y, spec, cfac, rows)
    if mono:
        raw = film_transfer_mono(D, Y, spec, cfac)
        r0, r1 = rows
        n00, n01, n10, n11 = _inv2(raw[0][r0], raw[1][r0], raw[0][r1], raw[1][r1])
        thin_y = [tuple(raw[0][i] * n00 + raw[1][i] * n10 for i in range(4)),
                  tuple(raw[0][i] * n01 + raw[1][i] * n11 for i in range(4))]
        thin_w = ((n00, n01), (n10, n11))
        thin = film_growth(spec, cfac) < MONO_GROWTH
        block = [tuple(np.where(thin, thin_y[k][i], block[k][i]) for i in range(4)) for k in range(2)]
        w_block = tuple(tuple(np.where(thin, thin_w[r][c], w_block[r][c]) for c in range(2)) for r in range(2))
    return (block, w_block) if return_w else block


def _film_stable_block(D, Y, spec, cfac, rows=(0, 1), return_w=True):
    """Stable (scattering-equivalent) propagation of the plane spanned by Y through a film.

    Returns the renormalised plane X_f + G_b Z_b N (L G_f^-1 X_f) using only decaying
    exponentials, and the 2x2 normalisation C = N g_f^-1 that maps
    top coordinates to bottom coordinates.
    """
    sf, pf, sb, pb = spec
    Afi, Bfi, mf = pair_exponential(sf, pf, -cfac)  # G_f^{-1} on F : exp(-cfac K_f), decaying
    Ab, Bb, mb = pair_exponential(sb, pb, cfac)     # G_b on B     : exp(+cfac K_b), decaying
    zf, zb = [], []
    for y in Y:
        z, _ = project_forward(D, y, sf, pf, sb, pb)
        zf.append(z)
        zb.append(tuple(y[i] - z[i] for i in range(4)))
    r0, r1 = rows
    n00, n01, n10, n11 = _inv2(zf[0][r0], zf[1][r0], zf[0][r1], zf[1][r1])
    xf = [tuple(zf[0][i] * n00 + zf[1][i] * n10 for i in range(4)),
          tuple(zf[0][i] * n01 + zf[1][i] * n11 for i in range(4))]
    xb = [tuple(zb[0][i] * n00 + zb[1][i] * n10 for i in range(4)),
          tuple(zb[0][i] * n01 + zb[1][i] * n11 for i in range(4))]
    # g = L G_f^-1 X_f   (2x2)
    gcols = []
    for x in xf:
        dx = matvec(D, x)
        gx = tuple((Afi - Bfi * mf) * x[i] + Bfi * dx[i] for i in range(4))
        gcols.append((gx[r0], gx[r1]))
    # G_b X_b
    gb = []
    for x in xb:
        dx = matvec(D, x)
        gb.append(tuple((Ab - Bb * mb) * x[i] + Bb * dx[i] for i in range(4)))
    out = [tuple(xf[cidx][i] + gb[0][i] * gcols[cidx][0] + gb[1][i] * gcols[cidx][1] for i in range(4))
           for cidx in range(2)]
    if return_w:
        # W = N G: coordinates at the top of the film -> coordinates at its bottom (Y_new = M Y_old W)
        w = ((n00 * gcols[0][0] + n01 * gcols[0][1], n00 * gcols[1][0] + n01 * gcols[1][1]),
             (n10 * gcols[0][0] + n11 * gcols[0][1], n10 * gcols[1][0] + n11 * gcols[1][1]))
        return out, w
    return out


def substrate_plane(D, spec):
    """Basis of the forward invariant subspace: q_b(D) [e0 e1] = (D^2 - s_b D + p_b I) [e0 e1].

    q_b(D) annihilates the backward subspace and is invertible on the forward one, so its columns
    span the same plane as P_f [e0 e1] (r_* only depend on the plane) at one
    mat-vec per column: D e_k is just column k of D."""
    a, b, c, d, e, f, g, h, j = D
    sf, pf, sb, pb = spec
    zero = np.zeros_like(sf)
    cols = []
    for idx, col in enumerate(((a, zero, g, j), (b, e, h, -g))):
        t1 = matvec(D, col)
        cols.append(tuple(t1[i] - sb * col[i] + (pb if i == idx else 0.0) for i in range(4)))
    return cols


# ----------------------------------------------------------------------------- isotropic
def iso_delta(k, eps, mu):
    """Isotropic Berreman entries in the 9-parameter form."""
    z = np.zeros_like(eps + k)
    k2 = k * k
    return (z, z, z, mu - k2 / eps, z, -mu + z, z, k2 / mu - eps, eps + z)


def iso_film_transfer(k, eps, mu, Y, cfac):
    """exp(cfac D) Y for an isotropic film: D^2 = kz^2 I  =>  cosh(c kz) I + c sinhc(c kz) D."""
    D = iso_delta(k, eps, mu)
    kz2 = eps * mu - k * k
    A, B, _ = pair_exponential(np.zeros_like(kz2), -kz2, cfac)  # pair (+kz, -kz): s=0, p=-kz^2
    out = []
    for y in Y:
        dy = matvec(D, y)
        out.append(tuple(A * y[i] + B * dy[i] for i in range(4)))
    return out


def iso_forward_kz(k, eps, mu):
    kz = np.sqrt(eps * mu - k * k + 0j)
    fwd = np.where(np.abs(kz.imag) > IM_TOL, kz.imag > 0, kz.real > 0)
    return np.where(fwd, kz, -kz)


def iso_film_stable(k, eps, mu, Y, cfac, rows=(0, 1), return_w=False):
    D = iso_delta(k, eps, mu)
    kz = iso_forward_kz(k, eps, mu)
    ikz = 1.0 / kz
    phase2 = np.exp(-2.0 * cfac * kz)  # exp(2 i tau kz), decaying (cfac = -i tau)
    zf, zb = [], []
    for y in Y:
        dy = matvec(D, y)
        f = tuple(0.5 * (y[i] + dy[i] * ikz) for i in range(4))
        zf.append(f)
        zb.append(tuple(y[i] - f[i] for i in range(4)))
    r0, r1 = rows
    n00, n01, n10, n11 = _inv2(zf[0][r0], zf[1][r0], zf[0][r1], zf[1][r1])
    out = []
    for (ca, cb) in ((n00, n10), (n01, n11)):
        out.append(tuple(zf[0][i] * ca + zf[1][i] * cb + phase2 * (zb[0][i] * ca + zb[1][i] * cb)
                         for i in range(4)))
    if return_w:
        ph = np.exp(-cfac * kz)  # exp(i tau kz), decaying
        return out, ((n00 * ph, n01 * ph), (n10 * ph, n11 * ph))
    return out


# ----------------------------------------------------------------------------- ambient + reduction
def exit_iso_plane(cf, n_exit):
    """Columns 0 and 2 of the isotropic exit matrix (reference layers.py:190-238)."""
    z = np.zeros_like(cf)
    o = np.ones_like(cf)
    return [(z, o, -n_exit * cf, z), (cf, z, z, n_exit * o)]


def reflect_from_plane(Y, inv_c, inv_nc, inv_n):
    """Prism matrix applied to the plane + 2x2 minors (reference layers.py:109-152,
    structure.py:285-304).  The common factor 1/2 of the prism matrix cancels."""
    g = []
    for y in Y:
        g.append((y[1] - y[2] * inv_nc, y[1] + y[2] * inv_nc,
                  y[0] * inv_c + y[3] * inv_n, -y[0] * inv_c + y[3] * inv_n))
    (g00, g10, g20, g30), (g02, g12, g22, g32) = g
    bottom = g00 * g22 - g02 * g20
    ib = 1.0 / bottom
    r_pp = (g00 * g32 - g30 * g02) * ib
    r_ps = (g00 * g12 - g10 * g02) * ib
    r_sp = (g30 * g22 - g32 * g20) * ib
    r_ss = (g10 * g22 - g12 * g20) * ib
    return r_pp, r_ps, r_sp, r_ss


def mueller(pp, ps, sp, ss):
    """Real Mueller matrix M = Re(A F A^-1) written out (reference mueller.py:266-307)."""
    cj = np.conj
    F = [[pp * cj(pp), pp * cj(ps), ps * cj(pp), ps * cj(ps)],
         [pp * cj(sp), pp * cj(ss), ps * cj(sp), ps * cj(ss)],
         [sp * cj(pp), sp * cj(ps), ss * cj(pp), ss * cj(ps)],
         [sp * cj(sp), sp * cj(ss), ss * cj(sp), ss * cj(ss)]]
    rows = [[F[0][c] + F[3][c] for c in range(4)], [F[0][c] - F[3][c] for c in range(4)],
            [F[1][c] + F[2][c] for c in range(4)], [1j * (F[1][c] - F[2][c]) for c in range(4)]]
    M = np.empty(np.shape(pp) + (4, 4))
    for r in range(4):
        c0, c1, c2, c3 = rows[r]
        M[..., r, 0] = 0.5 * (c0 + c3).real
        M[..., r, 1] = 0.5 * (c0 - c3).real
        M[..., r, 2] = 0.5 * (c1 + c2).real
        M[..., r, 3] = (0.5j * (c2 - c1)).real
    return M


# ----------------------------------------------------------------------------- power bookkeeping
def flux_form(Y):
    """Hermitian 2x2 form of S_z = 1/2 Re(Ex Hy* - Ey Hx*) on the plane Y c (reference
    fields.py:45-51): S_z = h00 |c0|^2 + h11 |c1|^2 + Re(c0 conj(c1) h01)."""
    def q(u, v):
        return u[0] * np.conj(v[3]) - u[1] * np.conj(v[2])
    y0, y1 = Y
    return 0.5 * q(y0, y0).real, 0.5 * q(y1, y1).real, 0.5 * (q(y0, y1) + np.conj(q(y1, y0)))


def flux_eval(h, c):
    h00, h11, h01 = h
    return h00 * np.abs(c[0]) ** 2 + h11 * np.abs(c[1]) ** 2 + (c[0] * np.conj(c[1]) * h01).real


def incident_coords(Y, inv_c, inv_nc, inv_n, a_s, a_p):
    """Solve [[G00, G02], [G20, G22]] c = (a_s, a_p) with G = M_prism Y (1/2 included)."""
    g = []
    for y in Y:
        g.append((0.5 * (y[1] - y[2] * inv_nc), 0.5 * (y[0] * inv_c + y[3] * inv_n)))
    (g00, g20), (g02, g22) = g
    idet = 1.0 / (g00 * g22 - g02 * g20)
    return (g22 * a_s - g02 * a_p) * idet, (g00 * a_p - g20 * a_s) * idet


# ----------------------------------------------------------------------------- thin films: monolithic cubic
def film_transfer_mono(D, Y, spec, cfac):
    """exp(cfac D) Y as ONE cubic polynomial in D (3 mat-vecs per column):
    E(x) = alpha_b(x) + (alpha_f(x) - alpha_b(x)) q_b(x) h(x)  mod  q_f(x) q_b(x).
    Only safe when the exponentials of the four modes are of comparable size (thin films)."""
    sf, pf, sb, pb = spec
    Af, Bf, mf = pair_exponential(sf, pf, cfac)
    Ab, Bb, mb = pair_exponential(sb, pb, cfac)
    af0, af1 = Af - Bf * mf, Bf
    ab0, ab1 = Ab - Bb * mb, Bb
    h0, h1 = projector_coeffs(sf, pf, sb, pb)
    d0, d1 = af0 - ab0, af1 - ab1
    g0, g1, g2 = d0 * h0, d0 * h1 + d1 * h0, d1 * h1
    c3, c2, c1, c0 = -(sf + sb), pf + pb + sf * sb, -(sf * pb + sb * pf), pf * pb
    a3 = g1 - sb * g2 - g2 * c3
    a2 = g0 - sb * g1 + pb * g2 - g2 * c2
    a1 = -sb * g0 + pb * g1 - g2 * c1 + ab1
    a0 = pb * g0 - g2 * c0 + ab0
    out = []
    for y in Y:
        v1 = matvec(D, y)
        v2 = matvec(D, v1)
        v3 = matvec(D, v2)
        out.append(tuple(a0 * y[i] + a1 * v1[i] + a2 * v2[i] + a3 * v3[i] for i in range(4)))
    return out


def film_growth(spec, cfac):
    """max |Re(cfac lambda)| over the four modes."""
    sf, pf, sb, pb = spec
    g = 0.0
    for s, p in ((sf, pf), (sb, pb)):
        m = 0.5 * s
        h = np.sqrt(m * m - p)
        g = np.maximum(g, np.maximum(np.abs((cfac * (m + h)).real), np.abs((cfac * (m - h)).real)))
    return g


# ----------------------------------------------------------------------------- un-tilted crystals: pair-symmetric film
PAIR_ODD = 1e-12        # |c3|^2 <= PAIR_ODD |c2|, |c1|^2 <= PAIR_ODD |c2|^3: odd coefficients are a small perturbation
PAIR_SEPARATION = 1e-6  # |y0 - y1|^2 >= PAIR_SEPARATION (|y0|^2 + |y1|^2): the pair mixing loses <= 3 digits
PAIR_STEP = 1e-20       # |last polish step|^2 relative: converged


def _cosh_sinhc(z):
    """cosh(z) and sinh(z)/z, series below |z| = 0.25."""
    z = np.asarray(z, dtype=complex)
    z2 = z * z
    small = np.abs(z2) < 0.0625
    zs = np.where(small, 1.0, z)
    e = np.exp(zs)
    ei = 1.0 / e
    ch_big, sh_big = 0.5 * (e + ei), 0.5 * (e - ei) / zs
    ch = 1 + z2 / 2 * (1 + z2 / 12 * (1 + z2 / 30 * (1 + z2 / 56 * (1 + z2 / 90 * (1 + z2 / 132)))))
    sh = 1 + z2 / 6 * (1 + z2 / 20 * (1 + z2 / 42 * (1 + z2 / 72 * (1 + z2 / 110 * (1 + z2 / 156)))))
    return np.where(small, ch, ch_big), np.where(small, sh, sh_big)


def pair_symmetric_split(c3, c2, c1, c0, iters=3):
    """Quartic with small odd coefficients (un-tilted crystal + the reference's hidden 1e-8 tilt):
    factors (x^2 + u x + v)(x^2 + b1 x + b0) grouped as the +-mu pairs of the biquadratic limit,
    u, b1 ~ tilt.  Start (0, -y0); the Bairstow Jacobian there is -(y0 - y1) I up to O(tilt), so
    the polish is the fixed-point step (u, v) += (r1, r0) / (y0 - y1): linear convergence at rate
    ~ tilt |mu| / |y0 - y1|, three steps."""
    y0, y1 = _quad_roots(c2, c0)
    dy = y0 - y1
    idy = 1.0 / np.where(dy == 0, 1.0, dy)
    u, v = np.zeros_like(y0), -y0
    done = np.zeros(np.shape(y0), dtype=bool)
    for it in range(iters):   # two steps; a third only where the second still moved
        b1 = c3 - u
        b0 = c2 - v - u * b1
        du, dv = (c1 - u * b0 - v * b1) * idy, (c0 - v * b0) * idy
        u, v = np.where(done, u, u + du), np.where(done, v, v + dv)
        if it >= 1:
            done = done | ((np.abs(dv) ** 2 <= PAIR_STEP * np.abs(v) ** 2)
                           & (np.abs(du) ** 4 <= PAIR_STEP ** 2 * np.abs(v) ** 2))
    b1 = c3 - u
    b0 = c2 - v - u * b1
    n2 = np.abs(c2) ** 2
    ok = (np.abs(c3) ** 4 <= PAIR_ODD * n2) & (np.abs(c1) ** 4 <= PAIR_ODD * n2 ** 3)
    ok &= np.abs(dy) ** 2 >= PAIR_SEPARATION * (np.abs(y0) ** 2 + np.abs(y1) ** 2)
    ok &= done
    return u, v, b1, b0, ok


def film_pair_symmetric(D, Y, cfac):
    """exp(cfac D) Y for a film whose quartic is a slightly perturbed biquadratic.

    Root pairs delta_i +- mu_i (delta_i ~ tilt): on each pair exp(c x) is interpolated by
        alpha_i(x) = e^{c delta_i} [cosh(c mu_i) + (x - delta_i) sinh(c mu_i) / mu_i],
    entire and even in mu_i: no forward/backward classification, one complex exponential per pair.
    The two interpolants are merged into one cubic by the same projector formula as the general
    thin-film path.  Returns (columns, ok)."""
    c3, c2, c1, c0 = charpoly(D)
    u, v, b1, b0, ok = pair_symmetric_split(c3, c2, c1, c0)
    coef = []
    for (uu, vv) in ((u, v), (b1, b0)):
        delta = -0.5 * uu
        mu = np.sqrt(delta * delta - vv)
        z = cfac * mu
        ok = ok & (np.abs(z.real) < MONO_GROWTH)
        A, S = _cosh_sinhc(z)
        x = cfac * delta
        ok = ok & (np.abs(x) ** 2 < 1e-10)
        ed = 1 + x * (1 + 0.5 * x)
        coef.append((delta, ed * A, ed * cfac * S))
    (d0, A0, B0), (d1, A1, B1) = coef
    sf, pf, sb, pb = -u, v, -b1, b0
    af0, af1 = A0 - B0 * d0, B0
    ab0, ab1 = A1 - B1 * d1, B1
    h0, h1 = projector_coeffs(sf, pf, sb, pb)
    e0, e1 = af0 - ab0, af1 - ab1
    g0, g1, g2 = e0 * h0, e0 * h1 + e1 * h0, e1 * h1
    a3 = g1 - sb * g2 - g2 * c3
    a2 = g0 - sb * g1 + pb * g2 - g2 * c2
    a1 = -sb * g0 + pb * g1 - g2 * c1 + ab1
    a0 = pb * g0 - g2 * c0 + ab0
    out = []
    for y in Y:
        v1 = matvec(D, y)
        v2 = matvec(D, v1)
        v3 = matvec(D, v2)
        out.append(tuple(a0 * y[i] + a1 * v1[i] + a2 * v2[i] + a3 * v3[i] for i in range(4)))
    return out, ok


# ----------------------------------------------------------------------------- uniaxial crystals: factorised quartic
def _cosh_sinhc_w(w):
    """cosh(sqrt(w)) and sinh(sqrt(w))/sqrt(w) as entire functions of w (model of ho_berreman.cuh
    cosh_sinhc_pair / cosh_sinhc_w): Taylor series in w below |w| = 1 -- seven terms below 0.1, ten
    below 1 -- and the exponential form above; no complex square root on the series path."""
    import math

    w = np.asarray(w, dtype=complex)
    ca = [1.0 / math.factorial(2 * k) for k in range(9, -1, -1)]
    cb = [1.0 / math.factorial(2 * k + 1) for k in range(9, -1, -1)]

    def horner(first):
        a, b = np.full(w.shape, ca[first], dtype=complex), np.full(w.shape, cb[first], dtype=complex)
        for i in range(first + 1, 10):
            a, b = w * a + ca[i], w * b + cb[i]
        return a, b

    a7, b7 = horner(3)
    a10, b10 = horner(0)
    big = np.abs(w) >= 1.0
    z = np.sqrt(np.where(big, w, 1.0))
    e = np.exp(z)
    ei = 1.0 / e
    tiny = np.abs(w) < 0.1
    ch = np.where(big, 0.5 * (e + ei), np.where(tiny, a7, a10))
    sh = np.where(big, 0.5 * (e - ei) / z, np.where(tiny, b7, b10))
    return ch, sh


def _thin_pair(re_cm, w, limit=3.0):
    """|Re(c m)| + |Re sqrt(w)| < limit without the square root (ho_berreman.cuh thin_pair)."""
    lim = limit - np.abs(re_cm)
    r = 2.0 * lim * lim - np.real(w)
    return (lim > 0) & (r > 0) & (np.abs(w) ** 2 < r * r)


def uniaxial_split(c3, c2, qo):
    """Spectral split of a uniaxial crystal (model of ho_berreman.cuh split_uniaxial): the quartic
    factorises exactly, (x^2 + q_o)(x^2 + c3 x + q_e), q_o = kx^2 - mu eps_o, q_e = c2 - q_o; the four
    roots are two square roots away, the forward pair follows from the reference's rule on them.
    -> (sf, pf, sb, pb)."""
    lo = np.sqrt(-qo + 0j)
    le0, le1 = _quad_roots(c3, c2 - qo)
    roots = [lo, -lo, le0, le1]
    fwd, _ = classify_forward(roots)
    sf = sum(np.where(m, lam, 0.0) for m, lam in zip(fwd, roots))
    pf = np.prod([np.where(m, lam, 1.0) for m, lam in zip(fwd, roots)], axis=0)
    sb = sum(np.where(m, 0.0, lam) for m, lam in zip(fwd, roots))
    pb = np.prod([np.where(m, 1.0, lam) for m, lam in zip(fwd, roots)], axis=0)
    return sf, pf, sb, pb


def film_pair_uniaxial(D, Y, cfac, qo):
    """exp(cfac D) Y for a thin film of a uniaxial crystal at ANY orientation (model of
    ho_berreman.cuh pair_factors_uniaxial + pair_symmetric_film_coeffs): the ordinary pair +-mu_o
    (mu_o^2 = -q_o) and the extraordinary pair -c3/2 +- mu_e are known in closed form, so the cubic
    needs neither a split nor the odd quartic coefficients.  Returns (columns, ok)."""
    c3, c2, _, _ = charpoly(D)
    sf, pf, sb, pb = np.zeros_like(c3), qo + 0 * c3, -c3, c2 - qo
    d0, d1 = 0.5 * sf, 0.5 * sb
    m0, m1 = d0 * d0 - pf, d1 * d1 - pb
    ok = np.abs(m0 - m1) ** 2 >= PAIR_SEPARATION * (np.abs(m0) ** 2 + np.abs(m1) ** 2)
    cc = cfac * cfac
    w0, w1 = cc * m0, cc * m1
    x1 = cfac * d1
    ok = ok & _thin_pair(0.0, w0) & _thin_pair(np.real(x1), w1)
    A0, S0 = _cosh_sinhc_w(w0)
    A1, S1 = _cosh_sinhc_w(w1)
    e1 = np.where(np.abs(x1) ** 2 < 1e-10, 1 + x1 * (1 + 0.5 * x1), np.exp(x1))
    B0, B1 = cfac * S0, e1 * cfac * S1
    A1 = e1 * A1
    af0, af1 = A0 - B0 * d0, B0
    ab0, ab1 = A1 - B1 * d1, B1
    h0, h1 = projector_coeffs(sf, pf, sb, pb)
    e0_, e1_ = af0 - ab0, af1 - ab1
    g0, g1, g2 = e0_ * h0, e0_ * h1 + e1_ * h0, e1_ * h1
    q3, q2, q1, q0 = -(sf + sb), sf * sb + pf + pb, -(sf * pb + sb * pf), pf * pb   # q_0 q_1
    a3 = g1 - sb * g2 - g2 * q3
    a2 = g0 - sb * g1 + pb * g2 - g2 * q2
    a1 = -sb * g0 + pb * g1 - g2 * q1 + ab1
    a0 = pb * g0 - g2 * q0 + ab0
    out = []
    for y in Y:
        v1 = matvec(D, y)
        v2 = matvec(D, v1)
        v3 = matvec(D, v2)
        out.append(tuple(a0 * y[i] + a1 * v1[i] + a2 * v2[i] + a3 * v3[i] for i in range(4)))
    return out, ok
