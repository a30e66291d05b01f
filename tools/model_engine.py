"""Run a ``StackPlan`` through the NumPy model of the device algorithm (dev/test tool).

Same table inputs as the CUDA engine, same right-to-left sweep over layers with a
running 4x2 "plane" (the two columns of Gamma the reflection formulas need).
"""

from __future__ import annotations

import numpy as np

from hyperbolic_optics_b200 import plan as P
from tools import algo_model as am


def _grid(plan):
    f, a, b, t = np.meshgrid(np.arange(plan.F), np.arange(plan.A), np.arange(plan.B),
                             np.arange(plan.T), indexing="ij")
    return f, a, b, t


def _pick(table, idx):
    return table[idx if table.shape[0] > 1 else np.zeros_like(idx)]


def layer_delta(plan, lp, f, a, b):
    k = plan.kx[a]
    cb, sb = _pick(lp.cos_beta, b), _pick(lp.sin_beta, b)
    e = tuple(_pick(lp.eps, f)[..., i] for i in range(6))
    e = am.rotate_z(e, cb, sb)
    m = tuple(_pick(lp.mu, f)[..., i] for i in range(6))
    if not lp.mu_scalar:
        m = am.rotate_z(m, cb, sb)
    return am.berreman9(k + 0j, e, m)


def _ordinary_factor(plan, lp, f, k):
    """q_o = kx^2 - mu eps_o of a uniaxial layer (plan.eps_ordinary), else None."""
    if getattr(lp, "eps_ordinary", None) is None:
        return None
    eo = _pick(lp.eps_ordinary.reshape(-1, 1), f)[..., 0]
    mu = _pick(lp.mu, f)[..., 0]
    return k * k - mu * eo


def run(plan, backend="transfer", stats=None, uniaxial=True, **kw):
    """`uniaxial=False`: ignore the uniaxial hint (the general split on every layer)."""
    f, a, b, t = _grid(plan)
    k = plan.kx[a] + 0j
    k0 = plan.k0[f]
    stable = backend == "scattering"
    Y = None
    forms, maps = [], []  # per layer (bottom -> top): flux form at its top interface, W (stable)
    with np.errstate(all="ignore"):
        for lp in reversed(plan.layers):
            if lp.kind == P.KIND_ISO_EXIT:
                cf = lp.exit_cos[a]
                Y = am.exit_iso_plane(cf, lp.exit_n + 0j)
                forms.append(am.flux_form(Y)); maps.append(None)
                continue
            if lp.kind == P.KIND_ANISO_EXIT:
                D = layer_delta(plan, lp, f, a, b)
                qo = _ordinary_factor(plan, lp, f, k)
                if qo is not None and uniaxial:
                    # uniaxial crystal: forward / backward pairs straight from the factorised quartic
                    c3, c2, _, _ = am.charpoly(D)
                    spec = am.uniaxial_split(c3, c2, qo) + (None,)
                else:
                    spec = am.split_spectrum(D, stats=stats, **kw)
                    if stats is not None:
                        stats.setdefault("n_complex", []).append(spec[4])
                Y = am.substrate_plane(D, spec[:4])
                forms.append(am.flux_form(Y)); maps.append(None)
                continue
            d = _pick(lp.thickness, t)
            cfac = -1j * k0 * d
            if lp.kind == P.KIND_ISO_FILM:
                eps = lp.iso_eps + 0 * k
                mu = lp.iso_mu + 0 * k
                if stable:
                    Y, w = am.iso_film_stable(k, eps, mu, Y, cfac, return_w=True)
                else:
                    w = None
                    Y = am.iso_film_transfer(k, eps, mu, Y, cfac)
            else:
                D = layer_delta(plan, lp, f, a, b)
                qo = _ordinary_factor(plan, lp, f, k) if uniaxial else None
                if qo is not None:
                    c3, c2, _, _ = am.charpoly(D)
                    spec = am.uniaxial_split(c3, c2, qo) + (None,)
                else:
                    spec = am.split_spectrum(D, stats=stats, **kw)
                    if stats is not None:
                        stats.setdefault("n_complex", []).append(spec[4])
                Y_in = Y
                if stable:
                    Y, w = am.film_stable(D, Y_in, spec[:4], cfac, return_w=True)
                else:
                    Y, w = am.film_transfer(D, Y_in, spec[:4], cfac, return_w=True)
                if plan.fast_path or qo is not None:
                    # un-tilted or uniaxial crystals: the kernels take the pair closed form where it applies
                    Yp, ok = am.film_pair_uniaxial(D, Y_in, cfac, qo) if qo is not None else am.film_pair_symmetric(D, Y_in, cfac)
                    one, zero = np.ones_like(Yp[0][0]), np.zeros_like(Yp[0][0])
                    wp = ((one, zero), (zero, one))
                    if stable:
                        n00, n01, n10, n11 = am._inv2(Yp[0][0], Yp[1][0], Yp[0][1], Yp[1][1])
                        Yp = [tuple(Yp[0][i] * n00 + Yp[1][i] * n10 for i in range(4)),
                              tuple(Yp[0][i] * n01 + Yp[1][i] * n11 for i in range(4))]
                        wp = ((n00, n01), (n10, n11))
                    Y = [tuple(np.where(ok, Yp[k][i], Y[k][i]) for i in range(4)) for k in range(2)]
                    w = tuple(tuple(np.where(ok, wp[r][c], w[r][c]) for c in range(2)) for r in range(2))
                    if stats is not None:
                        stats.setdefault("pair_symmetric", []).append(float(np.mean(ok)))
            forms.append(am.flux_form(Y)); maps.append(w)
        r_pp, r_ps, r_sp, r_ss = am.reflect_from_plane(
            Y, plan.prism_inv_cos[a], plan.prism_inv_ncos[a], plan.prism_inv_n)
        mu4 = am.mueller(r_pp, r_ps, r_sp, r_ss)
        out = {"r_pp": r_pp, "r_ps": r_ps, "r_sp": r_sp, "r_ss": r_ss, "mueller": mu4}
        # power bookkeeping (reference fields.py:120-282): fluxes at every interface, top -> bottom
        inv_c, inv_nc = plan.prism_inv_cos[a], plan.prism_inv_ncos[a]
        s_inc = 0.5 / inv_nc
        for pol, (a_s, a_p) in (("p", (0.0, 1.0)), ("s", (1.0, 0.0))):
            c = am.incident_coords(Y, inv_c, inv_nc, plan.prism_inv_n, a_s, a_p)
            flux = []
            for h, w in zip(reversed(forms), reversed(maps)):
                flux.append(am.flux_eval(h, c))
                if w is not None:
                    c = (w[0][0] * c[0] + w[0][1] * c[1], w[1][0] * c[0] + w[1][1] * c[1])
            out[f"R_{pol}"] = 1.0 - flux[0] / s_inc
            out[f"T_{pol}"] = flux[-1] / s_inc
            out[f"A_{pol}"] = [(flux[i] - flux[i + 1]) / s_inc for i in range(len(flux) - 1)]
    return out


def jones_error(got, ref):
    """Per-point Frobenius-relative error of J = [[r_pp, r_ps], [r_sp, r_ss]]."""
    num = sum(np.abs(np.asarray(got[k]) - np.asarray(ref[k])) ** 2 for k in ("r_pp", "r_ps", "r_sp", "r_ss"))
    den = sum(np.abs(np.asarray(ref[k])) ** 2 for k in ("r_pp", "r_ps", "r_sp", "r_ss"))
    return np.sqrt(num / den)
