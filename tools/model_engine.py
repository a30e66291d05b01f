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


def run(plan, backend="transfer", stats=None, **kw):
    f, a, b, t = _grid(plan)
    k = plan.kx[a] + 0j
    k0 = plan.k0[f]
    stable = backend == "scattering"
    Y = None
    with np.errstate(all="ignore"):
        for lp in reversed(plan.layers):
            if lp.kind == P.KIND_ISO_EXIT:
                cf = lp.exit_cos[a]
                Y = am.exit_iso_plane(cf, lp.exit_n + 0j)
                continue
            if lp.kind == P.KIND_ANISO_EXIT:
                D = layer_delta(plan, lp, f, a, b)
                spec = am.split_spectrum(D, **kw)
                if stats is not None:
                    stats.setdefault("n_complex", []).append(spec[4])
                Y = am.substrate_plane(D, spec[:4])
                continue
            d = _pick(lp.thickness, t)
            cfac = -1j * k0 * d
            if lp.kind == P.KIND_ISO_FILM:
                eps = lp.iso_eps + 0 * k
                mu = lp.iso_mu + 0 * k
                if stable:
                    Y = am.iso_film_stable(k, eps, mu, Y, cfac)
                else:
                    Y = am.iso_film_transfer(k, eps, mu, Y, cfac)
            else:
                D = layer_delta(plan, lp, f, a, b)
                spec = am.split_spectrum(D, **kw)
                if stats is not None:
                    stats.setdefault("n_complex", []).append(spec[4])
                if stable:
                    Y = am.film_stable(D, Y, spec[:4], cfac)
                else:
                    Y = am.film_transfer(D, Y, spec[:4], cfac)
        r_pp, r_ps, r_sp, r_ss = am.reflect_from_plane(
            Y, plan.prism_inv_cos[a], plan.prism_inv_ncos[a], plan.prism_inv_n)
        mu4 = am.mueller(r_pp, r_ps, r_sp, r_ss)
    return {"r_pp": r_pp, "r_ps": r_ps, "r_sp": r_sp, "r_ss": r_ss, "mueller": mu4}


def jones_error(got, ref):
    """Per-point Frobenius-relative error of J = [[r_pp, r_ps], [r_sp, r_ss]]."""
    num = sum(np.abs(np.asarray(got[k]) - np.asarray(ref[k])) ** 2 for k in ("r_pp", "r_ps", "r_sp", "r_ss"))
    den = sum(np.abs(np.asarray(ref[k])) ** 2 for k in ("r_pp", "r_ps", "r_sp", "r_ss"))
    return np.sqrt(num / den)
