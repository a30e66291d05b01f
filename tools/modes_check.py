"""Diagnostic: materialised objects (transfer_matrix, layer matrices, wave profiles) of the CUDA
engine vs the reference fixtures tests/golden/modes_*.npz -- prints error statistics per layer."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
NAMES = ("simple_calcite", "incident_calcite", "multilayer_incident", "magnetic_gap_incident", "fullsweep_quartz",
         "dispersion_iso_exit", "thickness_axis_film", "c4_fullsweep_hbn_sic")
from tests.test_gpu_parity import _run_gpu  # noqa: E402


def sample(arr, strides, matrix_ndim):
    arr = np.asarray(arr)
    sl = tuple(slice(None, None, int(s) if n > 1 else 1) for n, s in zip(arr.shape[:4], strides))
    return arr[sl + (slice(None),) * matrix_ndim]


def rel(a, b, nd):
    ax = tuple(range(-nd, 0))
    with np.errstate(all="ignore"):
        return np.sqrt((np.abs(a - b) ** 2).sum(axis=ax) / (np.abs(b) ** 2).sum(axis=ax))


def main():
    for name in sys.argv[1:] or NAMES:
        fx = dict(np.load(ROOT / "tests" / "golden" / f"modes_{name}.npz"))
        s = _run_gpu(name, "transfer")
        st = fx["strides"]
        g = sample(s.transfer_matrix, st, 2)
        print(f"== {name}: gamma {s.transfer_matrix.shape} vs {tuple(fx['gamma_shape'])}")
        e = rel(g, fx["gamma"], 2)
        print(f"   gamma rel err: max {np.nanmax(e):.2e} median {np.nanmedian(e):.2e} nan {np.isnan(e).sum()}")
        ecol = np.abs(g - fx["gamma"]).max(axis=-2) / np.abs(fx["gamma"]).max(axis=(-1, -2))[..., None]
        print("   gamma per-column max abs/scale:", np.nanmax(ecol.reshape(-1, 4), axis=0))
        for i in range(int(fx["n_layers"])):
            layer = s.layers[i]
            m = np.asarray(layer.matrix)
            want_shape = tuple(fx[f"L{i}_shape"])
            em = rel(sample(m, st, 2), fx[f"L{i}_matrix"], 2)
            print(f"   L{i} {layer.type}: matrix {m.shape[:4]} vs {want_shape}: rel max {np.nanmax(em):.2e}")
            if f"L{i}_V" in fx:
                v, kz = layer.profile.tangential_modes()
                w, _ = layer.profile.full_modes()
                kz_s, v_s, w_s = sample(kz, st, 1), sample(v, st, 2), sample(w, st, 2)
                rk = fx[f"L{i}_kz"]
                gap = np.minimum(np.abs(rk[..., 0] - rk[..., 1]), np.abs(rk[..., 2] - rk[..., 3])) / np.abs(rk).max(axis=-1)
                ek = np.abs(kz_s - rk).max(axis=-1) / np.abs(rk).max(axis=-1)
                ev = np.abs(v_s - fx[f"L{i}_V"]).max(axis=(-1, -2))
                ew = np.abs(w_s - fx[f"L{i}_W"]).max(axis=(-1, -2)) / np.abs(fx[f"L{i}_W"]).max(axis=(-1, -2))
                p = np.stack([np.concatenate([getattr(layer.profile, f"transmitted_{c}"), getattr(layer.profile, f"reflected_{c}")], axis=-1)
                              for c in ("Px", "Py", "Pz")], axis=-2)
                ep = np.abs(sample(p, st, 2) - fx[f"L{i}_P"]).max(axis=(-1, -2))
                ok = gap > 1e-6
                print(f"      profile {kz.shape[:4]}: min gap {gap.min():.2e}; non-degenerate {ok.mean():.2f}: kz {np.max(ek[ok], initial=0):.2e} "
                      f"V {np.max(ev[ok], initial=0):.2e} W {np.max(ew[ok], initial=0):.2e} P {np.max(ep[ok], initial=0):.2e}; "
                      f"all: kz {ek.max():.2e} V {ev.max():.2e}")
                # forward / backward subspaces (well defined also in a degenerate eigenspace)
                for lo in (0, 2):
                    a, b = v_s[..., lo:lo + 2], fx[f"L{i}_V"][..., lo:lo + 2]
                    coef = np.linalg.solve(np.conj(np.swapaxes(b, -1, -2)) @ b, np.conj(np.swapaxes(b, -1, -2)) @ a)
                    res = np.abs(a - b @ coef).max(axis=(-1, -2))
                    print(f"      subspace cols {lo},{lo + 1}: residual max {res.max():.2e}")


if __name__ == "__main__":
    main()
