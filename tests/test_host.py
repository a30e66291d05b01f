"""Host-side logic (CPU): scenario grids, material tables, planner, error contract, and the
C-ABI library (loads, exports every symbol include/ho_b200.h declares -- no compute calls)."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from hyperbolic_optics_b200 import materials, plan as planmod
from hyperbolic_optics_b200.scenario import ScenarioSetup, present
from oracle import reference_port as oracle
from tests.golden.payloads import ALL, REFERENCE_BATTERY, angle_overrides

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("kind,sd", [
    ("Incident", {"type": "Incident"}),
    ("Azimuthal", {"type": "Azimuthal", "incidentAngle": 40}),
    ("Dispersion", {"type": "Dispersion", "frequency": 1460.0}),
    ("FullSweep", {"type": "FullSweep"}),
    ("Simple", {"type": "Simple", "incidentAngle": 45.0, "azimuthal_angle": 10.0, "frequency": 1460.0}),
])
def test_scenario_grids_equal_reference_port(kind, sd):
    sc = ScenarioSetup(sd)
    inc, az, freq = oracle.scenario_grids(sd)
    np.testing.assert_array_equal(np.asarray(sc.incident_angle), np.asarray(inc))
    if az is not None:
        np.testing.assert_array_equal(np.asarray(sc.azimuthal_angle), np.asarray(az))
    assert sc.frequency == freq


def test_default_shapes():
    assert ScenarioSetup({"type": "Incident"}).incident_angle.shape == (360,)
    sc = ScenarioSetup({"type": "Dispersion", "frequency": 900.0})
    assert sc.incident_angle.shape == (180,) and sc.azimuthal_angle.shape == (480,)
    sc = ScenarioSetup({"type": "FullSweep"})
    assert sc.incident_angle.shape == (180,) and sc.azimuthal_angle.shape == (120,)
    assert ScenarioSetup({"type": "Incident"}, n_incident=4096).incident_angle.shape == (4096,)


@pytest.mark.parametrize("name", ["Quartz", "Sapphire", "Calcite", "CalciteLower", "AlN", "SiC", "hBN", "GaN",
                                  "MoO3", "GalliumOxide"])
def test_material_tables_equal_oracle(name):
    mat = materials.create_material(name)
    freq = mat.frequency
    assert freq.shape == (410,)
    np.testing.assert_array_equal(freq, oracle.default_frequency(name))
    eps_ref, mu_ref = oracle.material_tensors(name, freq)
    np.testing.assert_allclose(materials.unpack_symmetric(mat.eps_packed(freq)), eps_ref, rtol=1e-15, atol=0)
    np.testing.assert_array_equal(materials.unpack_symmetric(mat.mu_packed(freq))[0], mu_ref[0])


def test_arbitrary_material():
    spec = {"eps_xx": {"real": 2.27, "imag": 0.001}, "eps_yy": -4.84, "eps_zz": "1+2j", "eps_xy": 0.3, "mu_r": 2.0}
    mat = materials.create_material(spec)
    assert mat.frequency is None
    eps_ref, mu_ref = oracle.material_tensors(spec, [1000.0])
    np.testing.assert_array_equal(materials.unpack_symmetric(mat.eps_packed([1000.0]))[0], eps_ref)
    np.testing.assert_array_equal(materials.unpack_symmetric(mat.mu_packed([1000.0]))[0], mu_ref)


def test_error_contract():
    with pytest.raises(NotImplementedError):
        materials.create_material("Unobtainium")
    with pytest.raises(NotImplementedError):
        ScenarioSetup({"type": "Nope"})
    sc = ScenarioSetup({"type": "Incident"})
    with pytest.raises(ValueError):
        planmod.build_plan(sc, [{"type": "Ambient Incident Layer", "permittivity": 5.0}, {"type": "Bogus Layer"}])
    two_swept = [{"type": "Ambient Incident Layer", "permittivity": 5.0},
                 {"type": "Isotropic Middle-Stack Layer", "thickness": [0.1, 0.2]},
                 {"type": "Crystal Layer", "material": "Quartz", "thickness": [0.1, 0.2]},
                 {"type": "Semi Infinite Anisotropic Layer", "material": "Quartz"}]
    with pytest.raises(ValueError):
        planmod.build_plan(ScenarioSetup({"type": "Incident"}), two_swept)
    with pytest.raises(ValueError):
        planmod.build_plan(ScenarioSetup({"type": "Incident"}),
                           [{"type": "Ambient Incident Layer", "permittivity": 5.0},
                            {"type": "Semi Infinite Isotropic Layer", "permittivity": 1.0}])  # no frequency


@pytest.mark.parametrize("name", list(ALL))
def test_plan_tables_match_oracle_stack(name):
    """kx, k0, prism scalings, rotated tensors and thickness of the planner == the port's arrays."""
    entry = ALL[name]
    inc, az = angle_overrides(entry)
    sc = ScenarioSetup(entry["payload"]["ScenarioData"])
    if inc is not None:
        sc.incident_angle = inc
    if az is not None:
        sc.azimuthal_angle = az
    plan = planmod.build_plan(sc, entry["payload"]["Layers"])
    with np.errstate(all="ignore"):
        st = oracle.build_stack(entry["payload"], inc, az)
    np.testing.assert_array_equal(plan.kx, st["kx"].reshape(-1))
    np.testing.assert_array_equal(plan.k0, st["k0"].reshape(-1))
    prism = st["layers"][0].matrix.reshape(-1, 4, 4)
    np.testing.assert_allclose(0.5 * plan.prism_inv_cos, prism[:, 2, 0].real, rtol=1e-15)
    np.testing.assert_allclose(0.5 * plan.prism_inv_ncos, prism[:, 1, 2].real, rtol=1e-15)
    assert plan.n_points == int(np.prod([plan.A, plan.B, plan.F, plan.T]))
    for lp, ref in zip(plan.layers, st["layers"][1:]):
        assert lp.type == ref.kind
        if lp.thickness is not None:
            np.testing.assert_allclose(lp.thickness, np.asarray(ref.thickness).reshape(-1), rtol=1e-15)
        if lp.kind == planmod.KIND_ISO_EXIT:
            np.testing.assert_allclose(lp.exit_cos, ref.matrix.reshape(-1, 4, 4)[:, 0, 2], rtol=1e-15)


def test_present_squeezes_like_reference():
    x = np.arange(2 * 3 * 1 * 1).reshape(2, 3, 1, 1)
    assert present(x).shape == (2, 3)
    assert present(np.zeros((1, 1, 1, 1))).shape == ()
    assert present(np.zeros((5, 1, 4, 1, 4, 4))).shape == (5, 4, 4, 4)


def test_library_exports_every_declared_symbol(built_library):
    header = (ROOT / "include" / "ho_b200.h").read_text()
    declared = set(re.findall(r"\b(ho_[a-z0-9_]+)\s*\(", header))
    assert {"ho_reflect", "ho_reflect_f32", "ho_polarimetry", "ho_material_table", "ho_layer_modes", "ho_transfer_matrix",
            "ho_measure_fp64_peak", "ho_launch_count", "ho_abi_version", "ho_last_error"} <= declared
    for sym in declared:
        assert hasattr(built_library, sym), f"{sym} declared in ho_b200.h but not exported"
    from hyperbolic_optics_b200 import _native

    version = int(re.search(r"#define HO_ABI_VERSION (\d+)", header).group(1))
    assert built_library.ho_abi_version() == version == _native.HO_ABI_VERSION
    assert built_library.ho_launch_count() == 0


def test_descriptor_validation_without_gpu(built_library):
    """Bad descriptors are rejected on the host before any CUDA call."""
    from hyperbolic_optics_b200 import _native as nat

    prob = nat.Problem()
    out = nat.Outputs()
    rc = built_library.ho_reflect(ctypes.byref(prob), ctypes.byref(out), 0, 0, None)
    assert rc == -1 and b"batch sizes" in built_library.ho_last_error()
    prob.A = prob.B = prob.F = prob.T = 1
    prob.n_layers = 99
    assert built_library.ho_reflect(ctypes.byref(prob), ctypes.byref(out), 0, 0, None) == -1
    req = nat.Polarimetry()
    assert built_library.ho_polarimetry(ctypes.byref(req), 16, None) == -1
    assert b"Jones planes" in built_library.ho_last_error()
    assert built_library.ho_polarimetry(ctypes.byref(req), 0, None) == 0  # empty range: nothing to do
    mat = nat.MaterialDesc()
    assert built_library.ho_material_table(ctypes.byref(mat), None, 4, None, None) == -1
    modes = nat.LayerModes()
    assert built_library.ho_layer_modes(ctypes.byref(prob), 0, ctypes.byref(modes), None) == -1
    assert built_library.ho_transfer_matrix(ctypes.byref(prob), None, 0, 0, None) == -1


def test_gather_fields_validation_without_gpu(built_library):
    """ho_outputs_t.peer_r / rebuild_* (fused all-gather) are checked on the host before any CUDA call."""
    from hyperbolic_optics_b200 import _native as nat

    kx = (ctypes.c_double * 1)(0.5)
    one = (ctypes.c_double * 2)(1.0, 0.0)
    prob = nat.Problem()
    prob.A = prob.B = prob.F = prob.T = 1
    prob.n_layers = 1
    for name in ("kx", "k0", "prism_inv_cos", "prism_inv_ncos"):
        setattr(prob, name, ctypes.addressof(kx))
    prob.layers[0].kind = nat.HO_ISO_EXIT if hasattr(nat, "HO_ISO_EXIT") else 4
    prob.layers[0].exit_cos = ctypes.addressof(one)
    dummy = ctypes.addressof(one)

    def call(out, n_end=1):
        rc = built_library.ho_reflect(ctypes.byref(prob), ctypes.byref(out), 0, n_end, None)
        return rc, built_library.ho_last_error()

    out = nat.Outputs()
    out.n_peers = nat.HO_MAX_PEERS + 1
    rc, msg = call(out)
    assert rc == -1 and b"n_peers" in msg
    out = nat.Outputs()
    out.n_rebuild = -1
    rc, msg = call(out)
    assert rc == -1 and b"n_rebuild" in msg
    out = nat.Outputs()
    out.n_peers = 1                                    # NULL destination
    rc, msg = call(out)
    assert rc == -1 and b"peer_r" in msg
    out = nat.Outputs()
    out.n_rebuild = 1
    out.rebuild_n[0] = 2                               # longer than the launched range (thread i <-> element i)
    out.rebuild_mueller[0] = dummy
    for k in range(4):
        out.rebuild_r[0][k] = dummy
    rc, msg = call(out)
    assert rc == -1 and b"rebuild_n" in msg
    out.rebuild_n[0] = 1
    out.rebuild_mueller[0] = None
    rc, msg = call(out)
    assert rc == -1 and b"rebuild_mueller" in msg
    out.rebuild_mueller[0] = dummy
    out.rebuild_r[0][2] = None
    rc, msg = call(out)
    assert rc == -1 and b"rebuild_r" in msg
    for k in range(4):
        out.rebuild_r[0][k] = dummy
    out.power, out.power_stride = dummy, 1             # gather launches carry no power / transmission outputs
    rc, msg = call(out)
    assert rc == -1 and b"power" in msg


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors of the ABI structs vs what a C compiler makes of include/ho_b200.h."""
    import subprocess

    from hyperbolic_optics_b200 import _native as nat

    probe = tmp_path / "probe.c"
    probe.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ho_b200.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ho_layer_t), sizeof(ho_problem_t), '
        'offsetof(ho_problem_t, layers), offsetof(ho_layer_t, eps), offsetof(ho_layer_t, exit_n), '
        'sizeof(ho_outputs_t), offsetof(ho_outputs_t, t_mode), sizeof(ho_polarimetry_t), '
        'offsetof(ho_polarimetry_t, j_post), offsetof(ho_polarimetry_t, mueller)); '
        'printf("%zu %zu %zu\\n", sizeof(ho_material_t), offsetof(ho_material_t, amp), offsetof(ho_material_t, rot)); '
        'printf("%zu %zu %zu\\n", offsetof(ho_problem_t, work), offsetof(ho_problem_t, work_capacity), offsetof(ho_problem_t, kx)); '
        'printf("%zu %zu %zu\\n", sizeof(ho_layer_modes_t), offsetof(ho_layer_modes_t, mA), offsetof(ho_layer_modes_t, pF)); '
        'printf("%zu %zu %zu %zu\\n", offsetof(ho_outputs_t, n_peers), offsetof(ho_outputs_t, peer_r), '
        'offsetof(ho_outputs_t, rebuild_n), offsetof(ho_outputs_t, rebuild_mueller)); '
        'return 0; }\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", f"-I{ROOT / 'include'}", str(probe), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(nat.LayerDesc), ctypes.sizeof(nat.Problem), nat.Problem.layers.offset,
                   nat.LayerDesc.eps.offset, nat.LayerDesc.exit_n.offset, ctypes.sizeof(nat.Outputs),
                   nat.Outputs.t_mode.offset, ctypes.sizeof(nat.Polarimetry), nat.Polarimetry.j_post.offset,
                   nat.Polarimetry.mueller.offset, ctypes.sizeof(nat.MaterialDesc), nat.MaterialDesc.amp.offset,
                   nat.MaterialDesc.rot.offset, nat.Problem.work.offset, nat.Problem.work_capacity.offset, nat.Problem.kx.offset,
                   ctypes.sizeof(nat.LayerModes), nat.LayerModes.mA.offset, nat.LayerModes.pF.offset,
                   nat.Outputs.n_peers.offset, nat.Outputs.peer_r.offset, nat.Outputs.rebuild_n.offset,
                   nat.Outputs.rebuild_mueller.offset]


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no product module may reference it."""
    for path in (ROOT / "hyperbolic_optics_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in text.replace("# oracle", ""), f"{path} mentions the oracle"
        assert "tools." not in text and "from tools" not in text


def test_thickness_sweep_argument_errors():
    """Same guards as the reference helper (sweep.py:70-83); they fire before any device work."""
    from hyperbolic_optics_b200 import ThicknessSweep

    payload = {"ScenarioData": {"type": "Incident"},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 50.0},
                          {"type": "Crystal Layer", "material": "Calcite", "thickness": 1.0, "rotationY": 90},
                          {"type": "Semi Infinite Isotropic Layer", "permittivity": 1.0}]}
    with pytest.raises(IndexError):
        ThicknessSweep(payload, 7, [0.5])
    with pytest.raises(ValueError):
        ThicknessSweep(payload, 0, [0.5])
    with pytest.raises(ValueError):
        ThicknessSweep(payload, 1, [])


def test_result_shape_follows_reference_broadcasting():
    """The reference's result shape is the broadcast of its layer matrices: axes that no layer
    depends on (azimuth without a crystal, frequency without a finite or dispersive layer) are
    squeezed away.  The planner collapses the same axes (random stacks of the differential test)."""
    from tests.test_gpu_random_stacks import random_case

    seen = set()
    for seed in range(1000, 1080):
        entry = random_case(seed)
        inc, az = angle_overrides(entry)
        try:
            ref = oracle.run(entry["payload"], backend="scattering", incident_angle=inc, azimuthal_angle=az)
        except np.linalg.LinAlgError:
            continue
        sc = ScenarioSetup(entry["payload"]["ScenarioData"])
        if inc is not None:
            sc.incident_angle = inc
        if az is not None:
            sc.azimuthal_angle = az
        plan = planmod.build_plan(sc, entry["payload"]["Layers"])
        squeezed = tuple(n for n in plan.fabt_shape if n != 1)
        assert squeezed == np.shape(ref["r_pp"]), f"seed {seed}: {plan.fabt_shape} vs {np.shape(ref['r_pp'])}"
        seen.add((entry["payload"]["ScenarioData"]["type"], len(squeezed)))
    assert len(seen) >= 5  # several scenario types and collapsed / full ranks were exercised


def _plan_for(layers, sd=None):
    sd = sd or {"type": "Incident"}
    return planmod.build_plan(ScenarioSetup(sd, n_incident=8), layers)


def test_fast_path_hint_and_scalar_detection():
    """Planner hints the kernels rely on: `fast_path` (every crystal keeps z as a principal axis up to
    the reference's hidden 1e-8 tilt), `eps_scalar` (isotropic crystal: closed-form modes) and the
    stored Ry*Rx rotation (identity for a scalar table)."""
    prism = {"type": "Ambient Incident Layer", "permittivity": 12.5}

    def film(material, **rot):
        return dict({"type": "Crystal Layer", "material": material, "thickness": 0.1}, **rot)

    def exit_(material, **rot):
        return dict({"type": "Semi Infinite Anisotropic Layer", "material": material}, **rot)

    untilted = _plan_for([prism, film("hBN"), exit_("SiC")])
    assert untilted.fast_path
    assert untilted.layers[1].eps_scalar and not untilted.layers[0].eps_scalar
    np.testing.assert_array_equal(untilted.layers[1].rotation_xy, np.eye(3))
    # the hidden tilt is in the table: xz of the rotated hBN tensor is ~1e-8 (eps_xx - eps_zz), not 0
    full = materials.unpack_symmetric(untilted.layers[0].eps)
    assert 0 < np.abs(full[:, 0, 2]).max() < 1e-6 * np.abs(full[:, 0, 0]).max()
    assert _plan_for([prism, film("hBN", rotationZ=37.0), exit_("Calcite", rotationY=90)]).fast_path
    assert _plan_for([prism, film("MoO3", rotationY=180), exit_("Quartz", rotationY=90, rotationZ=12)]).fast_path
    assert not _plan_for([prism, film("hBN", rotationY=30), exit_("SiC")]).fast_path
    # kernel hint: 1 = un-tilted (fast build), 2 = tilted with optically thin films (lean build), 0 = general
    assert untilted.kernel_hint == 1
    assert _plan_for([prism, film("hBN", rotationY=30), exit_("SiC")]).kernel_hint == 2
    assert _plan_for([prism, film("hBN", rotationY=30, thickness=5.0), exit_("SiC")]).kernel_hint == 0
    assert not _plan_for([prism, film("hBN"), exit_("Quartz", rotationX=45)]).fast_path
    # uniaxial crystals carry their ordinary principal value (start values of the spectral split: the
    # quartic of a uniaxial medium factorises); biaxial / monoclinic / isotropic ones do not
    tilted = _plan_for([prism, film("hBN", rotationY=30, rotationZ=10), exit_("Quartz", rotationY=50)])
    for lp, name in zip(tilted.layers, ("hBN", "Quartz")):
        eo = materials.create_material(name).eps_packed(tilted.frequency)[:, 0]
        np.testing.assert_array_equal(lp.eps_ordinary, eo)
        # ... and the factorisation holds for the rotated table: det(x - Delta) = (x^2 + q_o)(x^2 + c3 x + q_e)
        e = materials.unpack_symmetric(lp.eps)[7]
        k = 1.3
        delta = np.array([[-k * e[0, 2] / e[2, 2], -k * e[1, 2] / e[2, 2], 0, 1 - k * k / e[2, 2]], [0, 0, -1, 0],
                          [e[1, 2] * e[0, 2] / e[2, 2] - e[0, 1], e[1, 2] ** 2 / e[2, 2] - e[1, 1] + k * k, 0, k * e[1, 2] / e[2, 2]],
                          [e[0, 0] - e[0, 2] ** 2 / e[2, 2], -(e[1, 2] * e[0, 2] / e[2, 2] - e[0, 1]), 0, -k * e[0, 2] / e[2, 2]]])
        c = np.poly(delta)
        qo = k * k - eo[7]
        assert abs(c[3] - c[1] * qo) <= 1e-12 * abs(c[3]) + 1e-12 and abs(c[4] - qo * (c[2] - qo)) <= 1e-12 * abs(c[4]) + 1e-12
    assert untilted.layers[1].eps_ordinary is None                       # SiC: isotropic
    assert _plan_for([prism, film("MoO3", rotationY=20), exit_("SiC")]).layers[0].eps_ordinary is None   # biaxial
    assert _plan_for([prism, film("hBN"), exit_("GalliumOxide", rotationY=20)]).layers[1].eps_ordinary is None   # monoclinic
    # an isotropic crystal may be tilted at will: R (eps I) R^T = eps I
    assert _plan_for([prism, film("hBN"), exit_("SiC", rotationY=33, rotationX=12)]).fast_path
    # no crystal at all: nothing to hint against
    iso = _plan_for([prism, {"type": "Isotropic Middle-Stack Layer", "thickness": 0.2},
                     {"type": "Semi Infinite Isotropic Layer", "permittivity": 2.0}],
                    {"type": "Incident", "frequency": [900.0, 1000.0]})
    assert iso.fast_path


def test_result_buffers_are_leased_not_overwritten():
    """Host results are views of pooled page-locked buffers; the pool must never hand a buffer out
    again while a result array (or any slice of one) of an earlier run is still referenced."""
    import torch

    from hyperbolic_optics_b200.structure import _PinnedPool

    pool = _PinnedPool(allocate=lambda shape, dtype: torch.empty(shape, dtype=dtype))   # (pageable here: no GPU)
    first = pool.lease("r", (4, 10), torch.complex128)
    ptr = first.data_ptr()
    piece = first.numpy().reshape(4, 2, 5)[1]          # what Structure hands to the caller
    del first
    second = pool.lease("r", (4, 10), torch.complex128)
    assert second.data_ptr() != ptr and pool.n_buffers() == 2      # first still referenced through `piece`
    piece[...] = 7.0
    del second
    del piece
    third = pool.lease("r", (4, 10), torch.complex128)
    assert third.data_ptr() == ptr and pool.n_buffers() == 2       # released -> recycled
    held = third.numpy()
    other = pool.lease("r", (4, 20), torch.complex128)              # new shape: idle buffers of the old one are dropped
    assert pool.n_buffers() == 2 and held.shape == (4, 10)
    del other, third


def test_structure_dies_by_reference_count():
    """No reference cycle through structure.layers: dropping a Structure must release its (leased,
    page-locked) result buffers immediately, not at the next garbage collection."""
    import gc
    import weakref

    from hyperbolic_optics_b200 import Structure

    gc.disable()
    try:
        s = Structure()
        s.get_scenario({"type": "Incident", "frequency": [1000.0]})
        s.get_layers([{"type": "Ambient Incident Layer", "permittivity": 12.5},
                      {"type": "Semi Infinite Isotropic Layer", "permittivity": 2.0}])
        assert s.layers[0].matrix.shape == (360, 1, 1, 1, 4, 4) and s.layers[1].matrix.shape == (360, 1, 1, 1, 4, 4)
        assert getattr(s.layers[1], "profile", None) is None and s.transfer_matrix is None
        alive = weakref.ref(s)
        del s
        assert alive() is None
    finally:
        gc.enable()


def test_install_rebinds_the_reference_structure_and_restores_it():
    """dropin.install() / uninstall() (INTEGRATION.md route B) on the importable reference package: the
    name `Structure` is rebound in every module that holds it -- so `compose_jones`'s isinstance test
    and ThicknessSweep's own `Structure()` see the drop-in -- and restored afterwards.  No GPU needed:
    nothing is executed."""
    import importlib
    import sys

    ref_dir = ROOT / "baseline" / "_ref"
    if not (ref_dir / "hyperbolic_optics" / "structure.py").exists():
        pytest.skip("baseline/_ref is not populated (python tools/install_reference.py)")
    import hyperbolic_optics_b200 as hob

    sys.path.insert(0, str(ref_dir))
    try:
        ref_structure = importlib.import_module("hyperbolic_optics.structure")
        original = ref_structure.Structure
        assert original is not hob.Structure
        patched = hob.install()
        try:
            assert {"hyperbolic_optics.structure", "hyperbolic_optics.mueller", "hyperbolic_optics.jones",
                    "hyperbolic_optics.fields", "hyperbolic_optics.sweep"} <= set(patched)
            for name in patched:
                assert sys.modules[name].Structure is hob.Structure, name
            assert hob.dropin.installed()
            # the drop-in passes the reference's own type check (jones.py:404)
            jones = importlib.import_module("hyperbolic_optics.jones")
            assert isinstance(hob.Structure.__new__(hob.Structure), jones.Structure)
        finally:
            hob.uninstall()
        for name in patched:
            assert sys.modules[name].Structure is original, name
        assert not hob.dropin.installed()
    finally:
        sys.path.remove(str(ref_dir))
        for name in [m for m in sys.modules if m == "hyperbolic_optics" or m.startswith("hyperbolic_optics.")]:
            del sys.modules[name]
