"""ctypes binding of ``libho_b200.so`` (C ABI declared in ``include/ho_b200.h``).

The library is the product: if it is missing or fails to load, importing this module
raises -- there is no Python/NumPy fallback for the hot path.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HO_MAX_LAYERS = 24
HO_MAX_PEERS = 15
HO_ABI_VERSION = 15

HO_ISO_FILM, HO_ANISO_FILM, HO_ANISO_EXIT, HO_ISO_EXIT = 1, 2, 3, 4
HO_BACKEND_TRANSFER, HO_BACKEND_SCATTERING = 0, 1
HO_PROTOCOL_AUTO, HO_PROTOCOL_SINGLE, HO_PROTOCOL_FAST_FIXUP = 0, 1, 2
HO_HINT_NONE, HO_HINT_UNTILTED, HO_HINT_TILTED_THIN = 0, 1, 2


class c128(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class LayerDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_eps", C.c_int32), ("n_mu", C.c_int32), ("mu_scalar", C.c_int32),
        ("n_beta", C.c_int32), ("n_thickness", C.c_int32), ("eps_scalar", C.c_int32),
        ("eps", C.c_void_p), ("mu", C.c_void_p),
        ("cos_beta", C.c_void_p), ("sin_beta", C.c_void_p), ("thickness", C.c_void_p),
        ("iso_eps", c128), ("iso_mu", c128),
        ("exit_cos", C.c_void_p), ("exit_n", C.c_double),
        ("eps_ordinary", C.c_void_p),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("A", C.c_int32), ("B", C.c_int32), ("F", C.c_int32), ("T", C.c_int32),
        ("n_layers", C.c_int32), ("backend", C.c_int32), ("hint_fast_path", C.c_int32),
        ("protocol", C.c_int32), ("tsweep_groups", C.c_int32), ("max_ctas_per_sm", C.c_int32),
        ("work", C.c_void_p), ("work_capacity", C.c_int64),
        ("kx", C.c_void_p), ("k0", C.c_void_p),
        ("prism_inv_cos", C.c_void_p), ("prism_inv_ncos", C.c_void_p),
        ("prism_inv_n", C.c_double),
        ("layers", LayerDesc * HO_MAX_LAYERS),
    ]


class Outputs(C.Structure):
    _fields_ = [
        ("r_pp", C.c_void_p), ("r_ps", C.c_void_p), ("r_sp", C.c_void_p), ("r_ss", C.c_void_p),
        ("mueller", C.c_void_p), ("stokes", C.c_void_p),
        ("incident_stokes", C.c_double * 4),
        ("power", C.c_void_p), ("power_stride", C.c_int64),
        ("t_pp", C.c_void_p), ("t_ps", C.c_void_p), ("t_sp", C.c_void_p), ("t_ss", C.c_void_p),
        ("t_mode", C.c_void_p), ("t_mode_stride", C.c_int64),
        ("power_states", C.c_int32),
        ("n_peers", C.c_int32), ("peer_r", (C.c_void_p * 4) * HO_MAX_PEERS),
        ("n_rebuild", C.c_int32), ("rebuild_n", C.c_int64 * HO_MAX_PEERS),
        ("rebuild_r", (C.c_void_p * 4) * HO_MAX_PEERS), ("rebuild_mueller", C.c_void_p * HO_MAX_PEERS),
    ]


class OutputsF32(C.Structure):
    _fields_ = [
        ("r_pp", C.c_void_p), ("r_ps", C.c_void_p), ("r_sp", C.c_void_p), ("r_ss", C.c_void_p),
        ("mueller", C.c_void_p),
    ]


class Polarimetry(C.Structure):
    _fields_ = [
        ("j_pp", C.c_void_p), ("j_ps", C.c_void_p), ("j_sp", C.c_void_p), ("j_ss", C.c_void_p),
        ("s_pre", C.c_double * 4), ("m_post", C.c_double * 16),
        ("e_pre", c128 * 2), ("j_post", c128 * 4),
        ("stokes", C.c_void_p), ("polarisation", C.c_void_p),
        ("jones_vector", C.c_void_p), ("jones_stokes", C.c_void_p),
        ("ellipsometry", C.c_void_p),
        ("eigenvalues", C.c_void_p), ("eigenvectors", C.c_void_p), ("discriminant", C.c_void_p), ("overlap", C.c_void_p),
        ("decomposition", C.c_void_p), ("decomposition_matrices", C.c_void_p),
        ("mueller", C.c_void_p),
    ]


class LayerModes(C.Structure):
    _fields_ = [
        ("matrix", C.c_void_p), ("modes", C.c_void_p), ("kz", C.c_void_p), ("fields", C.c_void_p), ("poynting", C.c_void_p),
        ("mA", C.c_int32), ("mB", C.c_int32), ("mF", C.c_int32), ("mT", C.c_int32),
        ("pA", C.c_int32), ("pB", C.c_int32), ("pF", C.c_int32), ("reserved", C.c_int32),
    ]


HO_MAX_OSC = 8
HO_MAT_TOLO, HO_MAT_MONOCLINIC = 1, 2


class MaterialDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_osc", C.c_int32 * 3),
        ("eps_inf", C.c_double * 4),
        ("w_to", (C.c_double * HO_MAX_OSC) * 3), ("g_to", (C.c_double * HO_MAX_OSC) * 3),
        ("w_lo", (C.c_double * HO_MAX_OSC) * 3), ("g_lo", (C.c_double * HO_MAX_OSC) * 3),
        ("amp", (C.c_double * HO_MAX_OSC) * 2), ("alpha", C.c_double * HO_MAX_OSC),
        ("rot", C.c_double * 9),
    ]


EXPORTS = ("ho_reflect", "ho_reflect_f32", "ho_polarimetry", "ho_material_table", "ho_layer_modes", "ho_transfer_matrix",
           "ho_measure_fp64_peak", "ho_launch_count", "ho_abi_version", "ho_last_error")


def library_path() -> Path:
    override = os.environ.get("HO_B200_LIB")
    return Path(override) if override else Path(__file__).resolve().parent / "libho_b200.so"


def load():
    path = library_path()
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the reflection engine.")
    lib = C.CDLL(str(path))
    lib.ho_reflect.argtypes = [C.POINTER(Problem), C.POINTER(Outputs), C.c_int64, C.c_int64, C.c_void_p]
    lib.ho_reflect.restype = C.c_int
    lib.ho_reflect_f32.argtypes = [C.POINTER(Problem), C.POINTER(OutputsF32), C.c_int64, C.c_int64, C.c_void_p]
    lib.ho_reflect_f32.restype = C.c_int
    lib.ho_polarimetry.argtypes = [C.POINTER(Polarimetry), C.c_int64, C.c_void_p]
    lib.ho_polarimetry.restype = C.c_int
    lib.ho_material_table.argtypes = [C.POINTER(MaterialDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.ho_material_table.restype = C.c_int
    lib.ho_layer_modes.argtypes = [C.POINTER(Problem), C.c_int32, C.POINTER(LayerModes), C.c_void_p]
    lib.ho_layer_modes.restype = C.c_int
    lib.ho_transfer_matrix.argtypes = [C.POINTER(Problem), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.ho_transfer_matrix.restype = C.c_int
    lib.ho_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_void_p]
    lib.ho_measure_fp64_peak.restype = C.c_int
    lib.ho_launch_count.argtypes = []
    lib.ho_launch_count.restype = C.c_int64
    lib.ho_abi_version.argtypes = []
    lib.ho_abi_version.restype = C.c_int
    lib.ho_last_error.argtypes = []
    lib.ho_last_error.restype = C.c_char_p
    if lib.ho_abi_version() != HO_ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.ho_abi_version()} != expected {HO_ABI_VERSION}")
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


class NativeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise NativeError(f"libho_b200 error {rc}: {lib().ho_last_error().decode()}")
