"""Route B of INTEGRATION.md as code: install the CUDA-backed ``Structure`` *as* the reference's.

The reference has no plugin interface; its consumer modules bind the class name at import time
(``mueller.py:16``, ``jones.py:33``, ``fields.py:27``, ``sweep.py:37``), ``compose_jones`` tests
``isinstance(element, Structure)`` (``jones.py:404``) and ``ThicknessSweep`` instantiates
``Structure()`` itself (``sweep.py:90``) -- SURVEY 8b.  ``install()`` therefore rebinds that name
in every reference module that holds it; the reference's ``Mueller``, ``Jones``, ``compose_jones``,
``FieldProfile`` and ``ThicknessSweep`` then run unchanged on results of the CUDA engine
(``transfer_matrix`` / ``layers[i].matrix`` / ``.profile`` are materialised lazily on first read).

    import hyperbolic_optics, hyperbolic_optics_b200
    hyperbolic_optics_b200.install()            # hyperbolic_optics.Structure is now B200-backed
    ...
    hyperbolic_optics_b200.uninstall()          # restores the NumPy class
"""

from __future__ import annotations

import importlib
import sys

from .structure import Structure

_MODULES = ("hyperbolic_optics", "hyperbolic_optics.structure", "hyperbolic_optics.mueller", "hyperbolic_optics.jones",
            "hyperbolic_optics.fields", "hyperbolic_optics.sweep", "hyperbolic_optics.plots")
_saved = {}


def install(structure_cls=None):
    """Rebind ``Structure`` in the importable reference package ``hyperbolic_optics``; returns the
    list of patched module names.  Raises ``ImportError`` if the reference is not importable."""
    cls = structure_cls or Structure
    importlib.import_module("hyperbolic_optics")
    patched = []
    for name in _MODULES:
        try:
            mod = importlib.import_module(name)
        except ImportError:   # optional module (plots needs matplotlib)
            continue
        current = getattr(mod, "Structure", None)
        if current is None:
            continue
        if name not in _saved:
            _saved[name] = current
        setattr(mod, "Structure", cls)
        patched.append(name)
    return patched


def uninstall():
    """Undo ``install()``."""
    for name, original in list(_saved.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            setattr(mod, "Structure", original)
        del _saved[name]


def installed():
    return bool(_saved)
