"""hyperbolic_optics_b200 -- B200-native batched 4x4 Berreman reflection engine.

Drop-in for the hot path of hyperbolic-optics 0.3.0 (``Structure.execute`` ->
``r_pp, r_ps, r_sp, r_ss`` -> Mueller matrix): Python host code plans the stack and calls
hand-written sm_100a CUDA kernels through the C ABI in ``include/ho_b200.h``.
"""

from .dropin import install, installed, uninstall
from .fields import FieldProfile
from .materials import create_material, list_materials
from .mueller import Mueller
from .jones import Jones, compose_jones
from .scenario import ScenarioSetup
from .structure import Structure
from .sweep import ThicknessSweep

__version__ = "0.1.0"
__all__ = ["Structure", "ScenarioSetup", "Mueller", "Jones", "compose_jones", "FieldProfile", "ThicknessSweep", "create_material", "list_materials",
           "install", "uninstall", "installed", "__version__"]
