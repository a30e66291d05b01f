"""Place the UNMODIFIED reference under the git-ignored ``baseline/_ref/`` so that it travels to
the GPU box with the gpurun snapshot (``/root/reference`` does not exist there).

    python tools/install_reference.py            # from /root/reference

The contract's ``pip install --no-index --no-build-isolation --target baseline/_ref
/root/reference`` fails in this image: the reference's build backend is hatchling, which is not
in the offline wheelhouse (``ModuleNotFoundError: No module named 'hatchling'``, with and without
``--no-deps``).  The reference is pure Python with no build step, so a wheel install would place
exactly ``hyperbolic_optics/*.py`` + ``material_params.json`` -- this script copies those files
verbatim instead, plus the reference's golden snapshots (``tests/golden``) for the parity tests.
Nothing under ``baseline/_ref`` is tracked, imported by the product, or modified.
"""

from __future__ import annotations

import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DEST = ROOT / "baseline" / "_ref"


def install(source="/root/reference", force=False):
    src = Path(source)
    pkg = src / "hyperbolic_optics"
    if not pkg.is_dir():
        return False
    stamp = DEST / "hyperbolic_optics" / "__init__.py"
    if stamp.exists() and not force and stamp.stat().st_mtime >= (pkg / "__init__.py").stat().st_mtime:
        return True
    DEST.mkdir(parents=True, exist_ok=True)
    for name, sub in (("hyperbolic_optics", pkg), ("reference_golden", src / "tests" / "golden")):
        target = DEST / name
        if target.exists():
            shutil.rmtree(target)
        if sub.is_dir():
            shutil.copytree(sub, target, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    (DEST / "README").write_text(
        "Verbatim copy of /root/reference/hyperbolic_optics (v0.3.0) and its tests/golden, made by\n"
        "tools/install_reference.py.  Git-ignored; used only by bench.py --impl reference, the\n"
        "cpu_baseline leg and the drop-in tests.\n")
    return True


if __name__ == "__main__":
    ok = install(*(sys.argv[1:2] or ["/root/reference"]), force=True)
    print("installed" if ok else "reference not found", DEST)
