"""Import the REAL reference (`/root/reference/python-src/main.py`) -- test infrastructure only.

The reference module has an import-time side effect (main.py:7-9 removes and
recreates ``./frames`` in the current directory) and imports ``matplotlib.cm``
(main.py:3), which is not installed here.  We therefore import it from a
scratch working directory with an empty ``matplotlib`` stub in ``sys.modules``.
Only ``capture_snapshot`` (main.py:171) touches matplotlib, and we never call it.

This loader exists so that ``make_golden.py`` can produce golden vectors and so
that CPU tests can cross-check the oracle against the reference *when the
reference tree is present*.  `/root/reference` does not exist on the GPU box;
nothing marked ``gpu`` and nothing in bench.py/smoke() may call this.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("FDTD2D_REFERENCE_ROOT", "/root/reference")
_cached = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "python-src", "main.py"))


def load_reference_main():
    """Return the reference's ``main`` module (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl_cm = types.ModuleType("matplotlib.cm")
        mpl.cm = mpl_cm
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.cm"] = mpl_cm
    src = os.path.join(REFERENCE_ROOT, "python-src")
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="fdtd2d_ref_cwd_")
    saved_main = sys.modules.pop("main", None)
    try:
        os.chdir(scratch)  # main.py:7-9 rmtree/mkdir "frames" happens here, not in the repo
        sys.path.insert(0, src)
        import importlib

        mod = importlib.import_module("main")
    finally:
        os.chdir(cwd)
        if src in sys.path:
            sys.path.remove(src)
        # keep the module private: do not leave a top-level "main" around
        sys.modules.pop("main", None)
        if saved_main is not None:
            sys.modules["main"] = saved_main
    _cached = mod
    return mod
