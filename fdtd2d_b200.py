"""Import shim: the package directory is `fdtd-2d_b200/` (a hyphen cannot appear in a Python module
name), so `import fdtd2d_b200` loads that directory as the package `fdtd2d_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fdtd-2d_b200")
_spec = importlib.util.spec_from_file_location(
    "fdtd2d_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fdtd2d_b200"] = _mod
_spec.loader.exec_module(_mod)
