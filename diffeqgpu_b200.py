"""Import shim: the package directory is named `diffeqgpu.jl_b200` (not a valid Python
identifier), so load it by path and expose it as the module `diffeqgpu_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "diffeqgpu.jl_b200"
_spec = importlib.util.spec_from_file_location(
    "diffeqgpu_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["diffeqgpu_b200"] = _mod
_spec.loader.exec_module(_mod)
