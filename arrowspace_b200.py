"""Import shim: loads the package that lives in ``arrowspace-rs_b200/`` (hyphenated directory)
under the importable name ``arrowspace_b200``."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "arrowspace-rs_b200"
_spec = importlib.util.spec_from_file_location(
    "arrowspace_b200", _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["arrowspace_b200"] = _mod
_spec.loader.exec_module(_mod)
