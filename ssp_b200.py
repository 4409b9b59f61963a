"""Import shim: `import ssp_b200` loads the package directory `semantic-superpoint_b200/` (whose name is not a
valid Python identifier) under the module name `ssp_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "semantic-superpoint_b200")
_spec = importlib.util.spec_from_file_location(
    "ssp_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ssp_b200"] = _mod
_spec.loader.exec_module(_mod)
