"""Import shim: the product package lives in `elliptic-gnn-project_b200/` (a name Python
cannot import directly because of the hyphens); `import egnn_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "elliptic-gnn-project_b200")
_spec = importlib.util.spec_from_file_location(
    "egnn_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["egnn_b200"] = _mod
_spec.loader.exec_module(_mod)
