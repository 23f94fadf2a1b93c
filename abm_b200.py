"""Import shim: loads the package directory ``animal-movement-abm_b200/`` under the importable name ``abm_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "animal-movement-abm_b200")
_spec = importlib.util.spec_from_file_location("abm_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["abm_b200"] = _mod
_spec.loader.exec_module(_mod)
