"""Import alias for the product package.

The package directory is `holistic-robot-pose-estimation-study_b200/`, which is not a valid Python identifier; this
module loads it under the importable name `hrp_b200` (sub-modules resolve through the package's own __path__).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "holistic-robot-pose-estimation-study_b200")
_spec = importlib.util.spec_from_file_location("hrp_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hrp_b200"] = _mod
_spec.loader.exec_module(_mod)
