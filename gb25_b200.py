"""Import shim: the package directory is named ``gb-25_b200`` (after the reference repo), which is
not a valid Python identifier.  ``import gb25_b200`` loads that directory as a regular package."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gb-25_b200")
_spec = importlib.util.spec_from_file_location(
    "gb25_b200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_m = importlib.util.module_from_spec(_spec)
sys.modules["gb25_b200"] = _m
_spec.loader.exec_module(_m)
