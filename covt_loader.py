"""Imports the package directory `cov-tiles_b200/` (not a valid identifier) under the module name cov_tiles_b200."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if "cov_tiles_b200" in sys.modules:
        return sys.modules["cov_tiles_b200"]
    pkg = os.path.join(_ROOT, "cov-tiles_b200")
    spec = importlib.util.spec_from_file_location("cov_tiles_b200", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["cov_tiles_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
