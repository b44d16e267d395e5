"""Loader: makes the hyphenated package directory `node-cueball_b200/` importable as
`cueball_b200` (a Python module name cannot contain '-')."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "node-cueball_b200")
_spec = importlib.util.spec_from_file_location("cueball_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cueball_b200"] = _mod
_spec.loader.exec_module(_mod)
