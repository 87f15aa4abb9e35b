"""Import shim: the package directory `volume-rendering-based-on-distribution-data_b200/` has
hyphens in its name, so it cannot be imported by name.  `import vrdd_b200` loads it."""
import importlib.util
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "volume-rendering-based-on-distribution-data_b200")
_spec = importlib.util.spec_from_file_location("vrdd_b200", os.path.join(_PKG, "__init__.py"),
                                               submodule_search_locations=[_PKG])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vrdd_b200"] = _mod
_spec.loader.exec_module(_mod)
