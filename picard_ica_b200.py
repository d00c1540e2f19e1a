"""Import shim: the package directory is named `picard-ica_b200` (not an importable identifier), so this
module makes `import picard_ica_b200` resolve to it (sub-modules included, via __path__)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "picard-ica_b200")
_spec = _ilu.spec_from_file_location("picard_ica_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["picard_ica_b200"] = _mod
_spec.loader.exec_module(_mod)
