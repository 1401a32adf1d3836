"""Import shim: the package directory is named `spherical-sfm_b200/` (it carries the reference's
hyphen), which is not a valid Python identifier.  `import spherical_sfm_b200` loads that directory
as a regular package under this name."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spherical-sfm_b200")
_spec = importlib.util.spec_from_file_location("spherical_sfm_b200", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["spherical_sfm_b200"] = _mod
_spec.loader.exec_module(_mod)
