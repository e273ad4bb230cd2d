"""Import shim: the package directory is named ``pqp-for-mpc_b200`` (after the reference repository),
which is not a Python identifier; this module loads it under the importable name ``pqp_for_mpc_b200``."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pqp-for-mpc_b200")
_spec = _u.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
