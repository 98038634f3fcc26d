"""Import shim: ``import rfb200`` loads the package in ``recursivefactorization.jl_b200/``.

The package directory carries the project's name, which contains a dot and therefore cannot be
imported with a plain ``import`` statement.
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "recursivefactorization.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "rfb200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rfb200"] = _mod
_spec.loader.exec_module(_mod)
