"""Import alias: the package directory is named ``sp-nerf_b200`` (not a valid Python identifier),
so ``import spnerf_b200`` loads it from there and hands back the real package object."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sp-nerf_b200")
_spec = importlib.util.spec_from_file_location(
    "spnerf_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["spnerf_b200"] = _pkg
_spec.loader.exec_module(_pkg)
