"""Import alias for the product package.

The package directory is ``gb-nerf_b200/`` (not a Python identifier), so ``import gbnerf_b200`` loads it from
there and installs it under this module name, sub-modules included (``gbnerf_b200.render`` …).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gb-nerf_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
