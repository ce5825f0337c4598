"""Import shim: the package directory is named ``cs420-ray-tracer_b200`` (a hyphen is not a
valid module name), so ``import rtb200`` loads it as module ``cs420_ray_tracer_b200`` and
re-exports its public names."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cs420-ray-tracer_b200")
_NAME = "cs420_ray_tracer_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
globals().update({k: getattr(pkg, k) for k in dir(pkg) if not k.startswith("__")})
