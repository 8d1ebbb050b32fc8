"""Loader for the package directory `hevc-complexity-reduction_b200/` (a hyphen is not importable).

    import ethcnn_b200 as eb          # -> module hevc_complexity_reduction_b200
    net = eb.EthCnn(model_dir, ...)
"""
import importlib.util
import os
import sys

_NAME = "hevc_complexity_reduction_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hevc-complexity-reduction_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
PACKAGE_DIR = _DIR
