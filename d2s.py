"""Import shim: the product package lives in `dense2sparse-vit_b200/` (not a Python identifier).
`import d2s` registers it as `dense2sparse_vit_b200` and re-exports its submodules."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "dense2sparse-vit_b200")
_NAME = "dense2sparse_vit_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
ops = pkg.ops
engine = pkg.engine
variant_a = pkg.variant_a
variant_b = pkg.variant_b
perturbed_topk = pkg.perturbed_topk
patch = pkg.patch
runner = pkg.runner
losses = pkg.losses
_lib = pkg._lib
