"""dense2sparse-vit_b200: the B200 (sm_100a) token-sparsification hot path of Dense2Sparse-ViT / DynamicViT.

The directory name is not a Python identifier; import it through the shim at the repo root:

    import d2s                       # registers this package as `dense2sparse_vit_b200`
    from d2s import pkg, ops         # pkg.variant_a, pkg.variant_b, pkg.perturbed_topk, pkg.patch, ...

csrc/ holds the CUDA kernels and the C ABI (include/d2s.h); ops.py wraps them for torch tensors; engine.py /
layers.py / variant_a.py / variant_b.py / perturbed_topk.py mirror the reference's Python interface;
patch.py installs the kernels behind the reference's own modules.
"""
from . import _lib, ops, engine, layers, perturbed_topk, variant_a, variant_b, patch, runner, losses  # noqa: F401

__all__ = ["_lib", "ops", "engine", "layers", "perturbed_topk", "variant_a", "variant_b", "patch", "runner", "losses"]
