"""Drop-in installation behind the reference's own Python modules.

    import vit_models.dynamic_vit as dvit, vit_models.default_dynamic_vit as ddvit, vit_models.peturbed_topk as ptk
    from d2s import pkg;  pkg.patch.install(dvit=dvit, ddvit=ddvit, ptopk=ptk)

After install(), models built from the reference classes (and their checkpoints) run their hot path on the
d2s kernels: the reference's callers (train.py:43, evaluate.py:35-37) are untouched.  uninstall() restores
the original attributes.  Only attribute surfaces listed in SURVEY.md section 8b are replaced.
"""
from . import engine, ops
from .perturbed_topk import PerturbedTopK as _PTK

_saved = []


def _swap(obj, name, new):
    _saved.append((obj, name, getattr(obj, name)))
    setattr(obj, name, new)


def install(dvit=None, ddvit=None, ptopk=None):
    """Patch whichever reference modules are given (already-imported module objects)."""
    if ddvit is not None:   # Variant A, vit_models/default_dynamic_vit.py
        _swap(ddvit, "batch_index_select", ops.batch_index_select)
        _swap(ddvit.Attention, "softmax_with_policy", lambda self, attn, policy, eps=1e-6: engine.softmax_with_policy(attn, policy, eps))
        _swap(ddvit.Attention, "forward", lambda self, x, policy: engine.attention_forward(self, x, policy))
        _swap(ddvit.Block, "forward", lambda self, x, policy=None: engine.block_forward(self, x, policy))
        _swap(ddvit.PatchEmbed, "forward", lambda self, x: engine.patch_embed_forward(self, x))
        _swap(ddvit.PredictorLG, "forward", lambda self, x, policy: engine.predictor_a_forward(self, x, policy))
        _swap(ddvit.DefaultVisionTransformerDiffPruning, "forward", lambda self, x: engine.variant_a_forward(self, x))
        _swap(ddvit.DefaultVisionTransformerTeacher, "forward",
              lambda self, x: engine.teacher_forward(self, x, with_cls_attn=False))
    if dvit is not None:    # Variant B, vit_models/dynamic_vit.py
        _swap(dvit, "batch_index_select", ops.batch_index_select)
        _swap(dvit.Attention, "softmax_with_policy", lambda self, attn, policy, eps=1e-6: engine.softmax_with_policy(attn, policy, eps))
        _swap(dvit.Attention, "forward",
              lambda self, x, policy, return_cls_attn=False: engine.attention_forward(self, x, policy, return_cls_attn))
        _swap(dvit.Block, "forward",
              lambda self, x, policy=None, return_cls_attn=False: engine.block_forward(self, x, policy, return_cls_attn))
        _swap(dvit.PatchEmbed, "forward", lambda self, x: engine.patch_embed_forward(self, x))
        _swap(dvit.PredictorLG, "forward",
              lambda self, x, policy=None, current_sigma=0.0005, cls_attn=None:
              engine.predictor_b_forward(self, x, policy, current_sigma, cls_attn))
        _swap(dvit.VisionTransformerDiffPruning, "forward",
              lambda self, x, stacked_cls_attn_weights=None: engine.variant_b_forward(self, x, stacked_cls_attn_weights))
        _swap(dvit.VisionTransformerDiffPruning, "forward_cls_attn", lambda self, x: engine.variant_b_forward_cls_attn(self, x))
        _swap(dvit.VisionTransformerTeacher, "forward", lambda self, x: engine.teacher_forward(self, x, with_cls_attn=True))
        _swap(dvit.VisionTransformerTeacher, "forward_cls_attention",
              lambda self, x: engine.teacher_forward_cls_attention(self, x))
        if hasattr(dvit, "PerturbedTopK"):
            _swap(dvit, "PerturbedTopK", _PTK)
    if ptopk is not None:   # vit_models/peturbed_topk.py
        _swap(ptopk.PerturbedTopK, "__call__",
              lambda self, x, current_sigma=0.05, noise=None, seed=None:
              ops.perturbed_topk(x, self.k, self.num_samples, current_sigma, noise=noise, seed=seed))


def uninstall():
    while _saved:
        obj, name, old = _saved.pop()
        setattr(obj, name, old)
