"""Variant A -- upstream DynamicViT as shipped in the reference's vit_models/default_dynamic_vit.py.
Same class names, constructor arguments, return tuples and state-dict keys; forwards run on the d2s kernels."""
import torch
import torch.nn as nn

from . import engine, ops
from .layers import _VitBackbone

batch_index_select = ops.batch_index_select  # default_dynamic_vit.py:37-53


class PredictorLG(nn.Module):
    """Score predictor (default_dynamic_vit.py:304-330): (x (B,N,D), policy (B,N,1)) -> log-probs (B,N,2)."""

    def __init__(self, embed_dim=384):
        super().__init__()
        self.in_conv = nn.Sequential(nn.LayerNorm(embed_dim), nn.Linear(embed_dim, embed_dim), nn.GELU())
        self.out_conv = nn.Sequential(
            nn.Linear(embed_dim, embed_dim // 2), nn.GELU(),
            nn.Linear(embed_dim // 2, embed_dim // 4), nn.GELU(),
            nn.Linear(embed_dim // 4, 2), nn.LogSoftmax(dim=-1))

    def forward(self, x, policy):
        return engine.predictor_a_forward(self, x, policy)


class DefaultVisionTransformerDiffPruning(_VitBackbone):
    """default_dynamic_vit.py:333-487.  Training returns (logits, features, final_decision, [decisions]) when
    distill=True, else (logits, [decisions]); eval returns logits.  After an eval forward,
    `kept_token_indices` holds the per-stage kept indices (stage-relative, descending-score order)."""

    def __init__(self, *args, pruning_loc=None, token_ratio=None, distill=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.score_predictor = nn.ModuleList([PredictorLG(self.embed_dim) for _ in range(len(pruning_loc))])
        self.distill = distill
        self.pruning_loc = pruning_loc
        self.token_ratio = token_ratio
        self.kept_token_indices = []
        self._finish_init()

    def forward(self, x):
        return engine.variant_a_forward(self, x)


class DefaultVisionTransformerTeacher(_VitBackbone):
    """default_dynamic_vit.py:489-598: unpruned ViT returning (logits, tokens)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._finish_init()

    def forward(self, x):
        return engine.teacher_forward(self, x, with_cls_attn=False)


_ARCH = {  # factory hyper-parameters (default_dynamic_vit.py:641-781)
    "tiny": dict(patch_size=16, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4, qkv_bias=True),
    "small": dict(patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True),
    "base": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True),
}


def _student(arch, pruning_locs, keep_ratios, state_dict=None, **kwargs):
    model = DefaultVisionTransformerDiffPruning(pruning_loc=pruning_locs, token_ratio=keep_ratios, distill=True,
                                                **_ARCH[arch], **kwargs)
    if state_dict is not None:  # the reference downloads DeiT weights here; offline callers pass them in
        model.load_state_dict(state_dict.get("model", state_dict), strict=False)
    return model


def default_dynamic_vit_tiny_patch16_224_student(pruning_locs, keep_ratios, **kwargs):
    return _student("tiny", pruning_locs, keep_ratios, **kwargs)


def default_dynamic_vit_small_patch16_224_student(pruning_locs, keep_ratios, **kwargs):
    return _student("small", pruning_locs, keep_ratios, **kwargs)


def default_dynamic_vit_base_patch16_224_student(pruning_locs, keep_ratios, **kwargs):
    return _student("base", pruning_locs, keep_ratios, **kwargs)


def default_dynamic_vit_teacher(arch="small", state_dict=None, **kwargs):
    model = DefaultVisionTransformerTeacher(**_ARCH[arch], **kwargs)
    if state_dict is not None:
        model.load_state_dict(state_dict.get("model", state_dict), strict=False)
    return model
