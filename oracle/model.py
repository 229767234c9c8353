"""CPU oracle: the two pruning models and their teachers as pure functions of a state dict.
TEST INFRASTRUCTURE ONLY (see oracle/ops.py header for who may import this and how it is pinned).

The state dict uses the reference's parameter names: patch_embed.proj, cls_token,
pos_embed, blocks.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}, norm, head,
score_predictor.{p}.{in_conv,out_conv}.{idx}.  Citations are file:line under /root/reference/.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn.functional as F

from . import ops

INIT_N = 14 * 14  # hard-coded in the reference: vit_models/dynamic_vit.py:828, default_dynamic_vit.py:446


@dataclass
class VitCfg:
    embed_dim: int = 384
    depth: int = 12
    num_heads: int = 6
    patch_size: int = 16
    num_classes: int = 1000
    pruning_loc: List[int] = field(default_factory=lambda: [3, 6, 9])
    token_ratio: List[float] = field(default_factory=lambda: [0.7, 0.49, 0.343])
    # Variant B predictor options (vit_models/dynamic_vit.py:648-653)
    small_predictor: bool = False
    predictor_bn: bool = False
    predictor_loss_type: str = "kl_div"
    patch_score_threshold: Optional[float] = None


def embed(sd, cfg, img):
    """Conv patch embedding as a matmul + CLS + position embedding
    (vit_models/dynamic_vit.py:286-303, :816-824)."""
    B, C, Hh, Ww = img.shape
    p = cfg.patch_size
    gh, gw = Hh // p, Ww // p
    patches = img.view(B, C, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, C * p * p)
    w = sd["patch_embed.proj.weight"].reshape(cfg.embed_dim, -1)
    x = F.linear(patches, w, sd["patch_embed.proj.bias"])
    x = torch.cat([sd["cls_token"].expand(B, -1, -1), x], dim=1)
    return x + sd["pos_embed"]


def block(sd, cfg, i, x, policy=None, want_cls_attn=False):
    """Pre-LN transformer block (vit_models/dynamic_vit.py:263-283, default_dynamic_vit.py:234-237);
    block LayerNorms use eps=1e-6 (dynamic_vit.py:676)."""
    pre = f"blocks.{i}"
    h = F.layer_norm(x, (cfg.embed_dim,), sd[pre + ".norm1.weight"], sd[pre + ".norm1.bias"], 1e-6)
    res = ops.attention(h, sd[pre + ".attn.qkv.weight"], sd.get(pre + ".attn.qkv.bias"),
                        sd[pre + ".attn.proj.weight"], sd[pre + ".attn.proj.bias"], cfg.num_heads,
                        policy=policy, return_cls_attn=want_cls_attn)
    cls_attn = None
    if want_cls_attn:
        res, cls_attn = res
    x = x + res
    h = F.layer_norm(x, (cfg.embed_dim,), sd[pre + ".norm2.weight"], sd[pre + ".norm2.bias"], 1e-6)
    h = F.linear(F.gelu(F.linear(h, sd[pre + ".mlp.fc1.weight"], sd[pre + ".mlp.fc1.bias"])),
                 sd[pre + ".mlp.fc2.weight"], sd[pre + ".mlp.fc2.bias"])
    x = x + h
    return (x, cls_attn) if want_cls_attn else x


def _head(sd, cfg, x):
    x = F.layer_norm(x, (cfg.embed_dim,), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return F.linear(x[:, 0], sd["head.weight"], sd["head.bias"]), x[:, 1:]


def variant_a_eval(sd, cfg, img):
    """DynamicViT inference (vit_models/default_dynamic_vit.py:435-487, eval branch :461-468).
    Returns dict(logits, kept=[(B,K_s) int64 per stage, stage-relative, score order], scores=[(B,N_s,2)])."""
    x = embed(sd, cfg, img)
    B = x.shape[0]
    prev = torch.ones(B, INIT_N, 1, dtype=x.dtype, device=x.device)
    kept_all, score_all = [], []
    p = 0
    for i in range(cfg.depth):
        if i in cfg.pruning_loc:
            logp = ops.predictor_a(sd, f"score_predictor.{p}", x[:, 1:], prev).reshape(B, -1, 2)
            k = ops.num_keep(INIT_N, cfg.token_ratio[p])
            kept, _ = ops.select_topk(logp[:, :, 0], k, ops.ORDER_SCORE_DESC)
            x = ops.gather_tokens_with_cls(x, kept)
            prev = ops.batch_index_select(prev, kept)
            kept_all.append(kept)
            score_all.append(logp)
            p += 1
        x = block(sd, cfg, i, x)
    logits, feats = _head(sd, cfg, x)
    return dict(logits=logits, features=feats, kept=kept_all, scores=score_all)


def variant_a_train(sd, cfg, img, gumbels):
    """DynamicViT training forward with injected Gumbel noise (default_dynamic_vit.py:453-459, :470-472).
    The sequence never shrinks; pruning is simulated through the policy in every later block.
    Returns dict(logits, features, final_decision (B,196,1), decisions=[(B,196)], scores)."""
    x = embed(sd, cfg, img)
    B = x.shape[0]
    prev = torch.ones(B, INIT_N, 1, dtype=x.dtype, device=x.device)
    policy = torch.ones(B, INIT_N + 1, 1, dtype=x.dtype, device=x.device)
    decisions, score_all = [], []
    p = 0
    for i in range(cfg.depth):
        if i in cfg.pruning_loc:
            logp = ops.predictor_a(sd, f"score_predictor.{p}", x[:, 1:], prev).reshape(B, -1, 2)
            hard, _ = ops.gumbel_keep_decision(logp, gumbels[p], prev)
            decisions.append(hard.reshape(B, INIT_N))
            policy = torch.cat([torch.ones(B, 1, 1, dtype=hard.dtype, device=hard.device), hard], dim=1)
            prev = hard
            score_all.append(logp)
            p += 1
        x = block(sd, cfg, i, x, policy=policy)
    logits, feats = _head(sd, cfg, x)
    return dict(logits=logits, features=feats, final_decision=prev, decisions=decisions, scores=score_all)


def variant_b_forward(sd, cfg, img, training=False):
    """Dense2Sparse forward, top-k mode (vit_models/dynamic_vit.py:814-1015 with
    topk_selection=True, patch_score_threshold=None): identical hard selection in train and eval
    (:857-865), gather via ascending kept indices (:907-912 / :954-960), CLS attention row of every
    block collected (:925, :975, :986-989).
    Returns dict(logits, features, cls_attns=[(B,H,T_i-1)], pred_logits=[(B,N_s)], pred_scores, kept, dropped)."""
    x = embed(sd, cfg, img)
    kept_all, drop_all, logit_all, prob_all, cls_attns = [], [], [], [], []
    p = 0
    for i in range(cfg.depth):
        if i in cfg.pruning_loc:
            scores, probs = ops.predictor_b(sd, f"score_predictor.{p}", x[:, 1:], cfg.small_predictor,
                                            cfg.predictor_bn, cfg.predictor_loss_type, training)
            k = ops.num_keep(INIT_N, cfg.token_ratio[p])
            kept, dropped = ops.select_topk(probs, k, ops.ORDER_INDEX_ASC)
            x = ops.gather_tokens_with_cls(x, kept)
            kept_all.append(kept)
            drop_all.append(dropped)
            logit_all.append(scores)
            prob_all.append(probs)
            p += 1
        x, ca = block(sd, cfg, i, x, want_cls_attn=True)
        cls_attns.append(ca[:, :, 1:])
    logits, feats = _head(sd, cfg, x)
    return dict(logits=logits, features=feats, cls_attns=cls_attns, pred_logits=logit_all,
                pred_scores=prob_all, kept=kept_all, dropped=drop_all)


def variant_b_threshold_train(sd, cfg, img):
    """Dense2Sparse dynamic keep-ratio training branch (vit_models/dynamic_vit.py:880-894, :982-983):
    a 0/1 keep mask from the cumulative score threshold drives softmax_with_policy in the pruning
    block and every later block; no token is removed.  (The inference branch is broken in the
    reference: `score` undefined at :936.)"""
    x = embed(sd, cfg, img)
    B = x.shape[0]
    keep_mask = torch.ones(B, INIT_N + 1, dtype=x.dtype, device=x.device)
    logit = None
    p = 0
    for i in range(cfg.depth):
        if i in cfg.pruning_loc:
            logit, probs = ops.predictor_b(sd, f"score_predictor.{p}", x[:, 1:], cfg.small_predictor,
                                           cfg.predictor_bn, cfg.predictor_loss_type, True)
            m = ops.threshold_keep_mask(probs.detach(), cfg.patch_score_threshold)
            keep_mask = torch.cat([torch.ones(B, 1, dtype=x.dtype, device=x.device), m.to(x.dtype)], dim=1)
            p += 1
        x = block(sd, cfg, i, x, policy=keep_mask.unsqueeze(-1))
    logits, feats = _head(sd, cfg, x)
    return dict(logits=logits, features=feats, pred_logits=logit, keep_mask=keep_mask[:, 1:])


def variant_b_threshold_eval_intended(sd, cfg, img):
    """What the inference branch of the dynamic keep-ratio mode evidently intends (vit_models/dynamic_vit.py:935-949): the
    lines themselves cannot execute (`score` is never assigned, :936; a float mask is used as an index, :947), so nothing can
    be pinned against the reference here -- PARITY UNPINNED for this function.  Intent, restated: the cumulative-score
    threshold of the training branch on the eval predictor scores, then the dropped tokens are removed (variable length:
    one image per batch, mask_predictor.py:249-254); the pruning block and the later ones run without a policy."""
    assert img.shape[0] == 1, "variable-length pruning: one image per batch"
    x = embed(sd, cfg, img)
    kept_all, p = [], 0
    for i in range(cfg.depth):
        if i in cfg.pruning_loc:
            _, probs = ops.predictor_b(sd, f"score_predictor.{p}", x[:, 1:], cfg.small_predictor, cfg.predictor_bn,
                                       cfg.predictor_loss_type, False)
            m = ops.threshold_keep_mask(probs, cfg.patch_score_threshold)
            kept = torch.nonzero(m[0]).flatten().unsqueeze(0)
            x = ops.gather_tokens_with_cls(x, kept)
            kept_all.append(kept)
            p += 1
        x = block(sd, cfg, i, x)
    logits, feats = _head(sd, cfg, x)
    return dict(logits=logits, features=feats, kept=kept_all)


def teacher_forward(sd, cfg, img, want_cls_attn=True):
    """Unpruned ViT teacher (vit_models/dynamic_vit.py:1150-1176; default_dynamic_vit.py:581-598).
    Returns (logits, tokens (B,196,D), cls_attn (B,depth,H,197) or None)."""
    x = embed(sd, cfg, img)
    rows = []
    for i in range(cfg.depth):
        if want_cls_attn:
            x, ca = block(sd, cfg, i, x, want_cls_attn=True)
            rows.append(ca)
        else:
            x = block(sd, cfg, i, x)
    logits, feats = _head(sd, cfg, x)
    return logits, feats, (torch.stack(rows, dim=1) if want_cls_attn else None)
