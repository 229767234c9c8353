"""CPU oracle for the token-sparsification hot path of Dense2Sparse-ViT  --  TEST INFRASTRUCTURE ONLY.

This file restates, in this repo's own words, what the reference computes on the hot path.  It is the
checker for the CUDA kernels: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product (`dense2sparse-vit_b200/`) never does.

Parity pinning: the reference ships no tests or golden vectors, so every function
here is pinned against outputs of the UNMODIFIED reference executed in the build container
(`tests/golden/make_goldens.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).
Third-party arithmetic the path relies on (torch.nn.functional.gumbel_softmax, torch.sort tie order)
is pinned the same way, with the tie rule of SURVEY.md section 7 (hard part 1): compare in fp32,
lower token index first.

All tensors are CPU torch tensors; float math is fp32 unless a function says otherwise.
Citations are file:line under /root/reference/.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# selection
# ----------------------------------------------------------------------------------------------

ORDER_INDEX_ASC = 0   # Variant B: kept indices re-sorted ascending (vit_models/dynamic_vit.py:858-862)
ORDER_SCORE_DESC = 1  # Variant A: kept indices stay in descending-score order (vit_models/default_dynamic_vit.py:461-463)


def num_keep(init_n: int, ratio: float) -> int:
    """K = int(init_n * ratio) with truncation (vit_models/dynamic_vit.py:852, default_dynamic_vit.py:462)."""
    return int(init_n * ratio)


def select_topk(score: torch.Tensor, k: int, order: int):
    """score (B, N) fp32 -> (kept (B,k) int64, dropped (B,N-k) int64).

    Reference: argsort(descending) then a cut at k; Variant B sorts both halves ascending
    (vit_models/dynamic_vit.py:858-862), Variant A keeps score order and has no dropped list
    (vit_models/default_dynamic_vit.py:461-463).  torch.argsort is not stable; the rule adopted for
    the whole project is the stable one: equal scores -> lower index first; NaN sorts as largest
    (torch.sort semantics); -0.0 == +0.0.
    """
    assert score.dim() == 2
    srt = torch.sort(score.float(), dim=1, descending=True, stable=True).indices
    kept, dropped = srt[:, :k], srt[:, k:]
    if order == ORDER_INDEX_ASC:
        kept = torch.sort(kept, dim=1).values
        dropped = torch.sort(dropped, dim=1).values
    return kept.contiguous(), dropped.contiguous()


def threshold_keep_mask(score: torch.Tensor, threshold: float) -> torch.Tensor:
    """Dynamic keep-ratio mode, training branch (vit_models/dynamic_vit.py:880-890): ascending sort,
    cumulative sum, keep tokens whose cumulative mass exceeds the threshold.  Returns (B,N) bool."""
    val, idx = torch.sort(score.float(), dim=1, stable=True)
    keep_sorted = torch.cumsum(val, dim=-1) > threshold
    mask = torch.zeros_like(keep_sorted)
    mask.scatter_(1, idx, keep_sorted)
    return mask


# ----------------------------------------------------------------------------------------------
# gather / scatter of kept tokens
# ----------------------------------------------------------------------------------------------

def batch_index_select(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Per-image row gather (vit_models/default_dynamic_vit.py:37-53).  x (B,N,C) or (B,N); idx (B,K) int64."""
    if x.dim() == 3:
        b = torch.arange(x.shape[0], device=x.device).view(-1, 1)
        return x[b, idx]
    if x.dim() == 2:
        return torch.gather(x, 1, idx)
    raise NotImplementedError


def gather_tokens_with_cls(x: torch.Tensor, kept: torch.Tensor) -> torch.Tensor:
    """x (B,T,D) incl. CLS at row 0, kept (B,K) spatial indices -> (B,K+1,D): CLS then kept+1
    (vit_models/dynamic_vit.py:907-912 / 954-960, default_dynamic_vit.py:464-466)."""
    B = x.shape[0]
    rows = torch.cat([torch.zeros(B, 1, dtype=kept.dtype, device=kept.device), kept + 1], dim=1)
    return batch_index_select(x, rows)


def scatter_tokens_bwd(gout: torch.Tensor, kept: torch.Tensor, t_in: int) -> torch.Tensor:
    """Backward of gather_tokens_with_cls: (B,K+1,D) -> (B,t_in,D), zeros for dropped rows
    (autograd of torch.gather at vit_models/dynamic_vit.py:912; indices are unique per image)."""
    B, _, D = gout.shape
    rows = torch.cat([torch.zeros(B, 1, dtype=kept.dtype, device=kept.device), kept + 1], dim=1)
    gx = torch.zeros(B, t_in, D, dtype=gout.dtype, device=gout.device)
    gx[torch.arange(B, device=gout.device).view(-1, 1), rows] = gout
    return gx


# ----------------------------------------------------------------------------------------------
# Gumbel keep decision (Variant A training)
# ----------------------------------------------------------------------------------------------

def gumbel_keep_decision(logp: torch.Tensor, gumbel: torch.Tensor, prev_decision: torch.Tensor):
    """logp (B,N,2) log-probs, gumbel (B,N,2) injected noise, prev_decision (B,N,1) -> hard keep (B,N,1).

    vit_models/default_dynamic_vit.py:454 calls torch.nn.functional.gumbel_softmax(hard=True)
    (torch/nn/functional.py, third-party): y = softmax(logp + g); hard = onehot(argmax y);
    value = hard - y + y (exactly 0/1 for 2 classes); class 0 is "keep"; result is multiplied by the
    previous decision.  Also returns y[...,0] (the soft keep probability that carries the gradient).
    """
    y = torch.softmax((logp + gumbel) / 1.0, dim=-1)
    hard = torch.zeros_like(y).scatter_(-1, y.max(-1, keepdim=True)[1], 1.0)
    out = (hard - y) + y
    return out[..., 0:1] * prev_decision, y[..., 0]


def gumbel_keep_decision_bwd(gout: torch.Tensor, ysoft0: torch.Tensor, prev_decision: torch.Tensor):
    """d loss / d logp for the straight-through estimator: only y_soft carries gradient.
    gout (B,N,1) -> (B,N,2).  With y=(y0,1-y0): dL/dx0 = g*y0*(1-y0), dL/dx1 = -g*y0*(1-y0)."""
    g = (gout * prev_decision)[..., 0]
    t = g * ysoft0 * (1.0 - ysoft0)
    return torch.stack([t, -t], dim=-1)


# ----------------------------------------------------------------------------------------------
# policy-masked attention
# ----------------------------------------------------------------------------------------------

def softmax_with_policy(attn: torch.Tensor, policy: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """attn (B,H,T,T) scaled scores, policy (B,T,1) -> probabilities (B,H,T,T).

    vit_models/dynamic_vit.py:195-214 == default_dynamic_vit.py:185-199.  Not a masked softmax: the row
    max runs over ALL keys; the keep mask m_ij = p_j + (1-p_j)*[i==j] multiplies the exponentials; eps/T is
    added to every entry and eps to the denominator.  Closed form used here:
        P_ij = (exp(s_ij - max_j s_ij) * m_ij + eps/T) / (sum_j exp(.)*m_ij + eps)
    """
    B, H, T, _ = attn.shape
    p = policy.reshape(B, 1, 1, T).float()
    diag = torch.eye(T, device=attn.device).view(1, 1, T, T)
    m = p + (1.0 - p) * diag
    e = torch.exp(attn.float() - attn.float().amax(dim=-1, keepdim=True)) * m
    out = (e + eps / T) / (e.sum(dim=-1, keepdim=True) + eps)
    return out.to(attn.dtype)


def attention(x, wqkv, bqkv, wproj, bproj, num_heads, policy=None, return_cls_attn=False, scale=None):
    """Multi-head self-attention with optional keep policy and CLS-row side output
    (vit_models/dynamic_vit.py:216-236, default_dynamic_vit.py:201-216).  x (B,T,D)."""
    B, T, D = x.shape
    hd = D // num_heads
    scale = hd ** -0.5 if scale is None else scale
    qkv = F.linear(x, wqkv, bqkv).view(B, T, 3, num_heads, hd)
    q = qkv[:, :, 0].transpose(1, 2)
    k = qkv[:, :, 1].transpose(1, 2)
    v = qkv[:, :, 2].transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1) if policy is None else softmax_with_policy(s, policy)
    o = torch.matmul(p, v).transpose(1, 2).reshape(B, T, D)
    o = F.linear(o, wproj, bproj)
    if return_cls_attn:
        return o, p[:, :, 0, :]
    return o


def attention_core(qkv, num_heads, policy=None, scale=None, eps=1e-6):
    """The part of `attention` the CUDA kernel replaces: packed qkv (B,T,3,H,hd) -> (out (B,T,H*hd),
    cls_row (B,H,T)).  Same math as vit_models/dynamic_vit.py:220-234."""
    B, T, _, H, hd = qkv.shape
    scale = hd ** -0.5 if scale is None else scale
    q = qkv[:, :, 0].transpose(1, 2).float()
    k = qkv[:, :, 1].transpose(1, 2).float()
    v = qkv[:, :, 2].transpose(1, 2).float()
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1) if policy is None else softmax_with_policy(s, policy.reshape(B, T, 1), eps)
    o = torch.matmul(p, v).transpose(1, 2).reshape(B, T, H * hd)
    return o, p[:, :, 0, :].contiguous()


# ----------------------------------------------------------------------------------------------
# score predictors
# ----------------------------------------------------------------------------------------------

def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def _ln(sd, key, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def predictor_a(sd, prefix, x, policy):
    """Variant A predictor (vit_models/default_dynamic_vit.py:304-330).  x (B,N,D), policy (B,N,1)
    -> log-probs (B,N,2) [keep, drop].  LN -> Linear -> GELU; channel split into a local half and a
    policy-weighted global mean of the other half; 3 Linear layers; log-softmax."""
    h = F.gelu(_lin(sd, prefix + ".in_conv.1", _ln(sd, prefix + ".in_conv.0", x)))
    B, N, C = h.shape
    half = C // 2
    pooled = (h[:, :, half:] * policy).sum(dim=1, keepdim=True) / policy.sum(dim=1, keepdim=True)
    h = torch.cat([h[:, :, :half], pooled.expand(B, N, half)], dim=-1)
    h = F.gelu(_lin(sd, prefix + ".out_conv.0", h))
    h = F.gelu(_lin(sd, prefix + ".out_conv.2", h))
    return F.log_softmax(_lin(sd, prefix + ".out_conv.4", h), dim=-1)


def predictor_a_hidden(sd, prefix, x, policy):
    """Everything of predictor_a before its tail: returns the (B,N,D/4) post-GELU hidden the
    fused CUDA tail consumes."""
    h = F.gelu(_lin(sd, prefix + ".in_conv.1", _ln(sd, prefix + ".in_conv.0", x)))
    B, N, C = h.shape
    half = C // 2
    pooled = (h[:, :, half:] * policy).sum(dim=1, keepdim=True) / policy.sum(dim=1, keepdim=True)
    h = torch.cat([h[:, :, :half], pooled.expand(B, N, half)], dim=-1)
    h = F.gelu(_lin(sd, prefix + ".out_conv.0", h))
    return F.gelu(_lin(sd, prefix + ".out_conv.2", h))


def score_tail_a(hidden, w, b):
    """Tail of Variant A's predictor: Linear(D/4,2) + LogSoftmax (default_dynamic_vit.py:319-320)."""
    return F.log_softmax(F.linear(hidden.float(), w, b), dim=-1)


def predictor_b_layout(small: bool):
    """Layer indices inside `out_conv` for Variant B's LayerNorm predictors
    (vit_models/dynamic_vit.py:409-426 small, :491-531 large): list of (norm_idx, linear_idx)."""
    n_stages = 3 if small else 5
    return [(3 * s, 3 * s + 1) for s in range(n_stages)]


def _norm_b(sd, key, x, use_bn, training):
    if not use_bn:
        return _ln(sd, key, x)
    # BatchNormLayer (vit_models/dynamic_vit.py:350-367): BatchNorm1d over channels of (B,N,C)
    w, b = sd[key + ".bn.weight"], sd[key + ".bn.bias"]
    if training:
        mean = x.mean(dim=(0, 1))
        var = x.var(dim=(0, 1), unbiased=False)
    else:
        mean, var = sd[key + ".bn.running_mean"], sd[key + ".bn.running_var"]
    return (x - mean) / torch.sqrt(var + 1e-5) * w + b


def predictor_b_hidden(sd, prefix, x, small=False, use_bn=False, training=False):
    """Variant B predictor up to (not including) its last norm+Linear(D/4,1)
    (vit_models/dynamic_vit.py:536-547).  Returns the (B,N,D/4) activation entering the tail."""
    act = F.gelu if (small and not use_bn) else F.relu
    h = act(_lin(sd, prefix + ".in_conv.1", _norm_b(sd, prefix + ".in_conv.0", x, use_bn, training)))
    B, N, C = h.shape
    half = C // 2
    pooled = h[:, :, half:].mean(dim=1, keepdim=True)
    h = torch.cat([h[:, :, :half], pooled.expand(B, N, half)], dim=-1)
    stages = predictor_b_layout(small)
    for ni, li in stages[:-1]:
        h = act(_lin(sd, f"{prefix}.out_conv.{li}", _norm_b(sd, f"{prefix}.out_conv.{ni}", h, use_bn, training)))
    return h


def score_tail_b(hidden, ln_w, ln_b, w, b, loss_type="kl_div", ln_eps=1e-5):
    """Tail of Variant B's predictor: LayerNorm(D/4) + Linear(D/4,1) + flatten, then softmax over the
    N tokens (kl_div/mse) or sigmoid (bce)  (vit_models/dynamic_vit.py:424-426, 547-554).
    Returns (scores (B,N), keep_probs (B,N))."""
    h = hidden.float()
    if ln_w is not None:
        h = F.layer_norm(h, (h.shape[-1],), ln_w, ln_b, ln_eps)
    scores = F.linear(h, w, b).flatten(-2, -1)
    probs = torch.softmax(scores, dim=-1) if loss_type in ("kl_div", "mse") else torch.sigmoid(scores)
    return scores, probs


def predictor_b(sd, prefix, x, small=False, use_bn=False, loss_type="kl_div", training=False):
    """Full Variant B predictor -> (scores (B,N), keep_probs (B,N))."""
    h = predictor_b_hidden(sd, prefix, x, small, use_bn, training)
    ni, li = predictor_b_layout(small)[-1]
    if use_bn:
        h = _norm_b(sd, f"{prefix}.out_conv.{ni}", h, True, training)
        return score_tail_b(h, None, None, sd[f"{prefix}.out_conv.{li}.weight"],
                            sd[f"{prefix}.out_conv.{li}.bias"], loss_type)
    return score_tail_b(h, sd[f"{prefix}.out_conv.{ni}.weight"], sd[f"{prefix}.out_conv.{ni}.bias"],
                        sd[f"{prefix}.out_conv.{li}.weight"], sd[f"{prefix}.out_conv.{li}.bias"], loss_type)


# ----------------------------------------------------------------------------------------------
# PerturbedTopK
# ----------------------------------------------------------------------------------------------

def perturbed_topk_fwd(x, noise, k, sigma):
    """x (b,d), noise (b,nS,d) injected standard normal -> (indicators (b,k,d), egrad (b,k,d)).

    vit_models/peturbed_topk.py:27-51: per sample take the top-k of x + sigma*noise, sort the k indices
    ascending, one-hot them to (k,d) and average over samples.  Restated without the (b,nS,k,d) tensor:
    the j-th smallest kept index of each sample increments counts[j, idx]; the same pass accumulates the
    backward's expected-gradient tensor  egrad[j, idx] += noise[idx]  (peturbed_topk.py:77-78), scaled by
    1/(nS*sigma).  Counts are exact integers, so indicators are bit-identical to the reference;
    egrad is accumulated in fp64 and rounded once.
    """
    b, nS, d = noise.shape
    pert = x[:, None, :] + noise * sigma
    idx = torch.sort(pert, dim=-1, descending=True, stable=True).indices[:, :, :k]
    idx = torch.sort(idx, dim=-1).values                                  # (b,nS,k) ascending token index
    flat = (torch.arange(k).view(1, 1, k) * d + idx).reshape(b, nS * k)   # slot j*d + idx
    counts = torch.zeros(b, k * d, dtype=torch.float64)
    counts.scatter_add_(1, flat, torch.ones(b, nS * k, dtype=torch.float64))
    picked = torch.gather(noise, 2, idx).reshape(b, nS * k).double()
    eg = torch.zeros(b, k * d, dtype=torch.float64)
    eg.scatter_add_(1, flat, picked)
    indicators = (counts.float() / nS).view(b, k, d)
    egrad = (eg / nS / sigma).float().view(b, k, d)
    return indicators, egrad


def perturbed_topk_bwd(gout, egrad):
    """grad_x[b,d] = sum_k gout[b,k,d] * egrad[b,k,d]  (vit_models/peturbed_topk.py:79)."""
    return (gout.double() * egrad.double()).sum(dim=1).float()


def gelu_exact(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


# ----------------------------------------------------------------------------------------------
# residual add + LayerNorm
# ----------------------------------------------------------------------------------------------

def add_layernorm(x, y, weight, bias, eps, norm_row0=0):
    """What the reference computes with two ops in Block.forward (vit_models/dynamic_vit.py:263-283):
    s = x + y in the tensor dtype, then LayerNorm(s[:, norm_row0:]) (torch computes the statistics in fp32
    and rounds the result to the tensor dtype).  Returns (s, h)."""
    s = x if y is None else x + y
    h = F.layer_norm(s[:, norm_row0:].float(), (s.shape[-1],), weight.float(), bias.float(), eps).to(s.dtype)
    return s, h


# ----------------------------------------------------------------------------------------------
# predictor body pieces and token assembly (checkers for the fused inference kernels)
# ----------------------------------------------------------------------------------------------

def _act(x, act):
    return F.gelu(x) if act == "gelu" else (F.relu(x) if act == "relu" else x)


def pool_act(z, policy, act="gelu"):
    """z (B,N,C) = in_conv Linear output.  local = act(z)[..., :C/2]; pooled = policy-weighted mean over tokens of
    act(z)[..., C/2:] (vit_models/default_dynamic_vit.py:325-328; policy None = plain mean, dynamic_vit.py:542)."""
    h = _act(z, act)
    half = z.shape[-1] // 2
    local, glob = h[..., :half], h[..., half:]
    if policy is None:
        pooled = glob.float().mean(dim=1)
    else:
        p = policy.reshape(z.shape[0], z.shape[1], 1).float()
        pooled = (glob.float() * p).sum(dim=1) / p.sum(dim=1)
    return local.contiguous(), pooled.to(z.dtype)


def bias_act(u, bias, act="gelu"):
    """act(u + bias) with bias (B,C) broadcast over tokens or (C,) shared."""
    b = bias.unsqueeze(1) if bias.dim() == 2 else bias
    return _act(u + b.to(u.dtype), act)


def assemble_tokens(patches, cls_token, pos_embed):
    """cat(cls, patches) + pos_embed (vit_models/dynamic_vit.py:820-823)."""
    B = patches.shape[0]
    x = torch.cat([cls_token.reshape(1, 1, -1).expand(B, -1, -1).to(patches.dtype), patches], dim=1)
    return x + pos_embed.reshape(1, x.shape[1], -1).to(patches.dtype)


def linear_residual_ln(a, w, b, x, gamma, beta, eps):
    """attn.proj / mlp.fc2 + residual add + the next LayerNorm of Block.forward (vit_models/dynamic_vit.py:263-283) as the
    reference's bf16 modules compute them: Linear output rounded to bf16, residual sum rounded to bf16, LayerNorm statistics in
    fp32 and its output rounded to bf16.  Inputs are bf16 tensors; returns (x', h) in bf16."""
    y = F.linear(a.float(), w.float(), None if b is None else b.float()).bfloat16()
    s = (x.float() + y.float()).bfloat16()
    h = F.layer_norm(s.float(), (s.shape[-1],), gamma.float(), beta.float(), eps).bfloat16()
    return s, h
