"""CPU restatement (torch fp32) of the reference's loss-side consumers of the hot path: MaskLoss and BackboneLoss
(/root/reference/losses.py).  TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and bench.py's CPU legs, never by the
product package.  Pinned against the unmodified reference through tests/golden/golden_losses.npz
(tests/golden/make_loss_goldens.py).  Functions are stateless: the reference's running averages / metrics dict are
host bookkeeping and are restated in the product module, not here.
"""
import torch
import torch.nn.functional as F


def topk_mask(scores, keep_ratio):
    """MaskLoss.get_mask_from_pred_logits / get_mask_from_cls_attns (losses.py:121-164, two textual copies):
    1.0 for the int(N * keep_ratio) highest scores (argsort descending, scattered back to token order), else 0.0."""
    n_keep = int(scores.shape[-1] * keep_ratio)
    order = torch.argsort(scores, dim=-1, descending=True)
    mask = torch.zeros_like(scores, dtype=torch.float32)
    mask.scatter_(1, order[:, :n_keep], 1.0)
    return mask


def cls_attn_target(cls_attn_weights):
    """losses.py:47-50 / :70-73 / :82-85: teacher CLS attention (B, L, H, N+1) -> mean over layers, max over heads,
    drop the CLS column, renormalise over the N patch tokens -> (B, N)."""
    w = torch.mean(cls_attn_weights, dim=1)
    w, _ = torch.max(w, dim=1)
    return w[:, 1:] / torch.sum(w[:, 1:], dim=-1, keepdim=True)


def mask_loss(pred_logits, cls_attn_weights, kept_token_idx, keep_ratios, loss_type="kl_div"):
    """MaskLoss.forward (losses.py:32-119), branches "kl_div" (the else-branch, :81-104) and "mse" (:68-80).
    Returns (loss, [mask accuracy per stage]).  The "bce" branch of the reference reads undefined names
    (`args`, `self.mask_criterions`, losses.py:57-58) and cannot run."""
    target = cls_attn_target(cls_attn_weights)
    loss = 0
    accs = [0 for _ in keep_ratios]
    if loss_type == "bce":
        raise NameError("MaskLoss 'bce' branch reads undefined names in the reference (losses.py:57-58)")
    for i in range(len(kept_token_idx)):
        if loss_type == "mse":
            if i > 0:
                target = torch.gather(target, 1, kept_token_idx[i - 1])
                target = target / torch.sum(target, dim=1, keepdim=True)
            loss = loss + 100 * F.mse_loss(pred_logits[i], target, reduction="mean")
            continue
        if i > 0:
            ratio = keep_ratios[i] / keep_ratios[i - 1]
            gathered = torch.gather(target, 1, kept_token_idx[i - 1])
            gt = topk_mask(gathered, ratio)                       # before the renormalisation (losses.py:92-94)
            target = gathered / torch.sum(gathered, dim=1, keepdim=True)
        else:
            ratio = keep_ratios[i]
            gt = topk_mask(target, ratio)
        pred = topk_mask(F.softmax(pred_logits[i], dim=-1), ratio)
        loss = loss + F.kl_div(F.log_softmax(pred_logits[i], dim=-1), torch.log(target), log_target=True, reduction="batchmean")
        accs[i] = accs[i] + torch.sum(pred == gt) / pred.numel()
    return loss, accs


def backbone_loss(logits_s, token_s, logits_t, token_t, kept_token_idx, labels, patch_score_threshold=None, soft_labels=False):
    """BackboneLoss.forward (losses.py:182-241): CE (or soft-target CE under mixup) + KL(student || teacher) on the class
    logits + KL on the kept tokens' features, the teacher's tokens gathered with the LAST stage's kept indices
    (losses.py:212; stage-relative indices into the full-length teacher sequence - a reference defect that is kept)."""
    if soft_labels:
        cls_loss = torch.sum(-labels * F.log_softmax(logits_s, dim=-1), dim=-1).mean()
    else:
        cls_loss = F.cross_entropy(logits_s, labels)
    cls_kl = F.kl_div(F.log_softmax(logits_s, dim=-1), F.log_softmax(logits_t, dim=-1), reduction="batchmean", log_target=True)
    if patch_score_threshold is None:
        B, N, C = token_t.size()
        token_t = torch.gather(token_t, 1, kept_token_idx[-1].unsqueeze(-1).expand(-1, -1, C))
    else:
        # the reference's threshold branch (losses.py:216-217) never binds C, which the next line reads (:219)
        raise UnboundLocalError("BackboneLoss reads `C` before assignment when patch_score_threshold is set (losses.py:217-219)")
    token_s = token_s.reshape(-1, C)
    token_t = token_t.reshape(-1, C)
    token_kl = F.kl_div(F.log_softmax(token_s, dim=-1), F.log_softmax(token_t, dim=-1), reduction="batchmean", log_target=True)
    return cls_loss + cls_kl + token_kl, dict(cls=cls_loss, cls_kl=cls_kl, token_kl=token_kl)
