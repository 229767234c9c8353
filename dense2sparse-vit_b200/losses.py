"""Training losses that consume the hot path's outputs.

DistillDiffPruningLoss is what the reference's DDP script instantiates (`losses.DistillDiffPruningLoss(teacher_model=...,
clf_weight=1.0, base_criterion=CrossEntropyLoss())`, ddp_training.py:81) but its own losses.py no longer defines; it is
restated here from the upstream DynamicViT recipe the reference's Variant A model comes from (parity unpinned: there is
no reference implementation to run), with the weights of the reference's CLI defaults: ratio 2.0, distillation 0.5,
classification 1.0 (utils.py:234-244).  It consumes exactly the training tuple of
DefaultVisionTransformerDiffPruning.forward (default_dynamic_vit.py:481-485):
(logits, token features, final keep decision (B,196,1), [per-stage keep decisions (B,196)]).
"""
import torch
import torch.nn.functional as F


def _token_kl_rows(token_s, token_t):
    """KL(softmax(token_t) || softmax(token_s)) per token row, (B*N,) f32: F.kl_div(log_softmax(s), log_softmax(t),
    log_target=True) summed over the channels (losses.py:220-225).  One d2s pass over both tensors when it applies (CUDA,
    row-dense f32 | bf16 (B,N,C), C % 8 == 0, C <= 1024), torch's composition otherwise."""
    from . import ops
    B, N, C = token_s.shape
    if token_s.is_cuda and ops.token_kl_ok(token_s, token_t):
        return ops.token_kl_rows(token_s, token_t)
    lp = F.log_softmax(token_s.reshape(B * N, C).float(), dim=-1)
    lt = F.log_softmax(token_t.reshape(B * N, C).float(), dim=-1)
    return (lt.exp() * (lt - lp)).sum(dim=-1)


class DistillDiffPruningLoss(torch.nn.Module):
    def __init__(self, teacher_model, base_criterion=None, ratio_weight=2.0, distill_weight=0.5, clf_weight=1.0,
                 keep_ratio=(0.7, 0.49, 0.343), print_mode=False):
        super().__init__()
        self.teacher_model = teacher_model
        self.base_criterion = base_criterion or torch.nn.CrossEntropyLoss()
        self.ratio_weight, self.distill_weight, self.clf_weight = ratio_weight, distill_weight, clf_weight
        self.keep_ratio = list(keep_ratio)
        self.print_mode = print_mode
        self._pending = None

    def start_teacher(self, inputs, stream):
        """Launch the frozen teacher's forward on `stream` NOW (typically right before the student's forward on the current
        stream): the two forwards are independent, so their kernels interleave and each fills the other's launch and tail gaps;
        forward() on the same `inputs` then joins the stream and uses the outputs instead of running the teacher itself.
        Works under CUDA-graph capture (the side stream forks from and rejoins the capturing stream)."""
        cur = torch.cuda.current_stream(inputs.device)
        stream.wait_stream(cur)
        with torch.cuda.stream(stream), torch.no_grad():
            out = self.teacher_model(inputs)[:2]
        self._pending = (inputs, out, stream)

    def _teacher_outputs(self, inputs):
        pending, self._pending = self._pending, None
        if pending is not None and pending[0] is inputs:
            torch.cuda.current_stream(inputs.device).wait_stream(pending[2])
            return pending[1]
        with torch.no_grad():
            return self.teacher_model(inputs)[:2]

    def forward(self, inputs, outputs, labels):
        pred, token_pred, mask, out_pred_score = outputs
        # keep-ratio loss: the mean keep decision of every stage should hit its target ratio
        ratio_loss = 0.0
        for i, score in enumerate(out_pred_score):
            ratio_loss = ratio_loss + ((score.float().mean(dim=1) - self.keep_ratio[i]) ** 2).mean()
        cls_loss = self.base_criterion(pred.float(), labels)
        cls_t, token_t = self._teacher_outputs(inputs)
        cls_kl = F.kl_div(F.log_softmax(pred.float(), dim=-1), F.log_softmax(cls_t.float(), dim=-1),
                          reduction="batchmean", log_target=True)
        # token distillation over the KEPT tokens: KL(teacher || student) per token row, "batchmean" over the kept rows.  Written
        # as a masked mean over all rows (same value as indexing the kept rows first) so that every shape is static: no boolean
        # indexing, no host synchronisation -- the whole training step can be captured in a CUDA graph (runner.TrainStepRunner).
        B, N, C = token_pred.shape
        keep = (mask.reshape(B * N) > 0.5).float()
        kl_rows = _token_kl_rows(token_pred, token_t)
        token_kl = (kl_rows * keep).sum() / keep.sum().clamp_min(1.0)
        loss = (self.clf_weight * cls_loss + self.ratio_weight * ratio_loss / max(1, len(out_pred_score))
                + self.distill_weight * (cls_kl + token_kl))
        return loss, dict(cls=cls_loss.detach(), ratio=torch.as_tensor(ratio_loss).detach(), cls_kl=cls_kl.detach(),
                          token_kl=token_kl.detach())


# ---------------------------------------------------------------------------------------------------------------
# Loss-side consumers of the Dense2Sparse (Variant B) hot path: drop-ins for the reference's MaskLoss / BackboneLoss
# (losses.py:6-164, :167-241; called from train.py:46-48).  Same constructor arguments (`args` namespace, phase), same
# forward signatures, same `metrics` side effects and running averages.  The top-k masks reuse the d2s select kernel
# (the selection the model itself ran), the teacher-token gather the d2s gather kernel.
# ---------------------------------------------------------------------------------------------------------------
def _topk_mask(scores, keep_ratio):
    """MaskLoss.get_mask_from_pred_logits / get_mask_from_cls_attns (losses.py:121-164): 1.0 for the
    int(N * keep_ratio) highest scores, in token order."""
    from . import ops
    n_keep = int(scores.shape[-1] * keep_ratio)
    s = scores.detach().float()
    kept, _ = ops.select_topk(s, n_keep, ops.ORDER_SCORE_DESC, want_dropped=False)     # CUDA only: there is no CPU fallback
    return torch.zeros_like(s).scatter_(1, kept, 1.0)


class MaskLoss(torch.nn.Module):
    def __init__(self, args, phase):
        super().__init__()
        self.phase = phase
        self.keep_ratios = args.keep_ratios
        self.loss_type = args.mask_loss_type
        self.count = 1
        self.running_loss = 0
        self.runnings_accs = [0 for _ in self.keep_ratios]

    get_mask_from_pred_logits = staticmethod(_topk_mask)
    get_mask_from_cls_attns = staticmethod(_topk_mask)

    def forward(self, pred_logits, cls_attn_weights, kept_token_idx, metrics):
        if self.loss_type == "bce":
            # the reference's bce branch reads the undefined names `args` and `self.mask_criterions` (losses.py:57-58)
            raise NameError("MaskLoss: the 'bce' branch of the reference cannot run (losses.py:57-58 read undefined names)")
        mask_loss = 0
        mask_accs = [0 for _ in self.keep_ratios]
        w = torch.mean(cls_attn_weights.float(), dim=1)            # (B, H, N+1)   losses.py:82
        w, _ = torch.max(w, dim=1)                                 # (B, N+1)      :83
        target = w[:, 1:] / torch.sum(w[:, 1:], dim=-1, keepdim=True)
        for i in range(len(kept_token_idx)):
            logits = pred_logits[i].float()
            if self.loss_type == "mse":                            # losses.py:68-80
                if i > 0:
                    target = torch.gather(target, 1, kept_token_idx[i - 1])
                    target = target / torch.sum(target, dim=1, keepdim=True)
                mask_loss = mask_loss + 100 * F.mse_loss(logits, target, reduction="mean")
                continue
            if i > 0:                                              # losses.py:88-97
                ratio = self.keep_ratios[i] / self.keep_ratios[i - 1]
                gathered = torch.gather(target, 1, kept_token_idx[i - 1])
                gt = _topk_mask(gathered, ratio)
                target = gathered / torch.sum(gathered, dim=1, keepdim=True)
            else:
                ratio = self.keep_ratios[i]
                gt = _topk_mask(target, ratio)
            pred = _topk_mask(F.softmax(logits, dim=-1), ratio)
            mask_loss = mask_loss + F.kl_div(F.log_softmax(logits, dim=-1), torch.log(target), log_target=True,
                                             reduction="batchmean")
            mask_accs[i] = mask_accs[i] + torch.sum(pred == gt) / pred.numel()
        self.running_loss += mask_loss.detach().item()
        metrics[f"{self.phase}_mask_loss"] = self.running_loss / self.count
        for i, _ in enumerate(self.keep_ratios):
            self.runnings_accs[i] += mask_accs[i]
            metrics[f"{self.phase}_mask_acc_{i}"] = self.runnings_accs[i] / self.count
        self.count += 1
        return mask_loss


class BackboneLoss(torch.nn.Module):
    def __init__(self, args):
        super().__init__()
        if args.mixup > 0.:
            # soft-target cross entropy (timm.loss.SoftTargetCrossEntropy, timm==0.4.12): labels arrive mixed / smoothed
            self.base_criterion = lambda x, t: torch.sum(-t * F.log_softmax(x, dim=-1), dim=-1).mean()
        else:
            self.base_criterion = torch.nn.CrossEntropyLoss()
        self.patch_score_threshold = args.patch_score_threshold
        self.count = 1
        self.running_loss = 0
        self.running_cls_loss = 0
        self.running_token_kl_loss = 0
        self.running_token_dist_loss = 0
        self.runnings_acc = 0

    def forward(self, logits_s, token_s, logits_t, token_t, kept_token_idx, train_labels, metrics):
        from . import ops
        logits_s, logits_t = logits_s.float(), logits_t.float()
        cls_loss = self.base_criterion(logits_s, train_labels)
        cls_kl_loss = F.kl_div(F.log_softmax(logits_s, dim=-1), F.log_softmax(logits_t, dim=-1), reduction="batchmean",
                               log_target=True)
        if self.patch_score_threshold is not None:
            # the reference's threshold branch (losses.py:216-217) never binds `C`, which line :219 reads
            raise UnboundLocalError("BackboneLoss: `C` is read before assignment when patch_score_threshold is set "
                                    "(losses.py:217-219)")
        B, N, C = token_t.size()
        # teacher tokens of the kept positions: the LAST stage's indices into the full-length sequence (losses.py:212)
        token_t = ops.gather_tokens(token_t.detach(), kept_token_idx[-1], prepend_cls=False)
        if token_s.dim() == 3 and token_s.shape == token_t.shape:
            rows = token_s.shape[0] * token_s.shape[1]
            token_kl_loss = _token_kl_rows(token_s, token_t).sum() / rows           # "batchmean" over the kept rows
        else:
            token_s = token_s.reshape(-1, C).float()
            token_t = token_t.reshape(-1, C).float()
            token_kl_loss = F.kl_div(F.log_softmax(token_s, dim=-1), F.log_softmax(token_t, dim=-1), reduction="batchmean",
                                     log_target=True)
        backbone_loss = cls_loss + cls_kl_loss + token_kl_loss
        self.running_loss += backbone_loss.detach().item()
        self.running_cls_loss += cls_loss.detach().item()
        self.running_token_dist_loss += cls_kl_loss.detach().item()
        self.running_token_kl_loss += token_kl_loss.detach().item()
        metrics["train_backbone_loss"] = self.running_loss / self.count
        metrics["train_cls_loss"] = self.running_cls_loss / self.count
        metrics["train_token_kl_loss"] = self.running_token_dist_loss / self.count     # (sic: losses.py:236-237 swap the names)
        metrics["train_cls_kl_loss"] = self.running_token_kl_loss / self.count
        self.count += 1
        return backbone_loss
