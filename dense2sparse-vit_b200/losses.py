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


class DistillDiffPruningLoss(torch.nn.Module):
    def __init__(self, teacher_model, base_criterion=None, ratio_weight=2.0, distill_weight=0.5, clf_weight=1.0,
                 keep_ratio=(0.7, 0.49, 0.343), print_mode=False):
        super().__init__()
        self.teacher_model = teacher_model
        self.base_criterion = base_criterion or torch.nn.CrossEntropyLoss()
        self.ratio_weight, self.distill_weight, self.clf_weight = ratio_weight, distill_weight, clf_weight
        self.keep_ratio = list(keep_ratio)
        self.print_mode = print_mode

    def forward(self, inputs, outputs, labels):
        pred, token_pred, mask, out_pred_score = outputs
        # keep-ratio loss: the mean keep decision of every stage should hit its target ratio
        ratio_loss = 0.0
        for i, score in enumerate(out_pred_score):
            ratio_loss = ratio_loss + ((score.float().mean(dim=1) - self.keep_ratio[i]) ** 2).mean()
        cls_loss = self.base_criterion(pred.float(), labels)
        with torch.no_grad():
            cls_t, token_t = self.teacher_model(inputs)[:2]
        cls_kl = F.kl_div(F.log_softmax(pred.float(), dim=-1), F.log_softmax(cls_t.float(), dim=-1),
                          reduction="batchmean", log_target=True)
        B, N, C = token_pred.shape
        keep = mask.reshape(B * N) > 0.5
        tp, tt = token_pred.reshape(B * N, C)[keep].float(), token_t.reshape(B * N, C)[keep].float()
        if tp.shape[0] == 0:
            token_kl = token_pred.new_zeros(())
        else:
            token_kl = F.kl_div(F.log_softmax(tp, dim=-1), F.log_softmax(tt, dim=-1), reduction="batchmean", log_target=True)
        loss = (self.clf_weight * cls_loss + self.ratio_weight * ratio_loss / max(1, len(out_pred_score))
                + self.distill_weight * (cls_kl + token_kl))
        return loss, dict(cls=cls_loss.detach(), ratio=torch.as_tensor(ratio_loss).detach(), cls_kl=cls_kl.detach(),
                          token_kl=token_kl.detach())
