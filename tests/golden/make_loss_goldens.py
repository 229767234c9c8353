"""Generate tests/golden/golden_losses.npz by executing the UNMODIFIED reference losses (/root/reference/losses.py, loaded by
path under the timm stub of refload.py) on seeded synthetic inputs.  Build container only:

    python tests/golden/make_loss_goldens.py
"""
import importlib.util
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures as fx  # noqa: E402
import refload  # noqa: E402

torch.set_num_threads(4)
refload._install_timm_stub()
spec = importlib.util.spec_from_file_location("_d2s_ref_losses", os.path.join(refload.REF_ROOT, "losses.py"))
ref_losses = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_losses)


def main():
    A, meta = {}, {"cases": {}}
    for name, seed, ratios in (("two_stage", 500, (0.7, 0.49)), ("one_stage", 501, (0.7,)), ("three_stage", 502, (0.7, 0.49, 0.343))):
        inp = fx.loss_inputs(seed, ratios=ratios)
        meta["cases"][name] = dict(seed=seed, ratios=list(ratios))
        for loss_type in ("kl_div", "mse"):
            args = types.SimpleNamespace(keep_ratios=list(ratios), mask_loss_type=loss_type, batch_size=4, device="cpu")
            mod = ref_losses.MaskLoss(args, "train")
            pl = [p.clone().requires_grad_(True) for p in inp["pred_logits"]]
            metrics = {}
            out = mod(pl, inp["cls_attn"], inp["kept"], metrics)
            out.backward()
            A[f"{name}::{loss_type}::loss"] = out.detach()
            for i, p in enumerate(pl):
                A[f"{name}::{loss_type}::grad{i}"] = p.grad
            A[f"{name}::{loss_type}::metric_loss"] = torch.tensor(metrics["train_mask_loss"])
            for i in range(len(ratios)):
                A[f"{name}::{loss_type}::acc{i}"] = torch.as_tensor(metrics[f"train_mask_acc_{i}"], dtype=torch.float32)
            # second call: running averages (count = 2)
            out2 = mod([p.detach() for p in pl], inp["cls_attn"], inp["kept"], metrics)
            A[f"{name}::{loss_type}::metric_loss_2"] = torch.tensor(metrics["train_mask_loss"])
        for mix, lab in ((0.0, inp["labels"]), (0.8, inp["soft"])):
            args = types.SimpleNamespace(mixup=mix, patch_score_threshold=None)
            mod = ref_losses.BackboneLoss(args)
            ls, ts = inp["logits_s"].clone().requires_grad_(True), inp["token_s"].clone().requires_grad_(True)
            metrics = {}
            out = mod(ls, ts, inp["logits_t"], inp["token_t"], inp["kept"], lab, metrics)
            out.backward()
            tag = f"{name}::backbone{'_mix' if mix else ''}"
            A[f"{tag}::loss"] = out.detach()
            A[f"{tag}::grad_logits"], A[f"{tag}::grad_tokens"] = ls.grad, ts.grad
            for k, v in metrics.items():
                A[f"{tag}::{k}"] = torch.tensor(v)
        A[f"{name}::in_checksum"] = torch.tensor([fx.checksum(inp["cls_attn"]), fx.checksum(inp["token_t"])])
    # the mask helpers on their own (losses.py:121-164)
    sc = torch.softmax(fx.randn(510, 5, 196), dim=-1)
    A["mask_pred"] = ref_losses.MaskLoss.get_mask_from_pred_logits(sc, 0.7)
    A["mask_cls"] = ref_losses.MaskLoss.get_mask_from_cls_attns(sc, 0.49 / 0.7)
    # reference defects that the restatement keeps (recorded as the exception type)
    try:
        ref_losses.MaskLoss(types.SimpleNamespace(keep_ratios=[0.7], mask_loss_type="bce", batch_size=4, device="cpu"), "train")(
            [torch.randn(4, 196)], torch.rand(4, 4, 3, 197), [torch.zeros(4, 137, dtype=torch.long)], {})
        meta["bce_error"] = None
    except Exception as e:  # noqa: BLE001
        meta["bce_error"] = type(e).__name__
    try:
        ref_losses.BackboneLoss(types.SimpleNamespace(mixup=0.0, patch_score_threshold=0.9))(
            torch.randn(4, 16), torch.randn(4, 10, 32), torch.randn(4, 16), torch.randn(4, 196, 32),
            [torch.zeros(40, dtype=torch.long)], torch.zeros(4, dtype=torch.long), {})
        meta["threshold_error"] = None
    except Exception as e:  # noqa: BLE001
        meta["threshold_error"] = type(e).__name__
    fx.save_npz("golden_losses.npz", A, meta)
    print("wrote golden_losses.npz:", len(A), "arrays;", meta)


if __name__ == "__main__":
    main()
