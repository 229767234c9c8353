"""Training-step throughput (BASELINE.json configs[2]): DynamicViT DeiT-S/16 (Variant A) student with the predictor,
ratio + distillation losses against a frozen DeiT-S teacher, bf16 autocast, AdamW, batch 256 per GPU; DDP (NCCL
gradient all-reduce overlapped with backward) when launched under torchrun.  Rank 0 prints one JSON line.

    python scripts/bench_train.py [--batch 256] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/bench_train.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--variant", default="a", choices=["a", "b"],
                help="a: DynamicViT (3 stages, Gumbel decisions, DistillDiffPruningLoss); b: Dense2Sparse (1 stage @3, top-k, MaskLoss + BackboneLoss, train.py:40-53)")
ap.add_argument("--graph", action="store_true", help="capture the whole step (forward, loss, backward, AdamW) in a CUDA graph (Variant A, one GPU)")
ap.add_argument("--flat-adamw", action="store_true", help="runner.FlatAdamW (one d2s kernel, gradients written into its flat buffer) instead of torch.optim.AdamW; single GPU")
ap.add_argument("--freeze-backbone", action="store_true", help="train the predictors only (mask_predictor.py:219-225)")
args = ap.parse_args()
pkg = d2s.pkg
rank, local, world = pkg.runner.dist_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
if args.variant == "a":
    student = pkg.variant_a.DefaultVisionTransformerDiffPruning(pruning_loc=bench.LOCS, token_ratio=bench.RATIOS, distill=True,
                                                                **bench.DEIT_S).to(dev).train()
    teacher = pkg.variant_a.DefaultVisionTransformerTeacher(**bench.DEIT_S).to(dev).eval()
else:
    student = pkg.variant_b.VisionTransformerDiffPruning(pruning_loc=[3], token_ratio=[0.7], distill=True, topk_selection=True,
                                                         predictor_loss_type="kl_div", **bench.DEIT_S).to(dev).train()
    teacher = pkg.variant_b.VisionTransformerTeacher(**bench.DEIT_S).to(dev).eval()
for p in teacher.parameters():
    p.requires_grad_(False)
if args.freeze_backbone:
    for n, p in student.named_parameters():
        p.requires_grad_("score_predictor" in n)
model = pkg.runner.wrap_ddp(student, dev) if world > 1 else student
if args.variant == "a":
    crit = pkg.losses.DistillDiffPruningLoss(teacher, keep_ratio=bench.RATIOS)
else:
    import types
    mask_loss_fn = pkg.losses.MaskLoss(types.SimpleNamespace(keep_ratios=[0.7], mask_loss_type="kl_div", batch_size=args.batch,
                                                             device=dev), "train")
    backbone_loss_fn = pkg.losses.BackboneLoss(types.SimpleNamespace(mixup=0.0, patch_score_threshold=None))
metrics = {}
use_graph = args.graph and args.variant == "a" and world == 1
if args.flat_adamw and world == 1:
    opt = pkg.runner.FlatAdamW([p for p in student.parameters() if p.requires_grad], lr=5e-4, weight_decay=0.05)
else:
    opt = torch.optim.AdamW([p for p in student.parameters() if p.requires_grad], lr=5e-4, weight_decay=0.05, capturable=use_graph)
g = torch.Generator(device=dev).manual_seed(42 + rank)
x = torch.randn(args.batch, 3, 224, 224, device=dev, generator=g)
y = torch.randint(0, 1000, (args.batch,), device=dev, generator=g)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        if args.variant == "a":
            out = model(x)
            loss, parts = crit(x, out, y)
        else:   # train.py:40-53
            with torch.no_grad():
                logits_t, token_t, cls_attn_weights = teacher(x)
            logits_s, token_s, pred_logits, kept_token_idx = model(x)
            loss = (mask_loss_fn(pred_logits, cls_attn_weights, kept_token_idx, metrics)
                    + backbone_loss_fn(logits_s, token_s, logits_t, token_t, kept_token_idx, y, metrics))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


if use_graph:
    def fwd_loss(xx, yy):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return crit(xx, model(xx), yy)[0]
    n_cap = pkg._lib.launch_count()
    graphed = pkg.runner.TrainStepRunner(fwd_loss, opt, x, y, warmup=max(1, args.warmup), grads=getattr(opt, "grads", None),
                                         weight_cache=getattr(opt, "weight_cache", None))
    per_step_captured = (pkg._lib.launch_count() - n_cap) // (max(1, args.warmup) + 1)   # eager warm-ups + the capture pass
    step = graphed

n0 = pkg._lib.launch_count()
for _ in range(args.warmup):
    loss = step()
per_step = per_step_captured if use_graph else (pkg._lib.launch_count() - n0) // max(1, args.warmup)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if os.environ.get("D2S_PROFILE_ONE_STEP"):      # ncu --profile-from-start off: one extra step between start/stop
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
if rank == 0:
    t = float(ms.item()) / args.steps
    print(json.dumps({"metric": "training images/sec " + ("DynamicViT (Variant A, 3 stages)" if args.variant == "a" else "Dense2Sparse (Variant B, 1 stage@3)") + " DeiT-S kr=0.7 @224", "value": world * args.batch / (t / 1e3),
                      "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t,
                      "dtype": "bf16 autocast", "data": "synthetic", "scaling": "weak", "final_loss": float(loss.detach()),
                      "d2s_launches_per_step": int(per_step),
                      "config": {"workload": "student fwd+bwd + frozen teacher fwd + AdamW, " + ("ratio/distill losses" if args.variant == "a" else "MaskLoss(kl_div) + BackboneLoss"),
                                 "batch_per_gpu": args.batch, "freeze_backbone": args.freeze_backbone,
                                 "parallelism": f"DDP x{world} (NCCL gradient all-reduce)" if world > 1 else "single GPU", "cuda_graph": bool(use_graph),
                                 "optimizer": type(opt).__name__}}))
if world > 1:
    dist.destroy_process_group()
