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
student = pkg.variant_a.DefaultVisionTransformerDiffPruning(pruning_loc=bench.LOCS, token_ratio=bench.RATIOS, distill=True,
                                                            **bench.DEIT_S).to(dev).train()
teacher = pkg.variant_a.DefaultVisionTransformerTeacher(**bench.DEIT_S).to(dev).eval()
for p in teacher.parameters():
    p.requires_grad_(False)
if args.freeze_backbone:
    for n, p in student.named_parameters():
        p.requires_grad_("score_predictor" in n)
model = pkg.runner.wrap_ddp(student, dev) if world > 1 else student
crit = pkg.losses.DistillDiffPruningLoss(teacher, keep_ratio=bench.RATIOS)
opt = torch.optim.AdamW([p for p in student.parameters() if p.requires_grad], lr=5e-4, weight_decay=0.05)
g = torch.Generator(device=dev).manual_seed(42 + rank)
x = torch.randn(args.batch, 3, 224, 224, device=dev, generator=g)
y = torch.randint(0, 1000, (args.batch,), device=dev, generator=g)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x)
        loss, parts = crit(x, out, y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


n0 = pkg._lib.launch_count()
for _ in range(args.warmup):
    loss = step()
per_step = (pkg._lib.launch_count() - n0) // max(1, args.warmup)
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
    print(json.dumps({"metric": "training images/sec DynamicViT DeiT-S kr=0.7 @224", "value": world * args.batch / (t / 1e3),
                      "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t,
                      "dtype": "bf16 autocast", "data": "synthetic", "scaling": "weak", "final_loss": float(loss.detach()),
                      "d2s_launches_per_step": int(per_step),
                      "config": {"workload": "student fwd+bwd + frozen teacher fwd + AdamW, ratio/distill losses",
                                 "batch_per_gpu": args.batch, "freeze_backbone": args.freeze_backbone,
                                 "parallelism": f"DDP x{world} (NCCL gradient all-reduce)" if world > 1 else "single GPU"}}))
if world > 1:
    dist.destroy_process_group()
