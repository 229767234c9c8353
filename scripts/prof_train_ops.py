"""Which host-side ops launch the non-d2s kernels of the training step?  Runs bench.py's training step eagerly (no CUDA graph)
under torch.profiler and prints device time by (op, input shapes) and by kernel.  Diagnostic only: profiler timings are never a
bench value.

    python scripts/prof_train_ops.py [--batch 256] > gpurun_out/train_ops.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--rows", type=int, default=70)
args = ap.parse_args()
pkg = d2s.pkg
dev = torch.device("cuda", 0)
torch.manual_seed(0)
student = pkg.variant_a.DefaultVisionTransformerDiffPruning(pruning_loc=bench.LOCS, token_ratio=bench.RATIOS, distill=True, **bench.DEIT_S)
bench.seeded_weights(student)
student = student.to(dev).train()
teacher = pkg.variant_a.DefaultVisionTransformerTeacher(**bench.DEIT_S).to(dev).eval()
for p in teacher.parameters():
    p.requires_grad_(False)
crit = pkg.losses.DistillDiffPruningLoss(teacher, keep_ratio=bench.RATIOS)
opt = pkg.runner.FlatAdamW(student.parameters(), lr=5e-4, weight_decay=0.05)
grads, wcache = opt.grads, opt.weight_cache
x = torch.randn(args.batch, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (args.batch,), device=dev)


def fwd_loss(xx, yy):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return crit(xx, student(xx), yy)[0]


run = pkg.runner.TrainStepRunner(fwd_loss, opt, x, y, warmup=2, use_graph=False, grads=grads, weight_cache=wcache)
for _ in range(2):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    run()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=args.rows, max_name_column_width=60,
                                                         max_shapes_column_width=90))
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=args.rows, max_name_column_width=110))
# aten ops with device time, grouped by the innermost repo frame that issued them
import collections
by_site = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if not ev.name.startswith("aten::") or ev.self_device_time_total <= 0:
        continue
    site = "?"
    for fr in (ev.stack or []):
        if "/root/repo/" in fr or "dense2sparse" in fr:
            site = fr.split("/")[-1]
            break
    k = (ev.name, site, str(ev.input_shapes)[:70])
    by_site[k][0] += ev.self_device_time_total
    by_site[k][1] += 1
print("\n==== aten ops by call site (self device us, calls) ====")
for k, (t, n) in sorted(by_site.items(), key=lambda kv: -kv[1][0])[:90]:
    print(f"{t:9.1f} {n:4d}  {k[0]:28s} {k[1]:60s} {k[2]}")
