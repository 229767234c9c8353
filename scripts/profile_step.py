"""One eager inference step of the bench workload between cudaProfilerStart/Stop, for ncu:

    python scripts/profile_step.py [--batch 1024] [--variant a|b]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python scripts/profile_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--variant", default="a")
ap.add_argument("--warm", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
pkg = d2s.pkg
if args.variant == "a":
    model = pkg.variant_a.DefaultVisionTransformerDiffPruning(pruning_loc=bench.LOCS, token_ratio=bench.RATIOS, distill=True,
                                                              **bench.DEIT_S)
else:
    model = pkg.variant_b.VisionTransformerDiffPruning(pruning_loc=[3], token_ratio=[0.7], distill=True, topk_selection=True,
                                                       predictor_loss_type="kl_div", **bench.DEIT_S)
model = model.to(dev, torch.bfloat16).eval()
x = torch.randn(args.batch, 3, 224, 224, device=dev).to(torch.bfloat16)
with torch.no_grad():
    for _ in range(args.warm):
        model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", float((out[0] if isinstance(out, tuple) else out).float().abs().mean()))
