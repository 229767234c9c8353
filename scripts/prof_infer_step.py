"""Kernel list of one eager inference step (torch profiler, device time per launch, in launch order).
    python scripts/prof_infer_step.py [small|base|b] [batch]        (b: Dense2Sparse / Variant B, DeiT-S, one stage at block 3)
Diagnostic only: profiler timings are never a bench value."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import d2s, bench
pkg = d2s.pkg
dev = torch.device("cuda", 0)
arch = sys.argv[1] if len(sys.argv) > 1 else "small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (512 if arch == "base" else 1024)
if arch == "b":
    model = pkg.variant_b.VisionTransformerDiffPruning(pruning_loc=[3], token_ratio=[0.7], distill=True, topk_selection=True,
                                                       predictor_loss_type="kl_div", **bench.DEIT_S)
else:
    model = pkg.variant_a.DefaultVisionTransformerDiffPruning(pruning_loc=bench.LOCS, token_ratio=bench.RATIOS, distill=True,
                                                              **(bench.DEIT_S if arch == "small" else bench.DEIT_B))
bench.seeded_weights(model)
model = model.to(dev).eval().to(torch.bfloat16)
x = torch.randn(B, 3, 224, 224, device=dev).bfloat16()
with torch.no_grad():
    for _ in range(3): model(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(x)
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in evs)
print("total us", tot, "launches", len(evs))
for e in evs:
    print(f"{e.device_time_total:8.1f}  {e.name[:110]}")
