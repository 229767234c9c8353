import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import ProfilerActivity, profile
import d2s, bench
pkg = d2s.pkg
dev = torch.device("cuda", 0)
model = pkg.variant_a.DefaultVisionTransformerDiffPruning(pruning_loc=bench.LOCS, token_ratio=bench.RATIOS, distill=True, **bench.DEIT_S)
bench.seeded_weights(model)
model = model.to(dev).eval().to(torch.bfloat16)
x = torch.randn(1024, 3, 224, 224, device=dev).bfloat16()
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
