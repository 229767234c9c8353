import sys, copy; sys.path.insert(0, "/root/repo")
import torch, d2s, bench
pkg = d2s.pkg
dev = torch.device("cuda:0")
torch.manual_seed(0)
student0 = pkg.variant_a.DefaultVisionTransformerDiffPruning(pruning_loc=bench.LOCS, token_ratio=bench.RATIOS, distill=True, **bench.DEIT_S).to(dev).train()
teacher = pkg.variant_a.DefaultVisionTransformerTeacher(**bench.DEIT_S).to(dev).eval()
for p in teacher.parameters(): p.requires_grad_(False)
g = torch.Generator(device=dev).manual_seed(42)
x = torch.randn(64, 3, 224, 224, device=dev, generator=g); y = torch.randint(0, 1000, (64,), device=dev, generator=g)
for mode in ("eager", "graph"):
    torch.manual_seed(1)
    m = copy.deepcopy(student0)
    crit = pkg.losses.DistillDiffPruningLoss(teacher, keep_ratio=bench.RATIOS)
    opt = torch.optim.AdamW(m.parameters(), lr=5e-4, weight_decay=0.05, capturable=True)
    def f(xx, yy):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return crit(xx, m(xx), yy)[0]
    run = pkg.runner.TrainStepRunner(f, opt, x, y, warmup=3, use_graph=mode == "graph")
    print(mode, [round(float(run().detach()), 3) for _ in range(24)])
