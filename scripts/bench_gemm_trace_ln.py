import os, sys; sys.path.insert(0, "/root/repo")
import torch, d2s
ops = d2s.pkg.ops
D, T = 384, 197
M = 1024 * T
x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
fc2 = torch.nn.Linear(4 * D, D).cuda().bfloat16()
proj = torch.nn.Linear(D, D).cuda().bfloat16()
u = torch.randn(M, 4 * D, device="cuda", dtype=torch.bfloat16)
ln = torch.nn.LayerNorm(D, eps=1e-6).cuda().bfloat16()
buf = torch.zeros(148 * 8 * 8, dtype=torch.int64, device="cuda")
names = ["wait tmem_full", "wait xfull", "pass1", "release+bar+store x'+wait", "stats+bar", "pass2", "fence+bar+store h+wait"]
for nm, a, lin in (("fc2", u, fc2), ("proj", x, proj)):
    for dbg in (0, 8):
        os.environ["D2S_GEMM_DEBUG"] = str(dbg)
        os.environ.pop("D2S_GEMM_TRACE", None)
        for _ in range(3): ops.linear_residual_ln(a, lin.weight, lin.bias, x, ln.weight, ln.bias, 1e-6)
        os.environ["D2S_GEMM_TRACE"] = str(buf.data_ptr())
        ops.linear_residual_ln(a, lin.weight, lin.bias, x, ln.weight, ln.bias, 1e-6)
        torch.cuda.synchronize()
        t = buf.view(148, 8, 8).double()
        tiles = t[:, :, 7].mean().item()
        print(f"{nm} dbg={dbg}: tiles/CTA {tiles:.1f}; cycles per tile: " + ", ".join(f"{n} {t[:, :, i].mean().item() / tiles:.0f}" for i, n in enumerate(names)))
