import os, sys; sys.path.insert(0, "/root/repo")
import torch, d2s
ops = d2s.pkg.ops
D, T = 384, 197
M = 1024 * T
x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
fc1 = torch.nn.Linear(D, 4 * D).cuda().bfloat16()
buf = torch.zeros(148 * 8 * 8, dtype=torch.int64, device="cuda")
for act in (0, 1):
    for dbg in (0, 1):
        os.environ["D2S_GEMM_DEBUG"] = str(dbg)
        os.environ.pop("D2S_GEMM_TRACE", None)
        for _ in range(3): ops.linear_act(x, fc1.weight, fc1.bias, act, pair=True)
        os.environ["D2S_GEMM_TRACE"] = str(buf.data_ptr())
        ops.linear_act(x, fc1.weight, fc1.bias, act, pair=True)
        torch.cuda.synchronize()
        t = buf.view(148, 8, 8).double()
        tiles = t[:, :, 5].mean().item()
        names = ["wait tmem_full", "wait store-read + bar", "tcgen05.ld + wait", "math + STS", "release/fence/bar/store"]
        print(f"act={act} dbg={dbg}: tiles/CTA {tiles:.1f}; cycles per tile (mean over warps): " +
              ", ".join(f"{n} {t[:, :, i].mean().item() / tiles:.0f}" for i, n in enumerate(names)))
        print("   CTA0 warps:", (t[0, :, :5] / tiles).round().tolist())
