import os, sys; sys.path.insert(0, "/root/repo")
import torch, d2s
ops = d2s.pkg.ops
D, T = 384, 197
M = 1024 * T
h = torch.randn(M, D, device="cuda", dtype=torch.bfloat16); x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
fc1 = torch.nn.Linear(D, 4 * D).cuda().bfloat16(); fc2 = torch.nn.Linear(4 * D, D).cuda().bfloat16(); ln = torch.nn.LayerNorm(D, eps=1e-6).cuda().bfloat16()
buf = torch.zeros(74 * 8, dtype=torch.int64, device="cuda")
f = lambda: ops.mlp_residual_ln(h, fc1.weight, fc1.bias, fc2.weight, fc2.bias, x, ln.weight, ln.bias, 1e-6)
for _ in range(3): f()
os.environ["D2S_GEMM_TRACE"] = str(buf.data_ptr())
f(); torch.cuda.synchronize()
t = buf.view(74, 8).double()
ch = t[:, 7].mean().item()
names = ["issue+other", "wait s_empty", "wait a1_full", "wait w1_full", "wait p_full", "wait w2_full", "wait acc_empty"]
print(f"chunks per pair {ch:.0f}; MMA-thread cycles per chunk: " + ", ".join(f"{n} {t[:, i].mean().item() / ch:.0f}" for i, n in enumerate(names)))
