"""Issuer-side wait breakdown of the fused MLP kernel (clock64 totals per barrier; needs a trace build:

    D2S_NVCC_EXTRA=-DD2S_GEMM_TRACE_BUILD python -c "import __graft_entry__ as g; g.build()"; python scripts/bench_mlp_trace.py
"""
import os, sys; sys.path.insert(0, "/root/repo")
import torch, d2s
ops = d2s.pkg.ops
D, T = 384, 197
M = 1024 * T
h = torch.randn(M, D, device="cuda", dtype=torch.bfloat16); x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
fc1 = torch.nn.Linear(D, 4 * D).cuda().bfloat16(); fc2 = torch.nn.Linear(4 * D, D).cuda().bfloat16(); ln = torch.nn.LayerNorm(D, eps=1e-6).cuda().bfloat16()
buf = torch.zeros(74 * 6 * 8, dtype=torch.int64, device="cuda")
f = lambda: ops.mlp_residual_ln(h, fc1.weight, fc1.bias, fc2.weight, fc2.bias, x, ln.weight, ln.bias, 1e-6)
for _ in range(3): f()
os.environ["D2S_GEMM_TRACE"] = str(buf.data_ptr())
f(); torch.cuda.synchronize()
t = buf[:74 * 16].view(74, 2, 8).double()
names = ["issue+other", "wait s_empty", "wait a1_full", "wait w1_full", "wait p_full", "wait w2_full", "wait acc_empty"]
for w, nm in ((0, "G1 issuer"), (1, "G2 issuer")):
    ch = t[:, w, 7].mean().item()
    print(f"{nm}: {ch:.0f} chunks; cycles per chunk: " + ", ".join(f"{n} {t[:, w, i].mean().item() / ch:.0f}" for i, n in enumerate(names) if t[:, w, i].sum() > 0))
