import os, sys; sys.path.insert(0, "/root/repo")
import torch, d2s
ops = d2s.pkg.ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
D, T = 384, 197
M = 1024 * T
x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
fc1 = torch.nn.Linear(D, 4 * D).cuda().bfloat16()
fc2 = torch.nn.Linear(4 * D, D).cuda().bfloat16()
u = torch.randn(M, 4 * D, device="cuda", dtype=torch.bfloat16)
ln = torch.nn.LayerNorm(D, eps=1e-6).cuda().bfloat16()
for dbg in (0, 1, 2, 3, 8, 9, 10, 11):
    os.environ["D2S_GEMM_DEBUG"] = str(dbg)
    a = t(lambda: ops.linear_act(x, fc1.weight, fc1.bias, ops.ACT_NONE, pair=True))
    g = t(lambda: ops.linear_act(x, fc1.weight, fc1.bias, ops.ACT_GELU, pair=True))
    f = t(lambda: ops.linear_residual_ln(u, fc2.weight, fc2.bias, x, ln.weight, ln.bias, 1e-6))
    print(f"dbg={dbg}: fc1 pair no-act {a:.1f} us, gelu {g:.1f} us | fc2+LN {f:.1f} us")
