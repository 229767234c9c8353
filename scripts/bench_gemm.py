import sys; sys.path.insert(0, "/root/repo")
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
for T in (197, 138, 97, 68):
    M, K, N = 1024 * T, 384, 1536
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    lin = torch.nn.Linear(K, N).cuda().bfloat16()
    def ref():
        u = lin(x); ops.bias_act_(u, None, ops.ACT_GELU); return u
    a = t(ref); b = t(lambda: ops.linear_act(x, lin.weight, lin.bias, ops.ACT_GELU)); c = t(lambda: lin(x))
    fl = 2.0 * M * N * K
    print(f"T={T}: cuBLAS+GELU kernel {a:.1f} us | cuBLAS alone {c:.1f} us ({fl / c / 1e6:.0f} TF/s) | fused tcgen05 {b:.1f} us ({fl / b / 1e6:.0f} TF/s)")
M, K, N = 1024 * 197, 384, 1536
x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
lin = torch.nn.Linear(K, N).cuda().bfloat16()
for act, nm in ((0, "none"), (2, "relu"), (1, "gelu")):
    print(nm, round(t(lambda: ops.linear_act(x, lin.weight, lin.bias, act)), 1), "us")
