"""Times the fused MLP kernel (fc1 + GELU + fc2 + residual + LayerNorm) against the two pair GEMMs it replaces."""
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
D = 384
for T in (197, 138, 97, 68):
    M = 1024 * T
    h = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    fc1 = torch.nn.Linear(D, 4 * D).cuda().bfloat16()
    fc2 = torch.nn.Linear(4 * D, D).cuda().bfloat16()
    ln = torch.nn.LayerNorm(D, eps=1e-6).cuda().bfloat16()
    def two():
        u = ops.linear_act(h, fc1.weight, fc1.bias, ops.ACT_GELU)
        return ops.linear_residual_ln(u, fc2.weight, fc2.bias, x, ln.weight, ln.bias, 1e-6)
    a = t(two)
    b = t(lambda: ops.mlp_residual_ln(h, fc1.weight, fc1.bias, fc2.weight, fc2.bias, x, ln.weight, ln.bias, 1e-6))
    fl = 4.0 * M * D * 4 * D
    print(f"T={T}: fc1+GELU then fc2+add+LN {a:.1f} us | fused MLP {b:.1f} us ({fl / b / 1e6:.0f} TF/s)")
