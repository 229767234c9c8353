"""Fused MLP kernel: time against the number of row tiles per CTA pair (M = 74 pairs x 256 rows x k) -> fixed cost + cost per tile."""
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
fc1 = torch.nn.Linear(D, 4 * D).cuda().bfloat16(); fc2 = torch.nn.Linear(4 * D, D).cuda().bfloat16(); ln = torch.nn.LayerNorm(D, eps=1e-6).cuda().bfloat16()
for k in (1, 2, 3, 4, 6, 8, 11, 16):
    M = 74 * 256 * k
    h = torch.randn(M, D, device="cuda", dtype=torch.bfloat16); x = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    us = t(lambda: ops.mlp_residual_ln(h, fc1.weight, fc1.bias, fc2.weight, fc2.bias, x, ln.weight, ln.bias, 1e-6))
    us0 = t(lambda: ops.mlp_residual_ln(h, fc1.weight, fc1.bias, fc2.weight, fc2.bias, x, None, None, 1e-6, want_norm=False))
    print(f"k={k:2d} M={M}: {us:.1f} us ({us / k:.1f} per tile) | no LN {us0:.1f} us")
