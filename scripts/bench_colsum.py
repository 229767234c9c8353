"""Bias-gradient column sums at the training shapes: d2s_colsum_bf16 vs torch.sum(0) vs the cuBLASLt bias-gradient epilogue."""
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
M = 256 * 197
for N, K in ((1536, 384), (1152, 384), (384, 1536), (384, 384)):
    dy = torch.randn(M, N, device="cuda").bfloat16(); x = torch.randn(M, K, device="cuda").bfloat16()
    a = t(lambda: ops.colsum(dy)); b = t(lambda: dy.sum(0))
    c = t(lambda: ops.linear_wgrad(dy, x)); d = t(lambda: dy.t() @ x)
    by = M * N * 2
    print(f"N={N} K={K}: d2s colsum {a:.1f} us ({by / a / 1e3:.0f} GB/s) | torch sum(0) {b:.1f} us | cuBLASLt dW+db {c:.1f} us | torch dW alone {d:.1f} us")
u = torch.randn(M, 1536, device="cuda").bfloat16(); ga = torch.randn(M, 1536, device="cuda").bfloat16()
ug = u.clone().requires_grad_(True)
y = torch.nn.functional.gelu(ug)
a = t(lambda: ops.gelu_bwd_colsum(u, ga))
b = t(lambda: torch.autograd.grad(y, ug, ga, retain_graph=True))
print(f"GELU backward (50432 x 1536): d2s fused with the column sums {a:.1f} us | torch gelu_backward alone {b:.1f} us (+ colsum above)")
