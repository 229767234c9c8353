"""Host-side enqueue cost (us per call, tiny tensors so the GPU never throttles the loop) of a few d2s ops against torch ops."""
import sys, time; sys.path.insert(0, "/root/repo")
import torch, d2s
ops, lib = d2s.pkg.ops, d2s.pkg._lib
def host(fn, n=2000):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6
dy = torch.randn(64, 384, device="cuda").bfloat16(); x = torch.randn(64, 384, device="cuda").bfloat16()
out = torch.empty(384, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("torch sum(0)          ", round(host(lambda: dy.sum(0)), 1))
print("torch dy.t() @ x      ", round(host(lambda: dy.t() @ x), 1))
print("ops.colsum            ", round(host(lambda: ops.colsum(dy)), 1))
print("raw d2s_colsum_bf16   ", round(host(lambda: lib.call("d2s_colsum_bf16", dy.data_ptr(), 64, 384, out.data_ptr(), st)), 1))
print("ops.linear_wgrad      ", round(host(lambda: ops.linear_wgrad(dy, x)), 1))
w = torch.randn(384, device="cuda"); b = torch.randn(384, device="cuda")
print("ops.layer_norm        ", round(host(lambda: ops.layer_norm(x, w, b, 1e-6)), 1))
print("torch.empty           ", round(host(lambda: torch.empty(384, device="cuda")), 1))
