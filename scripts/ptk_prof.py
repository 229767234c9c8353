import sys; sys.path.insert(0, "/root/repo")
import torch, d2s
ops = d2s.pkg.ops
B, N, K, S = 256, 196, 98, 500
x = torch.softmax(torch.randn(B, N, device="cuda"), -1)
noise = torch.randn(B, S, N, device="cuda")
for _ in range(3):
    out = ops.perturbed_topk(x, K, S, 0.05, noise=noise)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    out = ops.perturbed_topk(x, K, S, 0.05, noise=noise)
e.record(); torch.cuda.synchronize()
print("ptopk fwd us", s.elapsed_time(e) * 100)
