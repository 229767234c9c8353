"""A few launches of the PerturbedTopK forward (BASELINE configs[3]: N=196, k=98, 500 samples, injected noise) at B=256, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
lib = d2s.pkg._lib
B, N, K, S = 256, 196, 98, 500
x = torch.softmax(torch.randn(B, N, device="cuda"), -1)
nz = [torch.randn(B, S, N, device="cuda") for _ in range(3)]
ind, eg = torch.empty(B, K, N, device="cuda"), torch.empty(B, K, N, device="cuda")
for n in nz * 2:
    lib.call("d2s_ptopk_fwd", x.data_ptr(), n.data_ptr(), B, N, K, S, 0.05, ind.data_ptr(), eg.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(ind.sum()))
