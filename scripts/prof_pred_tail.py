"""A few launches of d2s_predictor_a_tail_bf16 (second / third Linear + GELUs + Linear(., 2) + log-softmax + selection) at the three
stage shapes of the bench configuration, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
B, H = 1024, 192
bf = torch.bfloat16
w2, w3 = (torch.randn(H, 2 * H, device="cuda") * 0.05).to(bf), (torch.randn(H // 2, H, device="cuda") * 0.07).to(bf)
b3 = torch.zeros(H // 2, device="cuda", dtype=bf)
W, b = torch.randn(2, H // 2, device="cuda") * 0.1, torch.zeros(2, device="cuda")
pim = (torch.randn(B, H, device="cuda") * 0.3).to(bf)
for N, K in ((196, 137), (137, 96), (96, 67)):
    prev = (torch.rand(B, N, device="cuda") > 0.1).float()
    for _ in range(3):
        local = (torch.randn(B, N, H, device="cuda") * 0.5).to(bf)
        ops.predictor_a_tail(local, pim, w2, w3, b3, W, b, K, prev=prev)
torch.cuda.synchronize()
print("ok")
