"""A few launches of the fused predictor tail (score_tail_a with GELU on load + top-K + prev gather) at the bench shape, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
B, N, C, K = 1024, 196, 96, 137
hs = [torch.randn(B, N, C, device="cuda", dtype=torch.bfloat16) for _ in range(8)]
W, b = torch.randn(2, C, device="cuda") * 0.1, torch.zeros(2, device="cuda")
for h in hs:
    ops.score_tail_a(h, W, b, k=K, act_input=ops.ACT_GELU, want_prev_kept=True)
torch.cuda.synchronize()
print("ok")
