"""proj + row statistics -> MLP (norm2 inside, row statistics out) -> qkv projection (norm1 inside) at B = 1024, T = 197, a few eager
launches for ncu (the three kernels of a transformer block around the attention)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
bf = torch.bfloat16
D, HID, B, T = 384, 1536, 1024, 197
r = lambda *s, sc=1.0: (torch.randn(*s, device="cuda") * sc).to(bf)
Wp, bp = r(D, D, sc=D ** -0.5), r(D, sc=0.1)
W1, b1, W2, b2 = r(HID, D, sc=D ** -0.5), r(HID, sc=0.1), r(D, HID, sc=HID ** -0.5), r(D, sc=0.1)
Wq, bq = r(3 * D, D, sc=D ** -0.5), r(3 * D, sc=0.1)
g, bt = torch.ones(D, device="cuda", dtype=bf), torch.zeros(D, device="cuda", dtype=bf)
for _ in range(3):
    a, x = r(B, T, D), r(B, T, D)
    xs, st = ops.linear_residual_ln(a, Wp, bp, x, eps=1e-6, want_norm=False, want_stats=True)
    x2, st2 = ops.mlp_residual_ln(None, W1, b1, W2, b2, xs, None, None, 1e-6, want_norm=False, in_stats=st, in_ln_weight=g, in_ln_bias=bt,
                                  want_stats=True)
    qkv = ops.linear_act(x2, Wq, bq, ops.ACT_NONE, in_stats=st2, in_ln_weight=g, in_ln_bias=bt)
torch.cuda.synchronize()
print("ok")
