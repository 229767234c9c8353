"""The one-kernel MLP at B = 1024, T = 197 with a materialised input (hn) and with the input LayerNorm applied inside (x' + row
statistics), eager launches for ncu; prints CUDA-event timings of both (rotating inputs) when run without a profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
bf = torch.bfloat16
D, HID, M = 384, 1536, 1024 * 197
r = lambda *s, sc=1.0: (torch.randn(*s, device="cuda") * sc).to(bf)
W1, b1, W2, b2 = r(HID, D, sc=D ** -0.5), r(HID, sc=0.1), r(D, HID, sc=HID ** -0.5), r(D, sc=0.1)
g, bt = torch.ones(D, device="cuda", dtype=bf), torch.zeros(D, device="cuda", dtype=bf)
sets = []
for _ in range(3):
    x = r(M, D)
    st = torch.stack([x.float().mean(-1), torch.rsqrt(x.float().var(-1, unbiased=False) + 1e-6)], dim=-1).contiguous()
    hn = torch.nn.functional.layer_norm(x.float(), (D,)).to(bf)
    sets.append((x, st, hn))
for mode in ("hn", "lnin"):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(3):
        if rep == 1:
            ev[0].record()
        for x, st, hn in sets:
            if mode == "hn":
                ops.mlp_residual_ln(hn, W1, b1, W2, b2, x, g, bt, 1e-6)
            else:
                ops.mlp_residual_ln(None, W1, b1, W2, b2, x, g, bt, 1e-6, in_stats=st, in_ln_weight=g, in_ln_bias=bt)
    ev[1].record()
    torch.cuda.synchronize()
    print(mode, ev[0].elapsed_time(ev[1]) / 6 * 1e3, "us per launch")
