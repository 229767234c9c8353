"""proj + residual + LayerNorm -> MLP kernel, with the normalised copy (as before) and with per-row statistics + the LayerNorm applied
inside the MLP kernel: outputs must be bit-identical; timings of both pairs at the four token counts of the bench step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
bf = torch.bfloat16
D, HID = 384, 1536
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).to(bf)
Wp, bp = r(D, D, sc=D ** -0.5), r(D, sc=0.1)
W1, b1, W2, b2 = r(HID, D, sc=D ** -0.5), r(HID, sc=0.1), r(D, HID, sc=HID ** -0.5), r(D, sc=0.1)
g2, bt2 = (1 + 0.2 * torch.randn(D, device="cuda", generator=g)).to(bf), r(D, sc=0.2)
g1, bt1 = (1 + 0.2 * torch.randn(D, device="cuda", generator=g)).to(bf), r(D, sc=0.2)


def old(a, x, row0=0):
    xs, hn = ops.linear_residual_ln(a, Wp, bp, x, g2, bt2, 1e-6)
    return ops.mlp_residual_ln(hn, W1, b1, W2, b2, xs, g1, bt1, 1e-6, norm_row0=row0)


def new(a, x, row0=0):
    xs, st = ops.linear_residual_ln(a, Wp, bp, x, eps=1e-6, want_norm=False, want_stats=True)
    return ops.mlp_residual_ln(None, W1, b1, W2, b2, xs, g1, bt1, 1e-6, norm_row0=row0, in_stats=st, in_ln_weight=g2, in_ln_bias=bt2)


ok = True
for (B, T, row0) in [(3, 197, 0), (2, 138, 1), (7, 97, 0), (1, 1, 0), (5, 68, 0), (64, 197, 1), (300, 97, 0)]:
    a, x = r(B, T, D), r(B, T, D, sc=2.0)
    x = x + 3.0 * (torch.arange(D, device="cuda") % 7 == 0).to(bf)          # a few channels with a large mean
    o, n = old(a, x, row0), new(a, x, row0)
    same = all(torch.equal(p, q) for p, q in zip(o, n))
    fin = all(bool(torch.isfinite(q.float()).all()) for q in n)
    print(f"B={B} T={T} row0={row0}: identical {same}, finite {fin}")
    ok &= same and fin
print("OK" if ok else "MISMATCH")


def timed(fn, xs):
    for x in xs:
        fn(*x)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for x in xs:
            fn(*x)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * len(xs)) * 1e3


if "--time" in sys.argv:
    for T in (197, 138, 97, 68):
        xs = [(r(1024, T, D), r(1024, T, D)) for _ in range(3)]
        print(f"T={T}: proj+LN -> MLP {timed(old, xs):.1f} us | proj+stats -> MLP with the LayerNorm inside {timed(new, xs):.1f} us")
sys.exit(0 if ok else 1)
