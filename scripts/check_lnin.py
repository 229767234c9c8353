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


# ---- norm1: MLP kernel writes row statistics, the next qkv projection normalises its resident input rows itself ----
def chain_old(a, x):
    xs, st = ops.linear_residual_ln(a, Wp, bp, x, eps=1e-6, want_norm=False, want_stats=True)
    x2, hn = ops.mlp_residual_ln(None, W1, b1, W2, b2, xs, g1, bt1, 1e-6, in_stats=st, in_ln_weight=g2, in_ln_bias=bt2)
    return x2, ops.linear_act(hn, Wq, bq, ops.ACT_NONE)


def chain_new(a, x):
    xs, st = ops.linear_residual_ln(a, Wp, bp, x, eps=1e-6, want_norm=False, want_stats=True)
    x2, st2 = ops.mlp_residual_ln(None, W1, b1, W2, b2, xs, None, None, 1e-6, want_norm=False, in_stats=st, in_ln_weight=g2,
                                  in_ln_bias=bt2, want_stats=True)
    return x2, ops.linear_act(x2, Wq, bq, ops.ACT_NONE, in_stats=st2, in_ln_weight=g1, in_ln_bias=bt1)


Wq, bq = r(3 * D, D, sc=D ** -0.5), r(3 * D, sc=0.1)
ok2 = True
for (B, T) in [(3, 197), (2, 138), (7, 97), (1, 1), (5, 68), (64, 197), (300, 97), (1024, 68)]:
    a, x = r(B, T, D), r(B, T, D, sc=2.0)
    o, n = chain_old(a, x), chain_new(a, x)
    same = all(torch.equal(p, q) for p, q in zip(o, n))
    print(f"norm1 chain B={B} T={T}: identical {same}")
    ok2 &= same
print("OK" if ok2 else "MISMATCH")
if "--time" in sys.argv:
    for T in (197, 138, 97, 68):
        xs = [(r(1024, T, D), r(1024, T, D)) for _ in range(3)]
        print(f"T={T}: proj -> MLP -> qkv with hn materialised {timed(chain_old, xs):.1f} us | with row statistics {timed(chain_new, xs):.1f} us")
sys.exit(0 if (ok and ok2) else 1)
