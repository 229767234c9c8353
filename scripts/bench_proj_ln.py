"""proj + residual + LayerNorm pair GEMM with and without the LayerNorm output (what not materialising hn would buy), B = 1024."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
bf = torch.bfloat16
D = 384
W = (torch.randn(D, D, device="cuda") / D ** 0.5).to(bf)
b = (torch.randn(D, device="cuda") * 0.1).to(bf)
g, bt = torch.ones(D, device="cuda", dtype=bf), torch.zeros(D, device="cuda", dtype=bf)


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


for T in (197, 138, 97, 68):
    M = 1024 * T
    xs = [(torch.randn(M, D, device="cuda").to(bf), torch.randn(M, D, device="cuda").to(bf)) for _ in range(4)]
    t_ln = timed(lambda a, x: ops.linear_residual_ln(a, W, b, x, g, bt, 1e-6, want_norm=True), xs)
    t_no = timed(lambda a, x: ops.linear_residual_ln(a, W, b, x, None, None, 1e-6, want_norm=False), xs)
    print(f"T={T}: with LayerNorm output {t_ln:.1f} us, without {t_no:.1f} us")
