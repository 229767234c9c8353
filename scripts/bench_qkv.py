"""qkv projection (N = 1152, K = 384) on the CTA-pair tcgen05 GEMM with 192-column tiles against the library GEMM, at the four token
counts of the bench step (B = 1024); CUDA-graph timed on rotating inputs larger than the L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
bf = torch.bfloat16
D = int(sys.argv[1]) if len(sys.argv) > 1 else 384
W = (torch.randn(3 * D, D, device="cuda") / D ** 0.5).to(bf)
b = (torch.randn(3 * D, device="cuda") * 0.1).to(bf)


def timed(fn, xs):
    for x in xs:
        fn(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for x in xs:
            fn(x)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * len(xs)) * 1e3


for T in (197, 138, 97, 68):
    M = 1024 * T
    xs = [torch.randn(M, D, device="cuda").to(bf) for _ in range(4)]
    ref = torch.nn.functional.linear(xs[0], W, b)
    out = ops.linear_act(xs[0], W, b, ops.ACT_NONE)
    err = (out.float() - ref.float()).abs().max().item()
    t_lib = timed(lambda x: torch.nn.functional.linear(x, W, b), xs)
    t_d2s = timed(lambda x: ops.linear_act(x, W, b, ops.ACT_NONE), xs)
    fl = 2.0 * M * D * 3 * D
    print(f"T={T} M={M}: library {t_lib:.1f} us ({fl / t_lib / 1e6:.0f} TF/s) | pair GEMM 192-col tiles {t_d2s:.1f} us ({fl / t_d2s / 1e6:.0f} TF/s) | max |diff| {err:.3e}")
