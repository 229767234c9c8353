"""qkv projection + attention over the whole batch at once against the same two kernels run chunk by chunk through ONE small qkv
buffer that stays in L2 (is the qkv tensor's HBM round trip avoidable without fusing the kernels?).  B = 1024, CUDA-graph timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, d2s
ops = d2s.pkg.ops
bf = torch.bfloat16
D, H, B = 384, 6, 1024
W = (torch.randn(3 * D, D, device="cuda") / D ** 0.5).to(bf)
b = (torch.randn(3 * D, device="cuda") * 0.1).to(bf)


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for T in (197, 138, 97, 68):
    xs = [torch.randn(B, T, D, device="cuda").to(bf) for _ in range(3)]

    def full():
        for x in xs:
            ops.attention_core(ops.linear_act(x, W, b, ops.ACT_NONE), H)

    t_full = timed(full) / len(xs)
    res = [f"T={T}: whole batch {t_full:.1f} us"]
    for C in (64, 96, 128, 192, 256):
        def chunked():
            for x in xs:
                for c0 in range(0, B, C):
                    ops.attention_core(ops.linear_act(x[c0:c0 + C], W, b, ops.ACT_NONE), H)     # the allocator hands back the same block
        res.append(f"chunks of {C}: {timed(chunked) / len(xs):.1f} us")
    print(" | ".join(res))
