"""Times ops.attention_core of the tree given as argv[1] (graph-captured launches over rotating buffers > 2x L2)."""
import sys, os
root = os.path.abspath(sys.argv[1])
sys.path.insert(0, root)
import torch
import d2s
ops = d2s.ops
dev = torch.device("cuda:0")
H = 6
out = []
for T in (197, 138, 97, 68):
    B = 1024
    nsets = 4
    qs = [(torch.randn(B, T, 3 * H * 64, device=dev) * float(sys.argv[2] if len(sys.argv) > 2 else 0.7)).bfloat16() for _ in range(nsets)]
    for q in qs:
        ops.attention_core(q, H)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for q in qs:
            ops.attention_core(q, H)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(6):
            for q in qs:
                ops.attention_core(q, H)
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record(); torch.cuda.synchronize()
    out.append(round(s.elapsed_time(e) / (5 * 6 * nsets) * 1e3, 1))
print(os.path.basename(root), out, flush=True)
