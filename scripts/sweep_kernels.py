"""Standalone kernel sweep (BASELINE.json configs[3] and configs[4]): select, gather fwd/bwd, fused tails, PerturbedTopK
fwd/bwd, softmax_with_policy fwd/bwd, add+LayerNorm -- CUDA-event timing, algorithmic bytes / time vs
the measured HBM peak.  Prints a markdown table; `--json PATH` also dumps the rows.

    python scripts/sweep_kernels.py [--quick] [--json gpurun_out/sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--json", default=None)
args = ap.parse_args()
ops = d2s.pkg.ops
dev = torch.device("cuda:0")
PK = bench.peaks()
FLUSH = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
rows = []


def timed(fn, foot_bytes, iters=20):
    """Average ms per launch.  Working sets larger than L2 are timed back to back; smaller ones get an L2 flush
    (a 160 MB memset) between launches, with each launch bracketed by its own pair of events."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if foot_bytes > 200e6:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters, "inputs > L2"
    tot = 0.0
    for _ in range(iters):
        FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters, "L2 flushed"


def add(kernel, shape, algo_bytes, fn, foot=None):
    ms, how = timed(fn, algo_bytes if foot is None else foot)
    gbs = algo_bytes / ms / 1e6
    rows.append(dict(kernel=kernel, shape=shape, us=ms * 1e3, algo_MB=algo_bytes / 1e6, GBs=gbs, frac=gbs / PK["hbm"], l2=how))
    print(f"| {kernel} | {shape} | {ms * 1e3:.1f} | {algo_bytes / 1e6:.2f} | {gbs:.0f} | {100 * gbs / PK['hbm']:.1f} % | {how} |", flush=True)


print(f"HBM peak {PK['hbm']} GB/s ({PK['source']})\n")
print("| kernel | shape | us | algorithmic MB | GB/s | of HBM peak | L2 |")
print("|---|---|---|---|---|---|---|")
N = 196
Bs = [64, 1024] if args.quick else [1, 8, 64, 512, 1024, 4096]
for D in (384, 768):
    for ratio in ((0.7,) if args.quick else (0.3, 0.5, 0.7, 0.9)):
        K = int(N * ratio)
        for B in Bs:
            if D == 768 and B == 4096 and not args.quick:
                pass
            for dt, e in ((torch.bfloat16, 2),):
                x = torch.randn(B, N + 1, D, device=dev, dtype=dt)
                sc = torch.rand(B, N, device=dev)
                kept, dropped = ops.select_topk(sc, K, ops.ORDER_INDEX_ASC)
                if D == 384:
                    add("select_topk (index asc, +dropped)", f"B={B} N={N} K={K}", B * (4 * N + 8 * N),
                        lambda: ops.select_topk(sc, K, ops.ORDER_INDEX_ASC))
                add("gather_tokens fwd", f"B={B} T={N + 1} D={D} K={K} bf16", B * (2 * e * D * (K + 1) + 8 * (K + 1)),
                    lambda: ops.gather_tokens(x, kept))
                g = torch.randn(B, K + 1, D, device=dev, dtype=dt)
                add("scatter_tokens bwd", f"B={B} T={N + 1} D={D} K={K} bf16", B * (e * D * ((K + 1) + (N + 1)) + 8 * (K + 1)),
                    lambda: ops.scatter_tokens_bwd(g, kept, N + 1))
                del x, g
# fused tails
for B in ([1024] if args.quick else [64, 1024, 4096]):
    for C in (96, 192):
        hid = torch.randn(B, N, C, device=dev, dtype=torch.bfloat16)
        W2, b2 = torch.randn(2, C, device=dev) * 0.1, torch.zeros(2, device=dev)
        add("score_tail_a + select", f"B={B} N={N} C={C} K=137 bf16", B * (2 * N * C + 8 * N + 8 * 137),
            lambda: ops.score_tail_a(hid, W2, b2, k=137))
        lw, lb, W1, b1 = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.randn(1, C, device=dev) * 0.1, torch.zeros(1, device=dev)
        add("score_tail_b + select", f"B={B} N={N} C={C} K=137 bf16", B * (2 * N * C + 8 * N + 8 * N),
            lambda: ops.score_tail_b(hid, lw, lb, W1, b1, 137))
        del hid
# PerturbedTopK (configs[3]): DeiT-B tokens N=196, k=98, 500 samples
for B in ([8, 64] if args.quick else [1, 8, 64, 256]):
    S, K = 500, 98
    x = torch.softmax(torch.randn(B, N, device=dev), -1)
    noise = torch.randn(B, S, N, device=dev)
    xg = x.clone().requires_grad_(True)
    add("ptopk fwd (injected noise, +egrad)", f"B={B} N={N} k={K} S={S}", B * (4 * N + 4 * S * N + 8 * K * N),
        lambda: ops.perturbed_topk(xg, K, S, 0.05, noise=noise))
    add("ptopk fwd (in-kernel Philox)", f"B={B} N={N} k={K} S={S}", B * (4 * N + 8 * K * N),
        lambda: ops.perturbed_topk(xg, K, S, 0.05, seed=7))
    out = ops.perturbed_topk(xg, K, S, 0.05, noise=noise)
    gout = torch.randn_like(out)
    add("ptopk bwd", f"B={B} N={N} k={K}", B * (8 * K * N + 4 * N), lambda: torch.autograd.grad(out, xg, gout, retain_graph=True))
# softmax_with_policy fwd/bwd (training path), add+LayerNorm
for B, T in ([(256, 197)] if args.quick else [(64, 197), (256, 197), (256, 138)]):
    H = 6
    s = torch.randn(B, H, T, T, device=dev, dtype=torch.bfloat16)
    pol = (torch.rand(B, T, 1, device=dev) > 0.3).float()
    add("softmax_with_policy fwd", f"B={B} H={H} T={T} bf16", B * H * T * T * 2 * 2, lambda: ops.softmax_with_policy(s, pol))
    sg = s.clone().requires_grad_(True)
    pg = pol.clone().requires_grad_(True)
    o = ops.softmax_with_policy(sg, pg)
    go = torch.randn_like(o)
    add("softmax_with_policy bwd", f"B={B} H={H} T={T} bf16", B * H * T * T * 2 * 3,
        lambda: torch.autograd.grad(o, (sg, pg), go, retain_graph=True))
    del s, sg, o, go
for B, T, D in ([(1024, 197, 384)] if args.quick else [(1024, 197, 384), (1024, 68, 384), (256, 197, 768)]):
    x = torch.randn(B, T, D, device=dev, dtype=torch.bfloat16)
    y = torch.randn(B, T, D, device=dev, dtype=torch.bfloat16)
    w, b = torch.ones(D, device=dev, dtype=torch.bfloat16), torch.zeros(D, device=dev, dtype=torch.bfloat16)
    add("add_layernorm", f"B={B} T={T} D={D} bf16", B * T * D * 2 * 4, lambda: ops.add_layernorm(x, y, w, b, 1e-6))
if args.json:
    with open(args.json, "w") as fh:
        json.dump(dict(peaks=PK, rows=rows), fh, indent=1)
