"""d2s_predictor_a_fused_bf16 against the unfused path (library GEMMs + pool_act + bias_act + score_tail_a) and an fp32 torch
evaluation of the same bf16 weights; then timings of both at the three stage shapes of the bench configuration.

    python scripts/check_pred_fused.py [--time]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import d2s  # noqa: E402

eng, ops = d2s.engine, d2s.ops


def fp32_reference(pred, normed, prev, k):
    f = lambda t: t.detach().float()
    x = f(normed)
    z = torch.nn.functional.gelu(x @ f(pred.in_conv[1].weight).t() + f(pred.in_conv[1].bias))
    B, N, C = z.shape
    pol = torch.ones(B, N, 1, device=x.device) if prev is None else prev.reshape(B, N, 1).float()
    pooled = (z[:, :, C // 2:] * pol).sum(1, keepdim=True) / pol.sum(1, keepdim=True)
    h = torch.cat([z[:, :, :C // 2], pooled.expand(B, N, C // 2)], -1)
    oc = pred.out_conv
    h = torch.nn.functional.gelu(h @ f(oc[0].weight).t() + f(oc[0].bias))
    h = torch.nn.functional.gelu(h @ f(oc[2].weight).t() + f(oc[2].bias))
    return torch.log_softmax(h @ f(oc[4].weight).t() + f(oc[4].bias), -1)


def run(B, N, k, with_prev, seed):
    torch.manual_seed(seed)
    dev = torch.device("cuda")
    pred = d2s.variant_a.PredictorLG(384).to(dev)
    for p in pred.parameters():
        torch.nn.init.normal_(p, std=0.08 if p.dim() > 1 else 0.05)
    pred = pred.to(torch.bfloat16).eval()
    normed = torch.randn(B, N, 384, device=dev).to(torch.bfloat16)
    prev = (torch.rand(B, N, device=dev) > 0.3).float() if with_prev else None
    with torch.no_grad():
        eng._PRED_FUSED = False
        lp_u, kept_u, pk_u = eng.predictor_a_select(pred, normed, prev, k)
        eng._PRED_FUSED = True
        lp_f, kept_f, pk_f = eng.predictor_a_select(pred, normed, prev, k)
        ref = fp32_reference(pred, normed, prev, k)
    torch.cuda.synchronize()
    eu = (lp_u - ref).abs().max().item()
    ef = (lp_f - ref).abs().max().item()
    same = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(kept_u, kept_f)) / max(1, kept_u.numel())
    # the kept list must be exactly the stable descending top-k of the kernel's own scores
    order = torch.sort(lp_f[:, :, 0], dim=1, descending=True, stable=True).indices[:, :k]
    exact = torch.equal(order, kept_f)
    pk_ok = True if prev is None else torch.equal(pk_f, torch.gather(prev, 1, kept_f))
    print(f"B={B} N={N} k={k} prev={with_prev}: |logp - fp32| unfused {eu:.3e} fused {ef:.3e}; kept overlap {same:.4f}; "
          f"kept == sort(own scores) {exact}; prev_kept ok {pk_ok}; finite {bool(torch.isfinite(lp_f).all())}")
    return exact and pk_ok and ef < 5 * max(eu, 1e-2)


def timeit(B, N, k):
    dev = torch.device("cuda")
    pred = d2s.variant_a.PredictorLG(384).to(dev).to(torch.bfloat16).eval()
    sets = [torch.randn(B, N, 384, device=dev).to(torch.bfloat16) for _ in range(6)]
    out = {}
    for fused in (False, True):
        eng._PRED_FUSED = fused
        with torch.no_grad():
            for x in sets:
                eng.predictor_a_select(pred, x, None, k)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for x in sets:
                    eng.predictor_a_select(pred, x, None, k)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            out[fused] = e0.elapsed_time(e1) / (5 * len(sets)) * 1e3
    print(f"B={B} N={N}: unfused {out[False]:.1f} us, fused {out[True]:.1f} us")


if __name__ == "__main__":
    ok = True
    for (B, N, k, wp, seed) in [(3, 196, 137, False, 0), (5, 137, 96, True, 1), (4, 96, 67, True, 2), (2, 128, 64, False, 3),
                                (2, 129, 70, True, 4), (1, 17, 5, True, 5), (300, 196, 137, True, 6), (2, 256, 200, True, 7)]:
        ok &= run(B, N, k, wp, seed)
    print("OK" if ok else "MISMATCH")
    if "--time" in sys.argv:
        for (N, k) in [(196, 137), (137, 96), (96, 67)]:
            timeit(1024, N, k)
    sys.exit(0 if ok else 1)
