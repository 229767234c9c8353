"""Training-step kernels added in round 2, each timed alone inside a CUDA graph on rotating argument sets larger than twice the
L2 (bench.time_graphed): flat AdamW, token-KL rows, the predictors' local / global split (forward / backward), LayerNorm over
x[:, 1:] (forward / backward).  One JSON line per kernel with algorithmic bytes and the fraction of the measured HBM peak.

    python scripts/bench_train_kernels.py > gpurun_out/train_kernels.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

ops = d2s.pkg.ops
dev = torch.device("cuda", 0)
pk = bench.peaks()
B, N, D = 256, 196, 384


def report(name, shape, calls, nbytes):
    ms = bench.time_graphed(calls, torch)
    print(json.dumps({"kernel": name, "shape": shape, "us": ms * 1e3, "algo_bytes": nbytes, "gbs": nbytes / ms / 1e6,
                      "frac_hbm": nbytes / ms / 1e6 / pk["hbm"]}), flush=True)


def sets(nbytes, make):
    return [make() for _ in range(bench.nsets_for(nbytes))]


with torch.no_grad():
    # ---- AdamW over 22.8 M parameters (DeiT-S + predictors): 30 bytes per parameter
    n = 22_774_432
    nb = 30 * n
    S = sets(nb, lambda: (torch.randn(n, device=dev), torch.randn(n, device=dev) * 1e-3, torch.zeros(n, device=dev), torch.zeros(n, device=dev),
                          torch.empty(n, dtype=torch.bfloat16, device=dev)))
    lr = torch.full((1,), 5e-4, device=dev)
    step = torch.full((1,), 3.0, device=dev)
    report("adamw_flat", f"{n} parameters", [lambda a=a: ops.adamw_flat(a[0], a[1], a[2], a[3], a[4], 0, n, lr, step, 0.9, 0.999, 1e-8, 0.05) for a in S], nb)
    del S
    # ---- token KL: (B, 197, 384) fp32 student view x[:, 1:], fp32 teacher view
    nb = B * N * D * (4 + 4 + 4)
    S = sets(nb, lambda: (torch.randn(B, N + 1, D, device=dev), torch.randn(B, N + 1, D, device=dev)))
    report("token_kl_rows", f"B={B},N={N},C={D} f32/f32", [lambda a=a: ops.token_kl_rows(a[0][:, 1:], a[1][:, 1:]) for a in S], nb)
    S = sets(nb, lambda: (torch.randn(B, N + 1, D, device=dev), torch.randn(B, N + 1, D, device=dev).bfloat16()))
    report("token_kl_rows", f"B={B},N={N},C={D} f32/bf16", [lambda a=a: ops.token_kl_rows(a[0][:, 1:], a[1][:, 1:]) for a in S], B * N * D * (4 + 2 + 4))
    del S
    # ---- local / global split
    nb = 2 * 2 * B * N * D
    pol = (torch.rand(B, N, 1, device=dev) > 0.3).float()
    S = sets(nb, lambda: torch.randn(B, N, D, device=dev).bfloat16())
    report("pool_concat_fwd", f"B={B},N={N},C={D} bf16", [lambda a=a: ops.pool_concat_train(a, pol) for a in S], nb)
    del S
torch.cuda.empty_cache()
# backward kernels: called at the C-ABI level (autograd's engine would run them on the forward's stream, outside the capture)
_c, _p, _st, _dc = ops._call, ops._ptr, ops._stream, ops._dtype_code
with torch.no_grad():
    nb = 3 * 2 * B * N * D
    pol = (torch.rand(B, N, device=dev) > 0.3).float()

    def mk_pool():
        h = torch.randn(B, N, D, device=dev).bfloat16()
        out = torch.empty_like(h)
        pooled = torch.empty(B, D // 2, device=dev)
        wsum = torch.empty(B, device=dev)
        _c("d2s_pool_concat_fwd", _p(h), _p(pol), _dc(h), B, N, D, _p(out), _p(pooled), _p(wsum), _st(h))
        return h, torch.randn_like(h), pooled, wsum, torch.empty_like(h), torch.empty(B, N, device=dev)
    S = sets(nb, mk_pool)
    report("pool_concat_bwd", f"B={B},N={N},C={D} bf16",
           [lambda a=a: _c("d2s_pool_concat_bwd", _p(a[1]), _p(a[0]), _p(pol), _p(a[2]), _p(a[3]), _dc(a[0]), B, N, D, _p(a[4]), _p(a[5]), _st(a[0]))
            for a in S], nb)
    del S
    w = torch.ones(D, device=dev)
    b = torch.zeros(D, device=dev)
    nb = 2 * B * N * D + 2 * B * N * D
    S = sets(nb, lambda: torch.randn(B, N + 1, D, device=dev).bfloat16())
    report("layer_norm(x[:, 1:]) fwd", f"B={B},T={N + 1},D={D} bf16", [lambda a=a: ops.layer_norm(a, w, b, 1e-5, out_dtype=torch.bfloat16, row0=1) for a in S], nb)
    del S
    nb = 3 * 2 * B * N * D

    def mk_ln():
        x = torch.randn(B, N + 1, D, device=dev).bfloat16()
        h = torch.empty(B, N, D, device=dev, dtype=torch.bfloat16)
        stats = torch.empty(B * N, 2, device=dev)
        _c("d2s_layernorm_seg_fwd", _p(x), _dc(x), _p(w), _p(b), B * N, D, N, 1, 1e-5, _p(h), _dc(h), _p(stats), _st(x))
        return x, torch.randn_like(h), stats, torch.empty_like(x), torch.zeros(2, D, device=dev)
    S = sets(nb, mk_ln)
    report("layer_norm(x[:, 1:]) bwd", f"B={B},T={N + 1},D={D} bf16",
           [lambda a=a: _c("d2s_layernorm_seg_bwd", _p(a[1]), _dc(a[1]), _p(a[0]), _dc(a[0]), _p(a[2]), _p(w), B * N, D, N, 1, _p(a[3]), _p(a[4][0]),
                           _p(a[4][1]), _st(a[0])) for a in S], nb)
