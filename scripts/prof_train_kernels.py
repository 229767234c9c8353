"""Launches the round-2 training kernels a few times each at the training step's shapes, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'adamw_flat|token_kl|pool_concat|ln_bwd' -c 16 \
        -o gpurun_out/train_kernels python scripts/prof_train_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402

ops = d2s.pkg.ops
dev = torch.device("cuda", 0)
B, N, D = 256, 196, 384
n = 22_774_432
p, g, m, v = torch.randn(n, device=dev), torch.randn(n, device=dev) * 1e-3, torch.zeros(n, device=dev), torch.zeros(n, device=dev)
sh = torch.empty(n, dtype=torch.bfloat16, device=dev)
lr, st = torch.full((1,), 5e-4, device=dev), torch.full((1,), 3.0, device=dev)
s, t = torch.randn(B, N + 1, D, device=dev), torch.randn(B, N + 1, D, device=dev)
h = torch.randn(B, N, D, device=dev).bfloat16().requires_grad_(True)
pol = (torch.rand(B, N, 1, device=dev) > 0.3).float().requires_grad_(True)
x = torch.randn(B, N + 1, D, device=dev).bfloat16().requires_grad_(True)
y = torch.randn(B, N + 1, D, device=dev).bfloat16().requires_grad_(True)
w = torch.ones(D, device=dev, requires_grad=True)
b = torch.zeros(D, device=dev, requires_grad=True)
for _ in range(3):
    ops.adamw_flat(p, g, m, v, sh, 0, n, lr, st, 0.9, 0.999, 1e-8, 0.05)
    ops.token_kl_rows(s[:, 1:], t[:, 1:])
    out = ops.pool_concat_train(h, pol)
    out.backward(torch.ones_like(out))
    ssum, hn = ops.add_layer_norm_train(x, y, w, b, 1e-6)
    (hn.float().sum() + ssum.float().sum()).backward()
    hp = ops.layer_norm(x, w, b, 1e-6, row0=1)
    hp.float().sum().backward()
torch.cuda.synchronize()
