"""d2s_linear_residual_ln_bf16 at DeiT-B widths (N = 768; K = 768: attn.proj, K = 3072: mlp.fc2) for ncu:
    ncu --set full --clock-control none -k regex:gemm_pair -s 4 -c 4 -o gpurun_out/gemm768 python scripts/prof_gemm768.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402

ops = d2s.pkg.ops
dev = torch.device("cuda", 0)
M, N = 512 * 197, 768
x = torch.randn(M, N, device=dev).bfloat16()
g, b = torch.ones(N, device=dev).bfloat16(), torch.zeros(N, device=dev).bfloat16()
for K in (768, 3072):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.zeros(N, device=dev).bfloat16()
    for _ in range(4):
        ops.linear_residual_ln(a, w, bias, x, g, b, 1e-6)
torch.cuda.synchronize()
