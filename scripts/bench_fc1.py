"""fc1 -> GELU forward of the training step (M = 256 x 197 rows, 384 -> 1536): the dual-output tcgen05 GEMM (GELU and the
pre-activation from one kernel) against cuBLAS + torch GELU, graph-timed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import d2s, bench
ops = d2s.pkg.ops
M, K, N = 256 * 197, 384, 1536
xs = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(4)]
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
b = torch.zeros(N, device="cuda").bfloat16()
out = {}
out["d2s_dual_us"] = 1e3 * bench.time_graphed([lambda x=x: ops.linear_act(x, w, b, ops.ACT_GELU, want_pre=True) for x in xs], torch, launches=8)
out["d2s_single_us"] = 1e3 * bench.time_graphed([lambda x=x: ops.linear_act(x, w, b, ops.ACT_GELU) for x in xs], torch, launches=8)
out["cublas_plus_gelu_us"] = 1e3 * bench.time_graphed([lambda x=x: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b)) for x in xs], torch, launches=8)
out["cublas_us"] = 1e3 * bench.time_graphed([lambda x=x: torch.nn.functional.linear(x, w, b) for x in xs], torch, launches=8)
print(json.dumps(out))
