"""Inference throughput of Variant B (the reference author's Dense2Sparse model: one stage at block 3, keep ratio 0.7,
large LayerNorm predictor, top-k mode) through the same CUDA-graph runner as bench.py.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
torch.manual_seed(0)
pkg = d2s.pkg
model = pkg.variant_b.VisionTransformerDiffPruning(pruning_loc=[3], token_ratio=[0.7], distill=True, topk_selection=True,
                                                   predictor_loss_type="kl_div", **bench.DEIT_S)
runner = pkg.runner.InferenceRunner(model, B, dev, dtype=torch.bfloat16, use_graph=True, warmup=2)
runner.static_in.copy_(torch.randn(runner.static_in.shape, device=dev).to(torch.bfloat16))
for _ in range(5):
    runner.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    runner.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"metric": "images/sec Dense2Sparse (Variant B) DeiT-S 1 stage@3 kr=0.7 @224", "value": B / ms * 1e3, "unit": "images/s",
                  "ms_per_step": ms, "batch": B, "dtype": "bf16", "cuda_graph": True,
                  "outputs": "logits + 12 CLS-attention rows + predictor logits + kept indices"}))
