"""Time the d2s kernels alone at the bench step's shapes (CUDA events) and print one line per kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda:0")
rows, roof = bench.kernel_breakdown(d2s.pkg.ops, B, dev, torch, bench.peaks())
for k in rows:
    print(f"{k['kernel']:36s} {k['shape']:30s} {k['ms'] * 1e3:8.1f} us {k['gbs']:8.0f} GB/s" +
          (f" {k['tflops']:7.1f} TF/s" if 'tflops' in k else ""))
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in roof.items() if k != "tensor"})
