"""LayerNorm training kernels (d2s_add_layernorm_fwd / _bwd) alone at the training step's shape (256 x 197 rows of 384),
graph-timed with rotating buffers; algorithmic bytes / time against the measured HBM peak.

    python scripts/bench_ln_train.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import d2s  # noqa: E402
import bench  # noqa: E402

lib = d2s.pkg._lib
pk = bench.peaks()
rows, D = 256 * 197, 384
bf = torch.bfloat16
st = lambda: torch.cuda.current_stream().cuda_stream   # noqa: E731
sets = []
for _ in range(4):
    x, y, dh, gs = (torch.randn(rows, D, device="cuda").to(bf) for _ in range(4))
    sets.append(dict(x=x, y=y, dh=dh, gs=gs, s=torch.empty_like(x), h=torch.empty_like(x), dx=torch.empty_like(x),
                     stats=torch.empty(rows, 2, device="cuda")))
w, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
dgb = torch.zeros(2, D, device="cuda")
for s in sets:
    lib.call("d2s_add_layernorm_fwd", s["x"].data_ptr(), s["y"].data_ptr(), 1, w.data_ptr(), b.data_ptr(), rows, D, 1e-6,
             s["s"].data_ptr(), s["h"].data_ptr(), 1, s["stats"].data_ptr(), st())
out = {}
ms = bench.time_graphed([lambda s=s: lib.call("d2s_add_layernorm_fwd", s["x"].data_ptr(), s["y"].data_ptr(), 1, w.data_ptr(), b.data_ptr(),
                                              rows, D, 1e-6, s["s"].data_ptr(), s["h"].data_ptr(), 1, s["stats"].data_ptr(), st())
                         for s in sets], torch)
by = rows * D * 2 * 4
out["add_layernorm_fwd"] = {"us": ms * 1e3, "gbs": by / ms / 1e6, "frac_hbm": by / ms / 1e6 / pk["hbm"]}
ms = bench.time_graphed([lambda s=s: lib.call("d2s_add_layernorm_bwd", s["dh"].data_ptr(), 1, s["s"].data_ptr(), 1, s["stats"].data_ptr(),
                                              w.data_ptr(), s["gs"].data_ptr(), rows, D, s["dx"].data_ptr(), dgb[0].data_ptr(),
                                              dgb[1].data_ptr(), st()) for s in sets], torch)
out["add_layernorm_bwd"] = {"us": ms * 1e3, "gbs": by / ms / 1e6, "frac_hbm": by / ms / 1e6 / pk["hbm"]}
ms = bench.time_graphed([lambda s=s: lib.call("d2s_layernorm_bwd", s["dh"].data_ptr(), 1, s["s"].data_ptr(), 1, s["stats"].data_ptr(),
                                              w.data_ptr(), rows, D, s["dx"].data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), st())
                         for s in sets], torch)
by3 = rows * D * 2 * 3
out["layernorm_bwd"] = {"us": ms * 1e3, "gbs": by3 / ms / 1e6, "frac_hbm": by3 / ms / 1e6 / pk["hbm"]}
print(json.dumps(out))
