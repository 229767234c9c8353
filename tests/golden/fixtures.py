"""Deterministic inputs/weights shared by make_goldens.py (which feeds them to the real reference)
and by the tests (which feed them to the oracle and to the CUDA path).  Everything is derived from
integer seeds with torch's CPU generator, so fixtures only need to store outputs plus checksums."""
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def gen(seed):
    return torch.Generator().manual_seed(int(seed))


def randn(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=gen(seed)) * scale


def seeded_state_dict(shapes, seed):
    """shapes: {param name: shape tuple}.  Non-degenerate values: non-zero biases, LayerNorm gains
    around 1, weights with unit-ish fan-in gain so predictor scores are well spread (no near-ties)."""
    g = gen(seed)
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros(shape, dtype=torch.long)
        elif name.endswith("running_var"):
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif name in ("cls_token", "pos_embed"):
            sd[name] = torch.randn(shape, generator=g) * 0.2
        elif len(shape) == 1 and name.endswith("weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = torch.randn(shape, generator=g) / (fan_in ** 0.5)
    return sd


def checksum(t):
    t = t.detach().double()
    return [float(t.sum()), float(t.abs().sum())]


def sd_checksum(sd):
    s = a = 0.0
    for k in sorted(sd):
        if sd[k].is_floating_point():
            c = checksum(sd[k])
            s += c[0]
            a += c[1]
    return [s, a]


def save_npz(name, arrays, meta):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name), **out)


def load_npz(name):
    z = np.load(os.path.join(HERE, name))
    meta = json.loads(bytes(z["__meta__"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "__meta__"}
    return arrays, meta


# small model used for the model-level goldens (N=196 is hard-coded by the reference: 224/16 squared)
SMALL_CFG = dict(embed_dim=128, depth=4, num_heads=2, patch_size=16, num_classes=16)
