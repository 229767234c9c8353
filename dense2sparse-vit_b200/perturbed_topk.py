"""PerturbedTopK with the reference's call signature (vit_models/peturbed_topk.py:5-13), running on the
d2s kernels: no (b, nS, k, d) one-hot tensor, no host RNG round trip."""
import torch.nn as nn

from . import ops


class PerturbedTopK(nn.Module):
    """PerturbedTopK(k, num_samples=500, sigma=0.05)(x, current_sigma=0.05) -> indicators (b, k, d).

    The reference draws its noise with the host generator and copies it to the device every call
    (peturbed_topk.py:29).  Here:
      * `noise=` (b, num_samples, d): use exactly these standard-normal draws (bit-exact indicators vs the
        reference given the same draws);
      * otherwise the kernel draws its own noise (Philox) from `seed` -- same distribution, different
        stream, nothing materialised in HBM.
    """

    def __init__(self, k: int, num_samples: int = 500, sigma: float = 0.05):
        super().__init__()
        self.num_samples = num_samples
        self.sigma = sigma
        self.k = k

    def __call__(self, x, current_sigma=0.05, noise=None, seed=None):
        return ops.perturbed_topk(x, self.k, self.num_samples, current_sigma, noise=noise, seed=seed)


class PerturbedTopKFunction:
    """Name kept for callers that use PerturbedTopKFunction.apply(x, k, num_samples, sigma) directly."""

    @staticmethod
    def apply(x, k, num_samples=500, sigma=0.05, noise=None, seed=None):
        return ops.perturbed_topk(x, k, num_samples, sigma, noise=noise, seed=seed)
