"""Metric objects accepted wherever the kernels take ``inverse_mass_matrix`` (blackjax/mcmc/metrics.py).

A 1-D tensor is the diagonal Euclidean metric, a 2-D ``[D, D]`` tensor the dense one (``default_metric`` /
``gaussian_euclidean``, metrics.py:180-346, 701-729); :func:`gaussian_euclidean_low_rank` builds the low-rank-modified
metric of metrics.py:349-467, whose momentum draw, kinetic energy, velocity and U-turn test all cost O(D k) inside the warp
kernels (libbjx ``bjx_set_metric_low_rank``)."""
from typing import NamedTuple

import torch


class LowRankMetric(NamedTuple):
    """M^-1 = diag(sigma) (I + U (diag(lam) - I) U^T) diag(sigma): sigma [D] > 0, U [D, k] with orthonormal columns, lam [k] > 0."""
    sigma: torch.Tensor
    U: torch.Tensor
    lam: torch.Tensor


def gaussian_euclidean_low_rank(sigma, U, lam) -> LowRankMetric:
    """blackjax/mcmc/metrics.py:349-467."""
    sigma = torch.as_tensor(sigma, dtype=torch.float32)
    U = torch.as_tensor(U, dtype=torch.float32)
    lam = torch.as_tensor(lam, dtype=torch.float32)
    if sigma.ndim != 1 or U.ndim != 2 or lam.ndim != 1 or U.shape[0] != sigma.shape[0] or U.shape[1] != lam.shape[0]:
        raise ValueError("gaussian_euclidean_low_rank expects sigma [D], U [D, k], lam [k]")
    return LowRankMetric(sigma, U, lam)
