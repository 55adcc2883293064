"""Restatement of blackjax.diagnostics.potential_scale_reduction (blackjax/diagnostics.py:39-89).  TEST INFRASTRUCTURE."""
import numpy as np


def potential_scale_reduction(x, chain_axis=0, sample_axis=1):
    x = np.asarray(x, np.float64)
    assert x.shape[chain_axis] > 1
    n = x.shape[sample_axis]
    per_chain_mean = x.mean(axis=sample_axis, keepdims=True)
    per_chain_var = x.var(axis=sample_axis, ddof=1, keepdims=True)
    between = n * per_chain_mean.var(axis=chain_axis, ddof=1, keepdims=True)
    within = per_chain_var.mean(axis=chain_axis, keepdims=True)
    estimator = (n - 1) / n * within + between / n
    return np.sqrt(estimator / within).squeeze()
