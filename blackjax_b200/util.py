"""Driver loop mirroring ``blackjax.util.run_inference_algorithm`` (blackjax/util.py:150-213)."""
import torch

from . import random as bjx_random


def run_inference_algorithm(rng_key, inference_algorithm, num_steps, initial_state=None, initial_position=None,
                            transform=lambda state, info: (state, info), collect=True):
    """``keys = split(rng_key, num_steps)`` then ``num_steps`` calls of ``step`` (util.py:200-211).
    Each step key is split into one key per chain (step-major schedule).  ``transform`` picks what is kept;
    with ``collect=False`` only the final state is returned."""
    if (initial_state is None) == (initial_position is None):
        raise ValueError("Either `initial_state` or `initial_position` must be specified, but not both.")
    if initial_state is None:
        initial_state = inference_algorithm.init(initial_position)
    keys = bjx_random.split(rng_key, num_steps)
    state = initial_state
    history = []
    for t in range(num_steps):
        state, info = inference_algorithm.step(keys[t], state)
        if collect:
            history.append(transform(state, info))
    return state, history
