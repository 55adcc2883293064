"""MEADS warm-up: time per warm-up step (fold statistics + one GHMC transition of every chain + the shuffle every K steps).
usage: python scripts/bench_meads.py [C] [D] [K] [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import blackjax_b200 as bj

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
dev = "cuda:0"
tgt = bj.targets.DiagGaussian(np.logspace(-1, 1, D))
q0 = bj.random.normal(bj.random.split(bj.random.key(7, dev), C), (D,))
warm = bj.meads_adaptation(tgt, num_chains=C, num_folds=K)
warm.run(bj.random.key(0, dev), q0, num_steps=8)
torch.cuda.synchronize()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    (last, params), _ = warm.run(bj.random.key(1, dev), q0, num_steps=steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"MEADS C={C} D={D} K={K}: {ms / steps * 1e3:.1f} us per warm-up step ({steps} steps, {ms:.1f} ms); step_size "
          f"{float(params['step_size']):.4f} alpha {float(params['alpha']):.4f}")
