"""NUTS on the tensor-core dense path (config-2-like model: correlated Gaussian, dense mass matrix): time per transition.
usage: python scripts/bench_dense_nuts.py [C] [D] [depth] [eps] [transitions]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import blackjax_b200 as bj

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 8
eps = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
T = int(sys.argv[5]) if len(sys.argv) > 5 else 6
dev = "cuda:0"
rs = np.random.default_rng(0)
Q, _ = np.linalg.qr(rs.standard_normal((D, D)))
var = np.logspace(-1, 1, D)
cov = (Q * var) @ Q.T
cov = 0.5 * (cov + cov.T)
prec = (Q / var) @ Q.T
prec = 0.5 * (prec + prec.T)
tgt = bj.targets.DenseGaussian(prec.astype(np.float32))
imm = torch.from_numpy(cov.astype(np.float32)).to(dev)
q0 = 0.1 * bj.random.normal(bj.random.split(bj.random.key(7, dev), C), (D,))
kern = bj.nuts.build_kernel(inplace=True, max_tree_depth=depth)
state = bj.nuts.init(q0.clone(), tgt)
keys = bj.random.split(bj.random.key(0, dev), T + 2)
for t in range(2):
    state, info = kern(keys[t], state, tgt, eps, imm, depth)
torch.cuda.synchronize()
lf = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tot = torch.zeros((), dtype=torch.int64, device=dev)
for t in range(2, T + 2):
    state, info = kern(keys[t], state, tgt, eps, imm, depth)
    tot += info.num_integration_steps.sum()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
lf = int(tot)
print(f"dense NUTS C={C} D={D} depth<={depth} eps={eps}: {ms / T:.2f} ms per transition, {lf / ms * 1e3:.3e} leapfrogs/s, "
      f"mean tree {lf / (C * T):.1f}, depth histogram {torch.bincount(info.num_trajectory_expansions).tolist()}, "
      f"accept {float(info.acceptance_rate.mean()):.3f}")
