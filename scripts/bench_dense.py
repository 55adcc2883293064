"""BASELINE config 2: HMC, 1024-D correlated Gaussian, 65536 chains, dense mass matrix, 50 leapfrog steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np, torch
import blackjax_b200 as bj
from oracle import targets as ot
DEV = "cuda:0"
C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
L = int(sys.argv[3]) if len(sys.argv) > 3 else 50
K = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cov, prec = ot.correlated_gaussian(D, seed=0)
tgt = bj.targets.DenseGaussian(prec)
imm = torch.from_numpy(cov).to(DEV)
g_ = torch.Generator(device=DEV).manual_seed(0)
q0 = 0.1 * torch.randn(C, D, device=DEV, generator=g_)
st = bj.hmc.init(q0, tgt)
kern = bj.hmc.build_kernel(inplace=True)
keys = bj.random.split(bj.random.key(0, DEV), K + 1)
st, info = kern(keys[0], st, tgt, 0.5, imm, L)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(1, K + 1):
    st, info = kern(keys[t], st, tgt, 0.5, imm, L)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
flops = (2 * L + 3) * 2.0 * C * D * D
print(json.dumps({"config": "c2_hmc_dense_1024", "C": C, "D": D, "L": L, "ms_per_transition": ms,
                  "leapfrogs_per_s": C * L / (ms * 1e-3), "algorithmic_TFLOPs": flops / (ms * 1e-3) / 1e12,
                  "ms_per_leapfrog": ms / L, "acc": float(info.acceptance_rate.mean())}))
