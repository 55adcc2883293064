"""BASELINE config 0/1 shape (HMC, 100-D isotropic Gaussian, 1024 chains, diag mass, L=10): launch-bound, so run through the
native sampler loop (bjx_hmc_sample: no Python or host sync between transitions) and report microseconds per transition."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
import blackjax_b200 as bj
DEV = "cuda:0"
C, D, L, T_ = 1024, 100, 10, 2000
tgt = bj.targets.StdNormal(D)
imm = torch.ones(D, device=DEV)
st = bj.hmc.init(torch.randn(C, D, device=DEV), tgt)
key = bj.random.key(0, DEV)
bj.sample_hmc_native(key, st, tgt, 0.2, imm, L, 50, keep_history=False)
torch.cuda.synchronize()
res = {}
for name, kw in (("native", dict(keep_history=False)), ("native+history", dict(keep_history=True))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fin, hist, acc = bj.sample_hmc_native(key, st, tgt, 0.2, imm, L, T_, **kw)
    e1.record()
    torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) * 1e3 / T_
alg = bj.hmc(tgt, 0.2, imm, L)
keys = bj.random.split(key, 300)
s2 = st
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(300):
    s2, info = alg.step(keys[t], s2)
e1.record()
torch.cuda.synchronize()
res["python_loop"] = e0.elapsed_time(e1) * 1e3 / 300
print(json.dumps({"config": "c1_hmc_iso_1024x100_L10", "us_per_transition": res,
                  "leapfrogs_per_s_native": C * L / (res["native"] * 1e-6), "acc": float(acc.mean())}))
