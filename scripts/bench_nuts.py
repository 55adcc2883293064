"""NUTS throughput on BASELINE config 3 (Neal's funnel D=128, 65536 chains, diag mass, depth<=10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json, sys, time
import numpy as np, torch
import blackjax_b200 as bj
from blackjax_b200._engine import get_engine
DEV = "cuda:0"
C = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
T_ = int(sys.argv[3]) if len(sys.argv) > 3 else 8
eps = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
tgt = bj.targets.Funnel(D)
imm = torch.ones(D, device=DEV)
q = 0.1 * bj.random.normal(bj.random.split(bj.random.key(0, DEV), C), (D,))
st = bj.nuts.init(q, tgt)
kern = bj.nuts.build_kernel(inplace=True)
keys = bj.random.split(bj.random.key(1, DEV), T_ + 2)
for t in range(2):
    st, info = kern(keys[t], st, tgt, eps, imm, 10)
torch.cuda.synchronize()
eng = get_engine(st.position, tgt)
res = []
for t in range(2, T_ + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    st, info = kern(keys[t], st, tgt, eps, imm, 10)
    e1.record()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    n = info.num_integration_steps
    leaves, depth = eng.nuts_last_stats()
    res.append(dict(ms=e0.elapsed_time(e1), wall_ms=(w1 - w0) * 1e3, leapfrogs=int(n.sum()), mean_tree=float(n.float().mean()),
                    max_tree=int(n.max()), leaf_launches=leaves, depth=depth, div=int(info.is_divergent.sum()),
                    acc=float(info.acceptance_rate.mean())))
    print(json.dumps(res[-1]))
tot_ms = sum(r["ms"] for r in res)
tot_lf = sum(r["leapfrogs"] for r in res)
print(json.dumps({"C": C, "D": D, "transitions": T_, "leapfrogs_per_s": tot_lf / (tot_ms * 1e-3),
                                    "ms_per_transition": tot_ms / T_, "GBps_at_44D": 44.0 * D * tot_lf / (tot_ms * 1e-3) / 1e9}))
