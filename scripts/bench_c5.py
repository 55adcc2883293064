"""BASELINE config 5: HMC, hierarchical logistic regression (synthetic, 10000 params), chains sharded over GPUs
(131072 per GPU in the config), 20 leapfrog steps, diag mass.  Start in the typical set and eps = 0.005: the origin start
with eps = 0.02 pencilled in by SURVEY 8d is unstable for this centred model (acceptance 0).
    python scripts/bench_c5.py [chains_per_gpu] [transitions]      (torchrun for N > 1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np, torch
import blackjax_b200 as bj
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T_ = int(sys.argv[2]) if len(sys.argv) > 2 else 3
D, L, eps = 10000, 20, 0.005
x, bits = bj.targets.HierLogit.synthetic_data(D - 4, seed=1)
tgt = bj.targets.HierLogit(x, bits)
imm = torch.ones(D, device=dev)
g_ = torch.Generator(device=dev).manual_seed(rank)
q0 = torch.empty(C, D, device=dev)
q0[:, 0], q0[:, 1], q0[:, 2], q0[:, 3] = 0.5, float(np.log(0.7)), 1.0, -0.5
q0[:, 4:] = 0.5 + 0.7 * torch.randn(C, D - 4, device=dev, generator=g_)
st = bj.hmc.init(q0, tgt)
kern = bj.hmc.build_kernel(inplace=True)
keys = bj.random.split(bj.random.key(0, dev), T_ + 1)
ck = lambda t: bj.random.split(keys[t], C * world)[rank * C:(rank + 1) * C]
st, info = kern(ck(0), st, tgt, eps, imm, L)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(1, T_ + 1):
    st, info = kern(ck(t), st, tgt, eps, imm, L)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / T_], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    m = float(ms)
    print(json.dumps({"config": "c5_hmc_hier_logit_10000d", "n_gpus": world, "chains_per_gpu": C, "L": L,
                      "ms_per_transition": m, "leapfrogs_per_s": world * C * L / (m * 1e-3),
                      "sigmoid_evals_per_s": world * C * L * 8.0 * (D - 4) / (m * 1e-3),
                      "GBps_at_24D": 24.0 * D * world * C * L / (m * 1e-3) / 1e9, "acc": float(info.acceptance_rate.mean())}))
if world > 1:
    dist.destroy_process_group()
