"""BASELINE config 4: NUTS + window adaptation (step-size dual averaging + diagonal mass matrix), 512-D
ill-conditioned Gaussian (std = logspace(-1, 1, 512), tests/fixtures.py:74-78), chains sharded over the GPUs
of one node (32768 per GPU), shared (eps, M^-1): ONE NCCL all-gather of summary statistics per warm-up step.

    python scripts/bench_c4.py [chains_per_gpu] [warmup_steps]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4.py ...
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json, time
import numpy as np, torch
import blackjax_b200 as bj

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T_ = int(sys.argv[2]) if len(sys.argv) > 2 else 200
D = 512
scale = np.logspace(-1, 1, D)
tgt = bj.targets.DiagGaussian(scale)
# global chain c starts at normal(fold_in(key, c)) * 1 (over-dispersed relative to the small scales)
init_key = bj.random.key(7, dev)
q0 = bj.random.normal(bj.random.split(init_key, C * world)[rank * C:(rank + 1) * C], (D,))
warm = bj.window_adaptation(bj.nuts, tgt, shared=True, max_num_doublings=10)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
(state, params), hist = warm.run(bj.random.key(11, dev), q0, T_)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
imm = params["inverse_mass_matrix"].cpu().numpy()
rel = np.abs(imm / scale ** 2 - 1)
# a few sampling transitions with the adapted kernel, counting leapfrogs
nuts = bj.nuts(tgt, params["step_size"], params["inverse_mass_matrix"])
keys = bj.random.split(bj.random.key(13, dev), 5)
lf = 0
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = state
f0.record()
for k in keys:
    ck = bj.random.split(k, C * world)[rank * C:(rank + 1) * C]
    st, info = nuts.step(ck, st)
    lf += int(info.num_integration_steps.sum())
f1.record()
torch.cuda.synchronize()
tot = torch.tensor([float(lf), f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
if world > 1:
    lf_all = tot[:1].clone(); dist.all_reduce(lf_all); t_all = tot[1:].clone(); dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    tot = torch.cat([lf_all, t_all])
if rank == 0:
    print(json.dumps({"config": "c4_nuts_window_adaptation_512d", "n_gpus": world, "chains_per_gpu": C, "warmup_steps": T_,
                      "warmup_ms": float(ms), "ms_per_warmup_step": float(ms) / T_, "step_size": float(params["step_size"]),
                      "imm_rel_err_median": float(np.median(rel)), "imm_rel_err_max": float(rel.max()),
                      "eps_first5": [float(h) for h in hist[:5]],
                      "sampling_leapfrogs_per_s": float(tot[0]) / (float(tot[1]) * 1e-3),
                      "sampling_mean_tree": float(tot[0]) / (5 * C * world)}))
if world > 1:
    dist.destroy_process_group()
