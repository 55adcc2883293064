"""Worker of tests/test_gpu_round2.py::test_shared_warmup_bit_identical_on_one_and_two_gpus: the shared (one step size,
one metric) NUTS warm-up over C_total = chains_per_rank x world chains, sharded over `world` GPUs; rank 0 saves the
step-size history, the adapted metric and every chain's final state in global chain order.
usage: [torchrun ...] shared_warmup_worker.py out.npz chains_per_rank"""
import os
import sys

import numpy as np
import torch

import blackjax_b200 as bj

out, C = sys.argv[1], int(sys.argv[2])
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
D, T_ = 64, 120
tgt = bj.targets.DiagGaussian(np.logspace(-1, 1, D))
q0 = bj.random.normal(bj.random.split(bj.random.key(7, dev), C * world)[rank * C:(rank + 1) * C], (D,))
warm = bj.window_adaptation(bj.nuts, tgt, shared=True, max_num_doublings=6)
(state, params), hist = warm.run(bj.random.key(11, dev), q0, T_)
pos, lp = state.position, state.logdensity
if world > 1:
    allp = [torch.empty_like(pos) for _ in range(world)]
    alll = [torch.empty_like(lp) for _ in range(world)]
    dist.all_gather(allp, pos)
    dist.all_gather(alll, lp)
    pos, lp = torch.cat(allp), torch.cat(alll)
if rank == 0:
    np.savez(out, n_ranks=world, eps_history=np.asarray(hist), imm=params["inverse_mass_matrix"].cpu().numpy(),
             step_size=np.float32(params["step_size"]), position=pos.cpu().numpy(), logdensity=lp.cpu().numpy())
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
