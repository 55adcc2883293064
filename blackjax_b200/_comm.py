"""NCCL communicator for the one collective of the path (``bjx_allgather_stats``): created inside libbjx
(``bjx_nccl_comm_init_rank``) from a unique id that rank 0 broadcasts over the caller's ``torch.distributed`` process
group.  torch.distributed is plumbing here (rendezvous); the all-gather itself runs in libbjx on the engine's stream."""
import ctypes as C

import torch

from ._lib import check, lib

_COMMS = {}


def world(process_group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(process_group), dist.get_rank(process_group)


def nccl_comm(device, process_group=None):
    """(ncclComm_t as a ctypes void pointer or None, n_ranks, rank) for ``process_group`` on ``device``."""
    import torch.distributed as dist
    n, rank = world(process_group)
    if n == 1:
        return None, 1, 0
    device = torch.device(device)
    key = (id(process_group), device.index)
    if key not in _COMMS:
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            check(lib().bjx_nccl_unique_id(raw))
            buf = torch.tensor(list(raw), dtype=torch.uint8)
        if dist.get_backend(process_group) == "nccl":
            buf = buf.to(device)
        src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
        dist.broadcast(buf, src=src, group=process_group)
        raw = (C.c_ubyte * 128)(*buf.cpu().tolist())
        comm = C.c_void_p()
        check(lib().bjx_nccl_comm_init_rank(raw, n, rank, device.index, C.byref(comm)))
        _COMMS[key] = comm
    return _COMMS[key], n, rank
