"""Host-side plumbing for multi-GPU runs: one process per GPU, `torch.distributed` only carries the 128-byte NCCL
unique id to the other ranks; every data-path exchange happens inside libtopopt_b200.so (NCCL send/recv + allreduce)."""
from __future__ import annotations

import os

from ._lib import Context


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def broadcast_bytes(dist, payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcast a fixed-size byte string from `src` over an initialised torch.distributed group (nccl or gloo)."""
    import torch
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t = torch.tensor(list(payload), dtype=torch.uint8)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def create_distributed_context(dist, local_rank: int) -> Context:
    """Context on `cuda:local_rank` joined to an NCCL communicator spanning the process group."""
    import torch
    ctx = Context(local_rank)
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = Context.comm_unique_id() if rank == 0 else None
    dev = torch.device("cuda", local_rank) if dist.get_backend() == "nccl" else None
    uid = broadcast_bytes(dist, uid, 128, 0, dev)
    ctx.comm_init(world, rank, uid)
    return ctx
