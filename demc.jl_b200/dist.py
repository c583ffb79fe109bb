"""Multi-GPU host plumbing: one process per GPU (torchrun), chains sharded contiguously, archive
replicated.  Mirrors the *sync mode* of demcz_sample_par (src/demcz.jl:106-165 upstream): every K
generations all workers' new archive rows become visible to everyone — here by an in-place NCCL
all-gather inside libdemcz_cuda.so (csrc/demcz_comm.cu), not through a SharedArray.

torch.distributed is used only to agree on the NCCL unique id and for barriers / host-side
reductions; the data path never goes through torch.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

__all__ = ["shard", "env_rank_world", "broadcast_unique_id", "make_sharded_sampler", "open_peer_archives",
           "ensure_process_group", "pad_archive_for_sharding", "allgather_rows", "allreduce_sum", "barrier"]


def shard(N_global: int, rank: int, world: int) -> Tuple[int, int]:
    """(N_local, chain_offset) of the contiguous equal partition the in-place all-gather needs.
    The global chain id is the Philox counter, so results do not depend on `world`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if N_global % world != 0:
        raise ValueError(f"N={N_global} chains must divide evenly over {world} GPUs")
    n = N_global // world
    return n, n * rank


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def broadcast_unique_id(make_id, rank: int, device=None, group=None) -> bytes:
    """Rank 0 calls `make_id()` (-> 128 bytes, demcz_comm_unique_id); everyone gets the same bytes."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def make_sharded_sampler(logobj, d, *, N_global, rank=None, world=None, device=None, **kw):
    """Sampler for this rank's shard of `N_global` chains, with the NCCL communicator wired when
    world > 1 (torch.distributed must be initialised).  All other keywords as api.Sampler."""
    from . import _lib
    from .api import Sampler
    r, w, lr = env_rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    device = lr if device is None else device
    n_local, off = shard(N_global, rank, world)
    s = Sampler(logobj, d, N=n_local, rank=rank, world=world, N_global=N_global, chain_offset=off, device=device, **kw)
    if world > 1:
        import torch

        def make_id():
            buf = (C.c_uint8 * 128)()
            rc = s.lib.demcz_comm_unique_id(buf)
            if rc != _lib.OK:
                raise _lib.DemczError(rc, s.lib.demcz_last_error(None).decode())
            return bytes(buf)

        uid = broadcast_unique_id(make_id, rank, device=torch.device("cuda", device))
        s.comm_init(uid)
    return s


def open_peer_archives(s, rank=None, world=None, device=None, group=None):
    """After `s.set_archive(...)` on every rank: all-gather the 128-byte peer records and open the peers' archive
    replicas, so that appended rows are pushed straight into them over NVLink by the library's own kernel
    (instead of ncclAllGather).  No-op for world == 1."""
    import torch
    import torch.distributed as dist
    r, w, lr = env_rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    if world == 1:
        return
    dev = torch.device("cuda", lr if device is None else device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(s.peer_info()), dtype=torch.uint8).to(dev)
    allinfo = [torch.zeros(128, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allinfo, mine, group=group)
    s.open_peers(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allinfo))


def ensure_process_group(local_rank=0):
    """Initialise torch.distributed from the torchrun environment if the caller has not (NCCL when CUDA is there)."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")


def pad_archive_for_sharding(Zmat, N_global, Ngeneration, K):
    """A rank's appended segment must not straddle the end of the archive ring (demcz_comm.cu): make M0 and the capacity
    multiples of N by repeating the oldest rows in front (they stay ordinary archive members, drawn like any other row).
    Returns (Zmat_padded, capacity)."""
    M0 = Zmat.shape[0]
    pad = (-M0) % N_global
    if pad:
        Zmat = np.asfortranarray(np.concatenate([Zmat[:pad], Zmat], axis=0))
    cap = Zmat.shape[0] + N_global * (-(-Ngeneration // K))
    return Zmat, cap


def _dev(local_rank):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", local_rank) if dist.get_backend() == "nccl" else torch.device("cpu")


def allgather_rows(a, local_rank=0):
    """Concatenate every rank's array along axis 0 (rank order = global chain order); equal shapes on all ranks."""
    import torch
    import torch.distributed as dist
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a).to(_dev(local_rank))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy() for o in out], axis=0)


def allreduce_sum(x, local_rank=0):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev(local_rank))
    dist.all_reduce(t)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
