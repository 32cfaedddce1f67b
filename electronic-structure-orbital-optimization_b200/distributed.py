"""Multi-GPU plumbing: one process per GPU, the ERI tensor sharded by its first index.

Every rank holds rows [t0, t0+mloc) of g, computes its rows of dE/dU and a partial energy, and the
only communication is one sum all-reduce of M*N+1 doubles per evaluation, issued by liboo_b200 on
its own NCCL communicator (bootstrapped here through torch.distributed)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(M: int, rank: int, world: int) -> Tuple[int, int]:
    """(t0, mloc): contiguous ceil-div slabs of the first index; trailing ranks may get fewer rows.
    Every rank must own at least one row."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    if world > M:
        raise ValueError(f"cannot shard M={M} rows over {world} ranks")
    base, extra = divmod(M, world)
    t0 = rank * base + min(rank, extra)
    return t0, base + (1 if rank < extra else 0)


def pair_selected(t: int, q: int) -> bool:
    """Which slab of the pair {(t,q),(q,t)} the pair-symmetric mode streams (mirror of
    oo::pair_selected in csrc/oo_k2.cuh): checkerboard, so every row keeps about M/2 slabs."""
    if t == q:
        return True
    return (t < q) if ((t + q) % 2 == 0) else (t > q)


def pair_slab_list(M: int, t0: int = 0, mloc: int = None):
    """[(t, q), ...] of the slabs g[t, q, :, :] that pair-packed storage keeps for rows
    [t0, t0+mloc), in storage (= streaming) order: t ascending, then q ascending (mirror of
    oo_pair_slab_list / pair_row_count / pair_ith_q in csrc/oo_k2.cuh)."""
    mloc = M - t0 if mloc is None else mloc
    if t0 < 0 or mloc < 1 or t0 + mloc > M:
        raise ValueError(f"bad shard rows [{t0}, {t0 + mloc}) of {M}")
    return [(t, q) for t in range(t0, t0 + mloc) for q in range(M) if pair_selected(t, q)]


def attach_nccl(engine, group=None) -> None:
    """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed
    (any backend) broadcasts its 128 bytes, every rank joins."""
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    backend = dist.get_backend(group)
    dev = engine.device if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = type(engine).nccl_unique_id()
        buf.copy_(torch.tensor(list(uid), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    engine.attach_comm(bytes(buf.cpu().tolist()), rank, world)


def attach_peer_memory(engine, group=None) -> None:
    """Enable the all-reduce fused into the evaluation's last kernel: every rank exports the CUDA
    IPC handle of its exchange buffer, torch.distributed all-gathers the 64-byte handles, every
    rank maps its peers' buffers over NVLink.  All ranks must live on one node."""
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    backend = dist.get_backend(group)
    dev = engine.device if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(engine.peer_export()), dtype=torch.uint8, device=dev)
    gathered = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    handles = b"".join(bytes(t.cpu().tolist()) for t in gathered)
    engine.peer_attach(handles, rank, world)
    dist.barrier(group)
