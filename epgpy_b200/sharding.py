"""Atom-slab sharding over ranks (one process per GPU, torch.distributed).

Atoms are independent for every operator of the path (epgpy/functions.py:173-192 acts per grid point;
X couples only the pools of one atom), so the flattened grid is cut in contiguous slabs and nothing is
exchanged while the sequence runs.  The only collective is the optional final gather of the signal
slabs (NCCL over NVLink on the GPU box; gloo in the CPU tests)."""

import numpy as np


def slab(natoms, rank, world):
    """contiguous atom range (begin, count) of `rank`: the first natoms % world ranks hold one atom more"""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(int(natoms), int(world))
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def slabs(natoms, world):
    return [slab(natoms, r, world) for r in range(world)]


def gather_rows(local, natoms, group=None):
    """all-gather per-rank slabs `local[rows, count_r, npool]` into `[rows, natoms, npool]` on every rank.
    Ragged slabs are padded to the largest one for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = slabs(natoms, world)
    cmax = max(c for _, c in parts)
    rows, count, npool = local.shape
    if count != parts[rank][1]:
        raise ValueError(f"rank {rank} holds {count} atoms, expected {parts[rank][1]}")
    send = local if count == cmax else torch.cat(
        [local, local.new_zeros((rows, cmax - count, npool))], dim=1)
    recv = torch.empty((world, rows, cmax, npool), dtype=local.dtype, device=local.device)
    # the collective runs on flat real views (complex dtypes are not uniformly supported by the backends)
    flat = (lambda t: torch.view_as_real(t).reshape(-1)) if local.is_complex() else (lambda t: t.reshape(-1))
    dist.all_gather_into_tensor(flat(recv), flat(send.contiguous()), group=group)
    out = torch.empty((rows, natoms, npool), dtype=local.dtype, device=local.device)
    for r, (b, c) in enumerate(parts):
        out[:, b:b + c] = recv[r, :, :c]
    return out
