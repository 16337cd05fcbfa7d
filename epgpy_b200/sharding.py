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


def run_gather(plan, device, group=None, nchunk=8, out=None):
    """the north-star multi-GPU step: every rank runs ITS slab of the plan's atoms (sharding.slab) and the signal
    slabs are all-gathered over NCCL / NVLink, so that every rank ends with the whole [nadc][natoms][npool] signal in
    its HBM.  The slab is cut in `nchunk` column chunks: the all-gather of chunk j (NCCL's own stream,
    async_op=True) overlaps with the kernel of chunk j + 1, and so does the scatter of chunk j to its columns of `out`
    (one strided device copy per rank and chunk, on a side stream).  Ragged slabs are padded to the largest one for the collective.
    Enqueues on torch's current stream and returns `out` without synchronising."""
    import torch
    import torch.distributed as dist

    low = plan.low
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = slabs(low.natoms, world)
    begin, count = parts[rank]
    cmax = max(c for _, c in parts)
    nchunk = max(1, min(int(nchunk), cmax))
    per = -(-cmax // nchunk)
    dev = torch.device("cuda", device)
    cdt = torch.complex128 if low.dtype == "f64" else torch.complex64
    if low.nvar and low.njac:
        raise NotImplementedError("run_gather gathers read-out rows only; gather Jacobian slabs with gather_rows")
    if out is None:
        out = torch.empty((low.nadc, low.natoms, low.npool), dtype=cdt, device=dev)
    flat = lambda t: torch.view_as_real(t).reshape(-1)
    cur = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)  # scatters chunk j to its columns of `out` while the kernel of chunk j + 1 runs
    for b in range(0, cmax, per):
        w = min(per, cmax - b)               # columns of this chunk (same on every rank)
        c = max(0, min(w, count - b))        # atoms this rank really has in it
        send = torch.empty((low.nadc, w, low.npool), dtype=cdt, device=dev)
        recv = torch.empty((world, low.nadc, w, low.npool), dtype=cdt, device=dev)
        if c < w:
            send.zero_()
        if c:
            plan.run_strided(device, begin + b, c, send, w)
        work = dist.all_gather_into_tensor(flat(recv), flat(send), group=group, async_op=True)
        with torch.cuda.stream(side):
            work.wait()
            recv.record_stream(side)
            for r, (rb, rc) in enumerate(parts):
                cr = max(0, min(w, rc - b))
                if cr:
                    out[:, rb + b:rb + b + cr] = recv[r, :, :cr]
    cur.wait_stream(side)
    out.record_stream(side)
    return out


def simulate(sequence, *, group=None, device=None, nchunk=4, dtype="float64", **kwargs):
    """distributed `simulate` (one process per GPU, torch.distributed initialised): every rank lowers the sequence,
    simulates its slab of the flattened grid and the slabs are all-gathered on the devices (run_gather).
    Returns the device tensor complex [nadc][natoms][npool] -- the whole dictionary in the HBM of every rank, ready
    for device-side matching -- together with the Lowered description (grid shape, ADC times)."""
    import torch

    from . import engine, lowering

    options = {k: kwargs.pop(k) for k in ("max_nstate", "kvalue") if k in kwargs}
    low = lowering.lower(sequence, options=options, dtype=engine.norm_dtype(dtype), **kwargs)
    plan = engine.Plan(low)
    dev = torch.cuda.current_device() if device is None else int(device)
    return run_gather(plan, dev, group=group, nchunk=nchunk), low
