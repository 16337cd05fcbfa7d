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
    # send / receive buffers of every chunk and the side stream are kept with the plan: allocating 2 x 16 GB per step
    # through the caching allocator (whose blocks, once recorded on the side stream, come back late) stalled the step
    key = ("gather", device, world, nchunk)
    cache = getattr(plan, "_gather_cache", None)
    if cache is None or cache[0] != key:
        bufs = []
        for b in range(0, cmax, per):
            w = min(per, cmax - b)
            bufs.append((torch.zeros((low.nadc, w, low.npool), dtype=cdt, device=dev),
                         torch.empty((world, low.nadc, w, low.npool), dtype=cdt, device=dev)))
        cache = plan._gather_cache = (key, bufs, torch.cuda.Stream(dev))
    _, bufs, side = cache  # side: scatters chunk j to its columns of `out` while the kernel of chunk j + 1 runs
    side.wait_stream(cur)
    for (send, recv), b in zip(bufs, range(0, cmax, per)):
        w = min(per, cmax - b)               # columns of this chunk (same on every rank)
        c = max(0, min(w, count - b))        # atoms this rank really has in it (the padding columns stay zero)
        if c:
            plan.run_strided(device, begin + b, c, send, w)
        work = dist.all_gather_into_tensor(flat(recv), flat(send), group=group, async_op=True)
        with torch.cuda.stream(side):
            work.wait()
            for r, (rb, rc) in enumerate(parts):
                cr = max(0, min(w, rc - b))
                if cr:
                    out[:, rb + b:rb + b + cr] = recv[r, :, :cr]
    cur.wait_stream(side)
    return out


class _DeviceArray:
    """a device buffer owned by libepgx (epgx_peer_alloc) seen by torch through the CUDA array interface"""

    def __init__(self, ptr, shape, complex_dtype):
        typestr = "<c16" if complex_dtype == "complex128" else "<c8"
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerWindow:
    """the gathered signal [nadc][natoms][npool] of one rank, mapped into every other rank of the node (CUDA IPC over
    NVLink; include/epgx.h `epgx_peer_*`).  Collective: every rank of `group` constructs it with the same arguments."""

    def __init__(self, low, device, group=None):
        import ctypes

        import torch
        import torch.distributed as dist

        from . import engine

        self.L = engine.lib()
        self.device, self.group = device, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cdt = torch.complex128 if low.dtype == "f64" else torch.complex64
        self.csz = 16 if low.dtype == "f64" else 8
        self.shape = (low.nadc, low.natoms, low.npool)
        nbytes = max(256, self.csz * low.nadc * low.natoms * low.npool)
        self.ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(device):
            engine._check(self.L.epgx_peer_alloc(nbytes, ctypes.byref(self.ptr), handle))
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(f"cuda:{device}")
            every = torch.empty((self.world, 64), dtype=torch.uint8, device=f"cuda:{device}")
            dist.all_gather_into_tensor(every.reshape(-1), mine, group=group)
            every = every.cpu().numpy()
            self.peers = []
            for r in range(self.world):
                if r == self.rank:
                    self.peers.append(self.ptr.value)
                    continue
                p = ctypes.c_void_p()
                engine._check(self.L.epgx_peer_open(every[r].tobytes(), ctypes.byref(p)))
                self.peers.append(p.value)
            self.tensor = torch.as_tensor(_DeviceArray(self.ptr.value, self.shape, "complex128" if low.dtype == "f64" else "complex64"),
                                          device=f"cuda:{device}")
            self.streams = [torch.cuda.Stream(device) for _ in range(min(4, max(1, self.world - 1)))]
        dist.barrier(group)

    def close(self):
        import torch.distributed as dist

        if getattr(self, "ptr", None) is None:
            return
        dist.barrier(self.group)  # nobody still pushes into a window that is about to go
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.L.epgx_peer_close(p)
        self.tensor = None
        self.L.epgx_peer_free(self.ptr)
        self.ptr = None


def run_gather_p2p(plan, window, nchunk=8):
    """the north-star multi-GPU step without a collective kernel: every rank runs its slab in `nchunk` column chunks,
    writing its own window in place, and pushes each finished chunk into the windows of all peers with the copy engines
    (pitched device -> device copies over NVLink, at the chunk's final columns) while the next chunk is computed.  Ends
    with a barrier: when it returns (after a stream synchronisation) every rank holds the whole [nadc][natoms][npool]
    signal in `window.tensor`."""
    import torch
    import torch.distributed as dist

    from . import engine

    low = plan.low
    world, rank, device = window.world, window.rank, window.device
    parts = slabs(low.natoms, world)
    begin, count = parts[rank]
    per = -(-max(1, count) // max(1, int(nchunk)))
    csz, L = window.csz * low.npool, window.L
    pitch = low.natoms * csz
    cur = torch.cuda.current_stream(device)
    with torch.cuda.device(device):
        for i, b in enumerate(range(0, count, per)):
            c = min(per, count - b)
            col = (begin + b) * csz
            plan.run_strided(device, begin + b, c, _Ptr(window.ptr.value + col), low.natoms)
            done = torch.cuda.Event()
            done.record(cur)
            # peers in ROTATED order (rank + 1, rank + 2, ...): at any moment every window receives from one rank instead
            # of all ranks writing into window 0 first, then window 1, ... (the schedule of an all-to-all)
            for k, d in enumerate(range(1, world)):
                r = (rank + d) % world
                st = window.streams[k % len(window.streams)]
                st.wait_event(done)
                engine._check(L.epgx_copy2d_device(window.peers[r] + col, pitch, window.ptr.value + col, pitch, c * csz, low.nadc,
                                                   st.cuda_stream))
        for st in window.streams:
            cur.wait_stream(st)
    # every rank has pushed its slab (its streams are ordered before `cur`): a device-side barrier closes the step
    token = torch.zeros(1, device=f"cuda:{device}")
    dist.all_reduce(token, group=window.group)
    return window.tensor


class _Ptr:
    """raw device address with the `data_ptr()` face of a tensor (engine.Plan.run_strided)"""

    def __init__(self, ptr):
        self._p = ptr

    def data_ptr(self):
        return self._p


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
