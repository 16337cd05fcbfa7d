"""N > 1 host logic on CPU: slab arithmetic and the final gather (world_size 2 and 3, gloo)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import tape_interp
from epgpy_b200 import lowering, sharding
from util import product_namespace


def test_slabs_cover_the_grid():
    for natoms in (1, 2, 7, 1000, 1000003):
        for world in (1, 2, 3, 8):
            parts = sharding.slabs(natoms, world)
            assert parts[0][0] == 0 and sum(c for _, c in parts) == natoms
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        sharding.slab(10, 3, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, full, natoms, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, c = sharding.slab(natoms, rank, world)
    local = torch.from_numpy(np.ascontiguousarray(full[:, b:b + c]))
    got = sharding.gather_rows(local, natoms)
    out[rank] = bool(np.array_equal(got.numpy(), full))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_of_signal_slabs(world):
    """every rank simulates its atom slab (here: takes it from the tape interpreter's result) and the
    gathered array equals the single-process result, ragged slabs included (35 atoms over 2 / 3 ranks)"""
    epg = product_namespace()
    case = cases.fisp(epg, 12, sizes=(5, 7, 1))
    low = lowering.lower(case["seq"])
    sig, _ = tape_interp.run(low)  # [nadc, natoms, npool]
    assert low.natoms == 35
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), sig, low.natoms, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
