"""N > 1 on real GPUs: one process per GPU over NCCL.  Every rank simulates its slab of the flattened grid and the
slabs are all-gathered (sharding.run_gather: chunked, overlapped with the slab kernels; sharding.gather_rows: one
padded collective) -- ragged slabs included.  Skipped on a one-GPU box; CPU / gloo counterpart: test_sharding_cpu.py."""

import os
import socket

import numpy as np
import pytest

import cases
import oracle_api
from util import RTOL64, product_namespace

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ref, out):
    import torch
    import torch.distributed as dist

    from epgpy_b200 import engine, lowering, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        epg = product_namespace()
        case = cases.fisp(epg, 40, sizes=(5, 7, 3))  # 105 atoms: ragged over 2 (53 + 52) and 4 ranks
        low = lowering.lower(case["seq"])
        plan = engine.Plan(low)
        res = {}
        for nchunk in (1, 3):
            full = sharding.run_gather(plan, rank, nchunk=nchunk)
            torch.cuda.synchronize()
            got = full.cpu().numpy().reshape((low.nadc,) + tuple(low.grid))
            res[f"run_gather{nchunk}"] = float(np.abs(got - ref).max() / np.abs(ref).max())
        # the same gather with copy-engine pushes into CUDA-IPC peer windows
        window = sharding.PeerWindow(low, rank)
        for nchunk in (1, 3):
            window.tensor.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            full = sharding.run_gather_p2p(plan, window, nchunk=nchunk)
            torch.cuda.synchronize()
            got = full.cpu().numpy().reshape((low.nadc,) + tuple(low.grid))
            res[f"p2p{nchunk}"] = float(np.abs(got - ref).max() / np.abs(ref).max())
        window.close()
        b, c = sharding.slab(low.natoms, rank, world)
        local, _ = plan.run(rank, b, c)
        full = sharding.gather_rows(local, low.natoms)
        got = full.cpu().numpy().reshape((low.nadc,) + tuple(low.grid))
        res["gather_rows"] = float(np.abs(got - ref).max() / np.abs(ref).max())
        # the public per-rank call: simulate(shard=...) returns the rank's slab on the host
        sl = epg.simulate(case["seq"], shard=(rank, world), device=rank)
        res["shard"] = float(np.abs(sl - ref.reshape(low.nadc, -1)[:, b:b + c]).max() / np.abs(ref).max())
        # the distributed entry point
        full, low2 = sharding.simulate(case["seq"], device=rank, nchunk=2)
        torch.cuda.synchronize()
        res["simulate"] = float(np.abs(full.cpu().numpy().reshape((low2.nadc,) + tuple(low2.grid)) - ref).max() / np.abs(ref).max())
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_gathered_over_nccl(world):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ref, _ = oracle_api.run(cases.fisp(oracle_api.epg, 40, sizes=(5, 7, 3)))
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ref, out), nprocs=world, join=True)
    for r in range(world):
        for key, err in out[r].items():
            assert err < RTOL64, (r, key, err)
