"""D2H scaling microbenchmark (one process per GPU, torchrun): every rank copies SIZE bytes from its GPU into its own
pinned host buffer, all ranks at once; reports per-rank and aggregate GB/s, the time to pin the buffer, and the same
for a pitched (2-D) copy like the one `Plan.run_to_host` issues.  Explains the end-to-end scaling of bench.py: the
per-rank result slabs all land in ONE host's memory.

    python -m torch.distributed.run --nproc-per-node N profiles/scripts/d2h_scaling.py [--gb 2]
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--gb", type=float, default=2.0)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(args.gb * (1 << 30))
dev = torch.empty(n, dtype=torch.uint8, device="cuda").fill_(1)
t0 = time.perf_counter()
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
t_pin = time.perf_counter() - t0


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


res = {}
for name in ("contiguous",):
    best = 0.0
    for _ in range(args.reps):
        sync()
        t0 = time.perf_counter()
        if name == "contiguous":
            host.copy_(dev, non_blocking=True)
        else:  # 1000 rows, each a column range of a wider buffer: what the chunked result copy looks like
            rows = 1000
            w = n // rows // 2
            host[: rows * 2 * w].view(rows, 2 * w)[:, :w].copy_(dev[: rows * 2 * w].view(rows, 2 * w)[:, :w], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nbytes = n if name == "contiguous" else rows * w
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = max(best, nbytes / float(t.item()) / 1e9)
    res[name] = {"per_rank_gbs": best, "aggregate_gbs": best * world}
if rank == 0:
    print(json.dumps({"n_gpus": world, "gb_per_rank": args.gb, "pin_seconds_per_gb": t_pin / args.gb, **res}))
if world > 1:
    dist.destroy_process_group()
