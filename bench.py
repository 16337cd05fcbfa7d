#!/usr/bin/env python
"""bench.py -- MRF FISP dictionary generation on B200 (BASELINE.json configs[2]).

One "step" = one pass of the hot path over the whole per-GPU batch: the 1000-TR FISP sequence
(`[T(180,0), E(20)] + [T(FA_i*B1, 90), E(3), ADC, E(TR_i-3), S(1)] x 1000`) for
100 x 100 x 100 = 1 M (T1, T2, B1) atoms per GPU, unbounded number of states, FP64.  With N > 1 GPUs the
B1 axis grows to 100*N values and the flattened grid is cut in N contiguous slabs (weak scaling, no
collective on the data path; `--gather` adds the final NCCL all-gather of the signal slabs).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference ...                      # CPU reference arm (oracle port, host cores)

Prints ONE JSON line (see the field notes in DESIGN.md section "Measurement").
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

NTR = 1000
GRID = (100, 100, 100)


def fisp_schedule(ntr, seed=0):
    rng = np.random.RandomState(seed)
    fa = 10 + 50 * np.abs(np.sin(np.arange(ntr) * np.pi / 200))
    tr = rng.uniform(11, 16, ntr)
    return fa, tr


def fisp_sequence(epg, T1, T2, B1, ntr=NTR, jac=False):
    """SURVEY.md 8d M3 (the reference-API form of BASELINE configs[2]); axes: T1 -> 0, T2 -> 1, B1 -> 2"""
    fa, tr = fisp_schedule(ntr)
    T1 = np.asarray(T1, dtype=float)
    T2 = np.asarray(T2, dtype=float)[None, :]
    B1 = np.asarray(B1, dtype=float)[None, None, :]
    o1 = {"order1": ["T1", "T2"]} if jac else {}
    seq = [epg.T(180, 0), epg.E(20, T1, T2, **o1)]
    for i in range(ntr):
        tk = {"order1": {"B1": {"alpha": fa[i]}}} if jac else {}
        seq.append([epg.T(fa[i] * B1, 90, **tk), epg.E(3, T1, T2, **o1), epg.ADC, epg.E(tr[i] - 3, T1, T2, **o1), epg.S(1)])
    return seq


def grid_axes(grid, world=1):
    T1 = np.linspace(300, 3000, grid[0])
    T2 = np.linspace(20, 300, grid[1])
    B1 = np.linspace(0.7, 1.2, grid[2] * world)
    return T1, T2, B1


# ------------------------------------------------------------------------------------------------ #
# CPU reference arm: the oracle port (numpy restatement of the reference's algorithm, full storage)
# ------------------------------------------------------------------------------------------------ #


def _cpu_worker(args):
    ntr, max_nstate, t1, t2, b1 = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import oracle_api

    seq = fisp_sequence(oracle_api.epg, t1, t2, b1, ntr)
    t0 = time.perf_counter()
    out = oracle_api.O.simulate(seq, max_nstate=max_nstate)
    return time.perf_counter() - t0, int(np.prod(out.shape[1:]))


def cpu_reference(ntr, max_nstate, cores, atoms_per_core, grid=GRID):
    """time the oracle on `cores` processes, each simulating its own slab of `atoms_per_core` atoms taken
    from the bench grid; returns (atoms/s over the wall clock of the slowest worker, atoms, seconds)"""
    import multiprocessing as mp

    T1, T2, B1 = grid_axes(grid)
    side = max(1, round(atoms_per_core ** (1 / 3)))
    jobs = []
    rng = np.random.RandomState(1)
    for c in range(cores):
        i, j, k = (rng.randint(0, max(1, n - side)) for n in grid)
        jobs.append((ntr, max_nstate, T1[i:i + side], T2[j:j + side], B1[k:k + side]))
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    atoms = sum(r[1] for r in res)
    return atoms / wall, atoms, wall, side


# ------------------------------------------------------------------------------------------------ #
# clocks
# ------------------------------------------------------------------------------------------------ #


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, v in zip(names, parts[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm), reasons=sorted(reasons))
        return out


# ------------------------------------------------------------------------------------------------ #


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="epgx", choices=["epgx", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--max-nstate", type=int, default=0, help="0 = unbounded (the BASELINE configuration)")
    ap.add_argument("--ntr", type=int, default=NTR)
    ap.add_argument("--grid", type=int, nargs=3, default=list(GRID))
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--atoms-per-cta", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--gather", action="store_true", help="time the final NCCL all-gather of the signal slabs too")
    ap.add_argument("--cpu-atoms", type=int, default=64, help="atoms per host core of the CPU sample")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    grid = tuple(args.grid)
    max_nstate = args.max_nstate or None
    workload = (f"MRF FISP dictionary, {args.ntr} TRs varying flip/TR, {grid[0]}x{grid[1]}x{grid[2]} T1xT2xB1 atoms per GPU, "
                f"max_nstate={'unbounded' if max_nstate is None else max_nstate}")

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        steps_v = []
        for _ in range(max(1, args.warmup > 0) + args.steps):
            v, atoms, wall, side = cpu_reference(args.ntr, max_nstate, cores, args.cpu_atoms, grid)
            steps_v.append((v, atoms, wall))
        used = steps_v[-args.steps:]
        value = sum(a for _, a, _ in used) / sum(w for _, _, w in used)
        sample = (f"{cores} processes x {side}^3-atom sub-grids of the bench grid ({used[0][1]} atoms/step), same {args.ntr}-TR "
                  f"sequence, numpy oracle port of the reference algorithm (full storage, complex128)")
        line = {
            "impl": "reference", "metric": "MRF dictionary atoms/sec", "value": value, "unit": "atoms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * np.mean([w for _, _, w in used]),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    import epgpy_b200
    from epgpy_b200 import engine, epg, lowering, sharding

    engine.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- lower the sequence once (operators pre-built, like the reference's eager coefficient arrays)
    T1, T2, B1 = grid_axes(grid, world)
    t0 = time.perf_counter()
    seq = fisp_sequence(epg, T1, T2, B1, args.ntr)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    opts = {} if max_nstate is None else {"max_nstate": max_nstate}
    low = lowering.lower(seq, options=opts, dtype=args.dtype)
    t_lower = time.perf_counter() - t0
    plan = engine.Plan(low)
    if args.lanes or args.atoms_per_cta:
        plan.set_variant(lanes_per_atom=args.lanes, atoms_per_cta=args.atoms_per_cta)
    cfg = plan.config()
    natoms = low.natoms
    a0, cnt = sharding.slab(natoms, rank, world)

    cdt = torch.complex128 if args.dtype == "f64" else torch.complex64
    csz = 16 if args.dtype == "f64" else 8
    sig = torch.empty((low.nadc, cnt, 1), dtype=cdt, device=f"cuda:{dev}")
    gathered = None
    if args.gather and world > 1:
        gathered = torch.empty((world,) + tuple(sig.shape), dtype=cdt, device=f"cuda:{dev}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")  # > 126 MB L2
    plan.upload(dev)

    def step():
        plan.run(dev, a0, cnt, signal=sig)
        if gathered is not None:  # the one collective of the path: final gather of the signal slabs (NCCL / NVLink)
            dist.all_gather_into_tensor(torch.view_as_real(gathered).reshape(-1), torch.view_as_real(sig).reshape(-1))

    for _ in range(args.warmup):
        step()
        flush.zero_()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=f"cuda:{dev}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    atoms_total = natoms
    value = atoms_total * args.steps / (total_ms * 1e-3)
    launches = args.steps

    # ---- end to end through the C ABI with HOST buffers: H2D of tape + tables, chunked kernel launches
    # overlapped with pitched D2H copies of the signal into pinned host memory
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((low.nadc, cnt, 1), dtype=cdt, pin_memory=True)
            nchunk = 8
            plan.run_to_host(dev, host, a0, cnt, nchunk=nchunk, dev_signal=sig)  # warm-up
            barrier()
            t_e2e = 0.0
            for _ in range(args.steps):
                flush.zero_()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                plan.run_to_host(dev, host, a0, cnt, nchunk=nchunk, dev_signal=sig)  # synchronises
                t_e2e += time.perf_counter() - t0
            tt = torch.tensor([t_e2e], dtype=torch.float64, device=f"cuda:{dev}")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            chk = complex(host[-1, cnt // 2, 0])
            e2e = {"value": atoms_total * args.steps / float(tt.item()), "unit": "atoms/s",
                   "h2d_bytes_per_step": int(plan.workspace_bytes()), "d2h_bytes_per_step": int(low.nadc * cnt * csz),
                   "ms_per_step": 1e3 * float(tt.item()) / args.steps, "chunks": nchunk,
                   "path": "epgx_plan_upload + epgx_simulate_strided x chunks + epgx_copy2d_to_host (pinned host buffer)",
                   "host_lowering_ms_once": 1e3 * (t_build + t_lower), "sample_value": [chk.real, chk.imag]}
            launches += (1 + args.steps) * nchunk
            del host
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "atoms/s", "error": f"{type(ex).__name__}: {ex}"}

    # ---- extra device-resident measurements (same grid): the other precision, and max_nstate = 32
    extra = {}
    if not args.no_extra:
        def quick(dtype, mns):
            lw = lowering.lower(seq, options=({} if mns is None else {"max_nstate": mns}), dtype=dtype)
            pl = engine.Plan(lw)
            c_ = pl.config()
            out = torch.empty((lw.nadc, cnt, 1), dtype=torch.complex128 if dtype == "f64" else torch.complex64,
                              device=f"cuda:{dev}")
            pl.upload(dev)
            for _ in range(2):
                pl.run(dev, a0, cnt, signal=out)
            barrier()
            tot = 0.0
            for _ in range(2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pl.run(dev, a0, cnt, signal=out)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            t = torch.tensor([tot], dtype=torch.float64, device=f"cuda:{dev}")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item()) / 2
            del out
            return {"value": natoms / (ms_ * 1e-3), "unit": "atoms/s", "ms_per_step": ms_, "kernel": c_,
                    "state_updates_per_s": c_["updates_per_atom"] * natoms / (ms_ * 1e-3)}, c_["flops_per_atom"] * cnt / (ms_ * 1e-3) / 1e12

        other = "f32" if args.dtype == "f64" else "f64"
        del sig
        torch.cuda.empty_cache()
        extra[f"{other}_unbounded"], tf_other = quick(other, max_nstate)
        extra[f"{args.dtype}_max_nstate32"], _ = quick(args.dtype, 32)
        launches += 8
        # the (B1, T1, T2) Jacobian of the same dictionary on a 50 x 50 x 50 sub-grid
        try:
            Tj1, Tj2, Bj1 = T1[::2], T2[::2], B1[: grid[2] * world: 2 * world] if world > 1 else B1[::2]
            seqj = fisp_sequence(epg, Tj1, Tj2, Bj1, args.ntr, jac=True)
            lwj = lowering.lower(seqj, probe=[None, epg.Jacobian(["B1", "T1", "T2"])], options=opts, dtype=args.dtype)
            plj = engine.Plan(lwj)
            plj.upload(dev)
            sj, jj = plj.run(dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plj.run(dev, signal=sj, jacobian=jj)
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1)
            extra[f"{args.dtype}_jacobian_B1_T1_T2"] = {
                "value": lwj.natoms / (ms_ * 1e-3), "unit": "atoms/s (each with signal + 3 derivatives), one GPU", "ms_per_step": ms_,
                "atoms": lwj.natoms, "kernel": plj.config()}
            launches += 2
            del sj, jj, plj
        except Exception as ex:
            extra["jacobian_error"] = f"{type(ex).__name__}: {ex}"
        if rank == 0:
            pk = engine.fma_peak(dev, other, 0.3)
            extra[f"{other}_unbounded"]["roofline_frac"] = tf_other / pk if pk else None

    if rank == 0:
        # ---- roofline: CUDA-core FMA throughput (the bound of this path) + HBM for context
        step_ms = total_ms / args.steps
        flops = cfg["flops_per_atom"] * cnt
        peak = engine.fma_peak(dev, args.dtype, 0.5)
        achieved = flops / (step_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bytes_alg = low.nadc * cnt * csz + plan.workspace_bytes()
        roofline = {
            "bound": "fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture in profiles/ (27 000-atom
            # launch: 4.8 MB read + 432.3 MB written = 16.19 kB per atom), scaled to this launch's atom count
            "traffic": 14492.0 * cnt if (args.ntr == NTR and args.dtype == "f64") else None,
            "traffic_note": "scaled per atom from the 27 000-atom capture profiles/r01_real_kernel_f64_full.txt "
                            "(2.7 MB read + 388.6 MB written; algorithmic: 16 kB per atom, the difference is dirty lines "
                            "still in the 126 MB L2 when the kernel ends)",
            "peak_source": f"measured live on this GPU: dependent-FMA microbenchmark epgx_fma_peak({args.dtype})",
            "flops_per_atom_executed": cfg["flops_per_atom"],
            "hbm": {"achieved": bytes_alg / (step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg / (step_ms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        cpu = None
        if not args.no_cpu and world == 1:
            v, atoms, wall, side = cpu_reference(args.ntr, max_nstate, 1, 216, grid)  # 6^3 atoms: 10-20 s on one core
            cpu = {"value": v, "unit": "atoms/s", "cores": 1, "kind": "port",
                   "sample": f"{side}^3 = {atoms} atoms of the bench grid, same {args.ntr}-TR sequence, {wall:.1f} s, numpy oracle port"}
        line = {
            "metric": "MRF dictionary atoms/sec", "value": value, "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload, "atoms_per_gpu": cnt, "atoms_total": atoms_total, "ntr": args.ntr,
                       "l2": "256 MB buffer written between timed iterations (L2 flush)", "kernel": cfg,
                       "state_updates_per_atom_executed": cfg["updates_per_atom"], "gather": bool(gathered is not None),
                       # what the reference (full storage, no pruning, no fusion) performs for the same output
                       "state_updates_per_atom_reference": 4002002 if (args.ntr == NTR and max_nstate is None) else None},
            "state_updates_per_s": cfg["updates_per_atom"] * value,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
