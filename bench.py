#!/usr/bin/env python
"""bench.py -- MRF FISP dictionary generation on B200 (BASELINE.json configs[2]).

One "step" = one pass of the hot path over the whole dictionary: the 1000-TR FISP sequence
(`[T(180,0), E(20)] + [T(FA_i*B1, 90), E(3), ADC, E(TR_i-3), S(1)] x 1000`) for
100 x 100 x 100 = 1 M (T1, T2, B1) atoms, unbounded number of states, FP64.

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference ...                      # CPU arm: the UNMODIFIED reference (baseline/_ref)

N > 1 (torchrun, one process per GPU): STRONG scaling -- the same 1 M-atom grid is cut in N contiguous slabs of the
flattened grid, every rank simulates its slab and the timed step ends with the one collective of the path, the
NCCL all-gather of the signal slabs (16 GB in total), so that every rank holds the whole dictionary
(epgpy_b200.sharding.run_gather).  `--scaling weak` keeps 1 M atoms per GPU instead (B1 axis x N, no gather).

Fields of the JSON line ("Measurement"):
  value     atoms/s, tape + coefficient table resident in HBM, CUDA events around each step, max over ranks
  e2e       atoms/s through the public API: epg.simulate(sequence) -- host lowering of the 5 002 operators, plan
            creation, H2D of the tape, chunked kernel launches overlapped with pitched D2H copies into the pinned result
            buffer, zero-copy reshape -- wall clock, every step a fresh call (no plan cache), max over ranks
  parity    >= 32 scattered atoms of the timed 1 M-atom output against the CPU oracle (outside the timed region)
  roofline  CUDA-core FMA throughput of the kernel against the live dependent-FMA microbenchmark
  cpu_baseline / --impl reference: the unmodified reference package on the host cores (baseline/_ref, installed from
            /root/reference by __graft_entry__.build(); the numpy oracle port only if it is missing)
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

NTR = 1000
GRID = (100, 100, 100)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
# sub-grid of the CPU arm: pairwise DIFFERENT axis sizes -- the reference's vectorised matmul mis-aligns operator
# axes when grid axes have equal sizes, so a 4 x 4 x 4 sample would time (and return) garbage
CPU_SUB = (4, 3, 5)
CPU_SUB_1CORE = (6, 5, 4)


def fisp_schedule(ntr, seed=0):
    rng = np.random.RandomState(seed)
    fa = 10 + 50 * np.abs(np.sin(np.arange(ntr) * np.pi / 200))
    tr = rng.uniform(11, 16, ntr)
    return fa, tr


def fisp_sequence(epg, T1, T2, B1, ntr=NTR, jac=False, flat=False):
    """SURVEY.md 8d M3 (the reference-API form of BASELINE configs[2]); axes: T1 -> 0, T2 -> 1, B1 -> 2
    (flat=True: the three parameter vectors share axis 0 -- a list of scattered atoms)"""
    fa, tr = fisp_schedule(ntr)
    T1 = np.asarray(T1, dtype=float)
    T2 = np.asarray(T2, dtype=float) if flat else np.asarray(T2, dtype=float)[None, :]
    B1 = np.asarray(B1, dtype=float) if flat else np.asarray(B1, dtype=float)[None, None, :]
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
# CPU arm: the unmodified reference (baseline/_ref); the numpy oracle port is the fallback
# ------------------------------------------------------------------------------------------------ #


def cpu_impl():
    """(kind, namespace, simulate(seq, max_nstate)) of the CPU arm"""
    if os.path.isdir(os.path.join(REF_DIR, "epgpy")):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import epgpy  # the UNMODIFIED reference package (pip --target install of /root/reference)

        def run(seq, max_nstate):
            return epgpy.simulate(seq, **({} if max_nstate is None else {"max_nstate": max_nstate}))

        return "reference", epgpy, run
    import oracle_api

    return "port", oracle_api.epg, lambda seq, max_nstate: oracle_api.O.simulate(seq, max_nstate=max_nstate)


def _cpu_worker(args):
    ntr, max_nstate, t1, t2, b1 = args
    kind, ns, run = cpu_impl()
    seq = fisp_sequence(ns, t1, t2, b1, ntr)  # operators pre-built: only `simulate` is timed
    t0 = time.perf_counter()
    out = run(seq, max_nstate)
    return time.perf_counter() - t0, int(np.prod(np.shape(out)[1:]))


def cpu_reference(ntr, max_nstate, cores, sub, grid=GRID):
    """time the CPU arm on `cores` processes, each simulating its own `sub`-shaped sub-grid taken from the bench grid;
    returns (atoms/s over the wall clock, atoms, seconds)"""
    import multiprocessing as mp

    T1, T2, B1 = grid_axes(grid)
    jobs = []
    rng = np.random.RandomState(1)
    for c in range(cores):
        i, j, k = (rng.randint(0, max(1, n - s)) for n, s in zip(grid, sub))
        jobs.append((ntr, max_nstate, T1[i:i + sub[0]], T2[j:j + sub[1]], B1[k:k + sub[2]]))
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    atoms = sum(r[1] for r in res)
    return atoms / wall, atoms, wall


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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--gather-chunks", type=int, default=8)
    ap.add_argument("--no-gather", action="store_true", help="N > 1: leave the final gather out of the timed step")
    ap.add_argument("--gather-impl", default="auto", choices=["auto", "p2p", "nccl"],
                    help="p2p: copy-engine pushes into peer windows (CUDA IPC over NVLink); nccl: chunked all-gather; auto: p2p "
                         "up to four ranks, nccl above (measured: 122.6 / 65.2 / 53.2 ms with p2p at N = 2 / 4 / 8 against "
                         "129.9 / 71.6 / 43.8 ms with nccl, profiles/scale_r02.json and scale_r02b_p2p.json)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--parity-atoms", type=int, default=32)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    grid = tuple(args.grid)
    max_nstate = args.max_nstate or None
    weak = args.scaling == "weak" and world > 1
    workload = (f"MRF FISP dictionary, {args.ntr} TRs varying flip/TR, {grid[0]}x{grid[1]}x{grid[2]} T1xT2xB1 atoms"
                f"{' per GPU' if weak else ' in total'}, max_nstate={'unbounded' if max_nstate is None else max_nstate}")

    if args.impl == "reference":
        if rank != 0:
            return 0
        kind, _, _ = cpu_impl()
        cores = os.cpu_count() or 1
        runs = []
        for _ in range(min(1, args.warmup) + args.steps):
            runs.append(cpu_reference(args.ntr, max_nstate, cores, CPU_SUB, grid))
        used = runs[-args.steps:]
        value = sum(a for _, a, _ in used) / sum(w for _, _, w in used)
        what = "the unmodified reference package epgpy (baseline/_ref)" if kind == "reference" else "numpy oracle port of the reference algorithm"
        sample = (f"{cores} processes x {'x'.join(map(str, CPU_SUB))}-atom sub-grids of the bench grid ({used[0][1]} atoms/step), same "
                  f"{args.ntr}-TR sequence, {what}, full storage, complex128; linear in atoms (BASELINE.md section 3)")
        line = {
            "impl": "reference", "metric": "MRF dictionary atoms/sec", "value": value, "unit": "atoms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([w for _, _, w in used])),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": cores, "kind": kind, "sample": sample},
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
        # NCCL's copy kernels on a HIGH-PRIORITY stream: the simulation fills every SM with CTAs that hold all its
        # registers, and at equal priority the all-gather of chunk j only got its CTAs placed once the kernels of all later
        # chunks had drained (measured: the whole gather time showed up as a tail behind the last kernel)
        opts_pg = None
        try:
            opts_pg = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev), pg_options=opts_pg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the sequence and its plan (device-timed part: operators, tape and tables are resident)
    T1, T2, B1 = grid_axes(grid, world if weak else 1)
    t0 = time.perf_counter()
    seq = fisp_sequence(epg, T1, T2, B1, args.ntr)
    t_build = time.perf_counter() - t0
    opts = {} if max_nstate is None else {"max_nstate": max_nstate}
    t0 = time.perf_counter()
    low = lowering.lower(seq, options=opts, dtype=args.dtype)
    t_lower = time.perf_counter() - t0
    plan = engine.Plan(low)
    if args.lanes or args.atoms_per_cta:
        plan.set_variant(lanes_per_atom=args.lanes, atoms_per_cta=args.atoms_per_cta)
    cfg = plan.config()
    natoms = low.natoms
    a0, cnt = sharding.slab(natoms, rank, world)
    gather = world > 1 and not weak and not args.no_gather

    cdt = torch.complex128 if args.dtype == "f64" else torch.complex64
    csz = 16 if args.dtype == "f64" else 8
    window, gather_impl = None, None
    if gather and (args.gather_impl == "p2p" or (args.gather_impl == "auto" and world <= 4)):
        try:  # every rank must agree: a failure anywhere sends all of them to the NCCL path
            window = sharding.PeerWindow(low, dev)
            ok = torch.ones(1, device=f"cuda:{dev}")
        except Exception as ex:
            window, ok = None, torch.zeros(1, device=f"cuda:{dev}")
            print(f"rank {rank}: peer window unavailable ({type(ex).__name__}: {ex}); using the NCCL all-gather", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok.item()) and window is not None:
            window = None
    if gather:
        gather_impl = "p2p" if window is not None else "nccl"
    sig = window.tensor if window is not None else torch.empty((low.nadc, natoms if gather else cnt, 1), dtype=cdt, device=f"cuda:{dev}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")  # > 126 MB L2
    plan.upload(dev)

    def step():
        if window is not None:  # slab kernels in chunks, each finished chunk pushed into every peer's window (copy engines)
            sharding.run_gather_p2p(plan, window, nchunk=args.gather_chunks)
        elif gather:  # slab kernels in chunks + NCCL all-gather of the signal slabs
            sharding.run_gather(plan, dev, nchunk=args.gather_chunks, out=sig)
        else:
            plan.run(dev, a0, cnt, signal=sig)

    for _ in range(args.warmup):
        step()
        flush.zero_()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    evs = []
    l0 = engine.LAUNCHES
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = engine.LAUNCHES - l0
    clocks = sampler.stop() if rank == 0 else {}
    total_ms = rank_max(sum(a.elapsed_time(b) for a, b in evs))
    value = natoms * args.steps / (total_ms * 1e-3)
    dev_sample = None
    if rank == 0:
        dev_sample = sig[:, :: max(1, sig.shape[1] // 97), 0].cpu().numpy()  # for the e2e == device cross-check

    # ---- end to end through the PUBLIC API: epg.simulate(sequence), lowering included, host result
    e2e, parity = None, None
    if not args.no_e2e:
        try:
            del sig
            torch.cuda.empty_cache()
            kw = dict(opts, dtype="float64" if args.dtype == "f64" else "float32", device=dev)
            if world > 1:
                kw["shard"] = (rank, world)
            t0 = time.perf_counter()
            out = epg.simulate(seq, **kw)  # first call: pins the result buffer (torch keeps the pages for the next calls)
            t_first = time.perf_counter() - t0
            del out
            barrier()
            t_e2e = 0.0
            l0, h0, d0 = engine.LAUNCHES, engine.H2D_BYTES, engine.D2H_BYTES
            for i in range(args.steps):
                flush.zero_()
                barrier()
                t0 = time.perf_counter()
                out = epg.simulate(seq, **kw)  # synchronises; returns a host array (nADC, *grid) / (nADC, slab atoms)
                t_e2e += time.perf_counter() - t0
                if i + 1 < args.steps:
                    del out  # the pinned pages go back to torch's host cache and serve the next call
            launches += engine.LAUNCHES - l0
            t_e2e = rank_max(t_e2e)
            real_rows = (engine.D2H_BYTES - d0) < args.steps * low.nadc * cnt * csz
            e2e = {"value": natoms * args.steps / t_e2e, "unit": "atoms/s",
                   # counted by the binding from the copies it issued (tape + tables in; result rows out: REAL parts only when
                   # the signal is real-valued, widened to complex128 by host threads -- Plan.run_to_host_real)
                   "h2d_bytes_per_step": int((engine.H2D_BYTES - h0) // args.steps), "d2h_bytes_per_step": int((engine.D2H_BYTES - d0) // args.steps),
                   "real_rows_over_pcie": bool(real_rows),
                   "ms_per_step": 1e3 * t_e2e / args.steps, "first_call_ms": 1e3 * t_first,
                   "path": "epg.simulate(sequence" + (", shard=(rank, world))" if world > 1 else ")") +
                           ": lowering + epgx_plan_create + epgx_plan_upload + chunked kernel launches overlapped with the D2H copy of "
                           "the previous chunk (real-valued signal: epgx_simulate_real -> pinned staging -> epgx_expand_real into "
                           "the complex128 result; else epgx_simulate_strided -> epgx_copy2d_to_host) + zero-copy reshape; a fresh "
                           "call per step (no plan cache)",
                   "host_lowering_ms": 1e3 * t_lower, "operator_construction_ms_outside": 1e3 * t_build,
                   "result_shape": list(out.shape), "result_dtype": str(out.dtype)}
            # ---- parity of the TIMED output: scattered atoms against the CPU oracle (outside the timed region)
            if rank == 0 and args.parity_atoms > 0:
                import oracle_api

                rng = np.random.RandomState(7)
                flat = out.reshape(out.shape[0], -1)  # (nADC, atoms of this rank), C order of the flattened grid
                pick = np.unique(np.concatenate([[0, flat.shape[1] - 1], rng.randint(0, flat.shape[1], args.parity_atoms)]))
                ii, jj, kk = np.unravel_index(pick + a0, low.atom_shape)
                ref = oracle_api.O.simulate(fisp_sequence(oracle_api.epg, T1[ii], T2[jj], B1[kk], args.ntr, flat=True),
                                            max_nstate=max_nstate)
                got = flat[:, pick]
                err = float(np.abs(got - ref).max() / np.abs(ref).max())
                tol = 1e-10 if args.dtype == "f64" else 1e-5
                cross = float(np.abs(flat[:, :: max(1, flat.shape[1] // 97)][:, :dev_sample.shape[1]] - dev_sample).max()) if world == 1 else None
                parity = {"atoms_checked": int(len(pick)), "parity_max_rel": err, "tol": tol, "ok": bool(err <= tol),
                          "against": "oracle/epg_oracle.py (numpy restatement of the reference, pinned by tests/golden)",
                          "e2e_vs_device_max_abs": cross}
            del out
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "atoms/s", "error": f"{type(ex).__name__}: {ex}"}

    # ---- extra device-resident measurements (rank-local slab of the same grid): the other precision, max_nstate = 32,
    # and the (B1, T1, T2) Jacobian
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
            ms_ = rank_max(tot) / 2
            del out
            return {"value": natoms / (ms_ * 1e-3), "unit": "atoms/s", "ms_per_step": ms_, "kernel": c_, "gather": False,
                    "state_updates_per_s": c_["updates_per_atom"] * natoms / (ms_ * 1e-3)}, c_["flops_per_atom"] * cnt / (ms_ * 1e-3) / 1e12

        other = "f32" if args.dtype == "f64" else "f64"
        torch.cuda.empty_cache()
        extra[f"{other}_unbounded"], tf_other = quick(other, max_nstate)
        extra[f"{args.dtype}_max_nstate32"], _ = quick(args.dtype, 32)
        # the (B1, T1, T2) Jacobian of the same dictionary on a 50 x 50 x 50 sub-grid, one GPU
        try:
            Tj1, Tj2, Bj1 = T1[::2], T2[::2], B1[::2][: grid[2] // 2]
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
            cj = plj.config()
            extra[f"{args.dtype}_jacobian_B1_T1_T2"] = {
                "value": lwj.natoms / (ms_ * 1e-3), "unit": "atoms/s (each with signal + 3 derivatives), one GPU", "ms_per_step": ms_,
                "atoms": lwj.natoms, "kernel": cj}
            del sj, jj, plj
        except Exception as ex:
            extra["jacobian_error"] = f"{type(ex).__name__}: {ex}"
        if rank == 0:
            pk = engine.fma_peak(dev, other, 0.3)
            extra[f"{other}_unbounded"]["roofline_frac"] = tf_other / pk if pk else None

    if rank == 0:
        # ---- roofline of the dominant kernel: CUDA-core FMA throughput (the bound of this path); the
        # kernel time is measured live with CUDA events around single launches over this rank's slab
        kms = []
        out = torch.empty((low.nadc, cnt, 1), dtype=cdt, device=f"cuda:{dev}")
        for i in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(dev, a0, cnt, signal=out)
            e1.record()
            torch.cuda.synchronize()
            kms.append(e0.elapsed_time(e1))
        del out
        kernel_ms = float(np.mean(kms[1:]))
        step_ms = total_ms / args.steps
        flops = cfg["flops_per_atom"] * cnt
        peak = engine.fma_peak(dev, args.dtype, 0.5)
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bytes_alg = low.nadc * cnt * csz + plan.workspace_bytes()
        roofline = {
            "bound": "fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "kernel": "real_kernel" if cfg["kernel"] == 2 else f"kernel variant {cfg['kernel']}",
            "kernel_ms_per_launch": kernel_ms, "kernel_share_of_step": kernel_ms / step_ms,
            # dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture in profiles/ (27 000-atom
            # launch), scaled to this launch's atom count
            "traffic": 14545.0 * cnt if (args.ntr == NTR and args.dtype == "f64") else None,
            "traffic_note": "scaled per atom from the 27 000-atom capture profiles/r02d_real_kernel_f64_full.txt "
                            "(2.7 MB read + 390.0 MB written; algorithmic: 16 kB per atom, the difference is dirty lines "
                            "still in the 126 MB L2 when the kernel ends)",
            "peak_source": f"measured live on this GPU: dependent-FMA microbenchmark epgx_fma_peak({args.dtype})",
            "flops_per_atom_executed": cfg["flops_per_atom"],
            # the real-valued kernels apply a pulse in 7 instructions / 11 flops per order since round 2 (shared
            # q = b s + u Z); rounds 1 counted the row-by-row form, 14 flops per order: the same launch at that count
            "frac_at_round1_flop_count": (achieved * 14.0 / 11.0 / peak if peak else None) if cfg["kernel"] == 2 else None,
            "hbm": {"achieved": bytes_alg / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        cpu = None
        if not args.no_cpu and world == 1:
            kind, _, _ = cpu_impl()
            v, atoms, wall = cpu_reference(args.ntr, max_nstate, 1, CPU_SUB_1CORE, grid)  # 120 atoms: 10-20 s on one core
            cpu = {"value": v, "unit": "atoms/s", "cores": 1, "kind": kind,
                   "sample": f"{'x'.join(map(str, CPU_SUB_1CORE))} = {atoms} atoms of the bench grid, same {args.ntr}-TR sequence, {wall:.1f} s, "
                             + ("unmodified reference package (baseline/_ref)" if kind == "reference" else "numpy oracle port")}
        line = {
            "metric": "MRF dictionary atoms/sec", "value": value, "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak" if weak else "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload, "atoms_per_gpu": cnt, "atoms_total": natoms, "ntr": args.ntr,
                       "l2": "256 MB buffer written between timed iterations (L2 flush)", "kernel": cfg,
                       "state_updates_per_atom_executed": cfg["updates_per_atom"], "gather": bool(gather),
                       "gather_impl": gather_impl,
                       "gather_note": ("final gather of the signal slabs inside the timed step, in %d chunks overlapped with the slab "
                                       "kernels (p2p: copy-engine pushes into CUDA-IPC peer windows over NVLink + a closing NCCL "
                                       "barrier; nccl: chunked all-gather); every rank ends with the whole [nadc][atoms] dictionary"
                                       % args.gather_chunks) if gather else None,
                       # what the reference (full storage, no pruning, no fusion) performs for the same output
                       "state_updates_per_atom_reference": 4002002 if (args.ntr == NTR and max_nstate is None) else None},
            "state_updates_per_s": cfg["updates_per_atom"] * value,
            "clocks": clocks, "e2e": e2e, "parity": parity, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "extra": extra,
        }
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            print(f"PARITY FAILURE: max rel err {parity['parity_max_rel']:.3e} > {parity['tol']}", file=sys.stderr)
            return 1
    if window is not None:
        window.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
