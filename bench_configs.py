#!/usr/bin/env python
"""Secondary measurements: every BASELINE.json configuration (and the Jacobian variants of SURVEY.md 8d)
through the engine, device-resident, CUDA-event timed.  Not the driver's bench (that is bench.py); used to
track the kernels that are not on the headline path.  Prints one JSON line per configuration."""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import bench  # noqa: E402


def cfg_readme(epg):
    import cases
    return cases.readme_mse(epg)["seq"], {}, None


def cfg_mse(epg):
    T2 = np.linspace(20, 300, 200)
    B1 = np.linspace(0.5, 1.2, 100)[None, :]
    T1 = np.linspace(500, 3000, 10)[None, None, :]
    exc, rfc, rlx, sh = epg.T(90 * B1, 90), epg.T(180 * B1, 0), epg.E(4.75, T1, T2), epg.S(1)
    return [exc] + [[sh, rlx, rfc, sh, rlx, epg.ADC]] * 17, {}, None


def cfg_fisp(epg, n=(100, 100, 100), ntr=1000, jac=False):
    T1, T2, B1 = bench.grid_axes(n)
    return bench.fisp_sequence(epg, T1, T2, B1, ntr, jac=jac), {}, (["B1", "T1", "T2"] if jac else None)


def cfg_fisp_pulse_jac(epg, ntr=1000, max_nstate=10):
    """M3J(ii): per-pulse flip-angle variables on 4 x 4 x 4 atoms, max_nstate = 10 (optim_mrf.py:96)"""
    fa, tr = bench.fisp_schedule(ntr)
    T1 = np.linspace(300, 3000, 4)
    T2 = np.linspace(20, 300, 4)[None, :]
    B1 = np.linspace(0.7, 1.2, 4)[None, None, :]
    names = [f"a{i:04d}" for i in range(ntr)]
    seq = [epg.T(180, 0), epg.E(20, T1, T2)]
    for i in range(ntr):
        seq.append([epg.T(fa[i] * B1, 90, order1={names[i]: {"alpha": B1}}), epg.E(3, T1, T2), epg.ADC,
                    epg.E(tr[i] - 3, T1, T2), epg.S(1)])
    return seq, {"max_nstate": max_nstate}, names


def cfg_optim_mrf(epg, ntr=400, max_nstate=10):
    """the reference's sequence-optimisation inner loop (examples/differentiation/optim_mrf.py:96-149): ONE atom,
    400 TRs, one flip-angle and one repetition-time variable per TR (800 variables), max_nstate = 10"""
    fa, tr = bench.fisp_schedule(ntr)
    names = [f"a{i:04d}" for i in range(ntr)] + [f"t{i:04d}" for i in range(ntr)]
    seq = [epg.T(180, 0), epg.E(20, 1000.0, 80.0)]
    for i in range(ntr):
        seq.append([epg.T(fa[i], 90, order1={names[i]: "alpha"}), epg.E(3, 1000.0, 80.0), epg.ADC,
                    epg.E(tr[i] - 3, 1000.0, 80.0, order1={names[ntr + i]: "tau"}), epg.S(1)])
    return seq, {"max_nstate": max_nstate}, names


def cfg_gre_diffusion(epg, ntr=500):
    T1 = np.linspace(400, 2000, 200)
    T2 = np.linspace(30, 200, 200)[None, :]
    seq, kv = [], [2, 1, -1]
    for n in range(ntr):
        ph = 117.0 * n * (n + 1) / 2
        seq.append([epg.T(15, ph), epg.E(2, T1, T2), epg.Adc(phase=-ph), epg.E(8, T1, T2), epg.S(kv), epg.D(10, 2e-3, k=kv)])
    return seq, {"kvalue": 500.0}, None


def cfg_mt_bssfp(epg, ntr=500):
    T1, T2, khi, f = [779.0, 779.0], [45.0, 12e-3], 4.3e-3, [1 - 0.117, 0.117]
    kmat = epg.exchange_matrix(khi, densities=f)
    offres = 1 / 5.0 * np.linspace(-0.5, 0.5, 101)
    FA = np.linspace(5, 60, 200)[None, None, :]
    pool = np.array([1.0, 0.0])[:, None, None]
    sat = epg.R(rL=[0, 0.0316])
    exg = epg.X(5.0, kmat, T1=T1, T2=T2, g=[offres])
    rf = [epg.T(FA * pool, 0.0) @ sat, epg.T(FA * pool, 180.0) @ sat]
    seq = [[rf[i % 2], exg] for i in range(ntr)] + [rf[0], epg.Adc(reduce=0)]
    return seq, {"init": epg.StateMatrix(density=f)}, None


def cfg_mt_bssfp_pulse_jac(epg, ntr=200):
    """BASELINE configs[4] with its Jacobian: per-pulse flip-angle variables (both pools), partial states carried through
    the exchange operator; 101 off-resonance x 16 flip-angle atoms"""
    T1, T2, khi, f = [779.0, 779.0], [45.0, 12e-3], 4.3e-3, [1 - 0.117, 0.117]
    kmat = epg.exchange_matrix(khi, densities=f)
    offres = 1 / 5.0 * np.linspace(-0.5, 0.5, 101)
    FA = np.linspace(5, 60, 16)[None, None, :]
    pool = np.array([1.0, 0.0])[:, None, None]
    sat = epg.R(rL=[0, 0.0316])
    exg = epg.X(5.0, kmat, T1=T1, T2=T2, g=[offres])
    names = [f"a{i:04d}" for i in range(ntr)]
    seq = []
    for i in range(ntr):
        seq += [epg.T(FA * pool, 0.0 if i % 2 == 0 else 180.0, order1={names[i]: {"alpha": pool}}) @ sat, exg, epg.Adc(reduce=0)]
    return seq, {"init": epg.StateMatrix(density=f), "propagate_nondiff": True}, names


CONFIGS = {
    "C1_readme_mse": cfg_readme,
    "C2_mse_grid_200k": cfg_mse,
    "C3_fisp_1M": cfg_fisp,
    "C3J_fisp_jac_B1_T1_T2_125k": lambda epg: cfg_fisp(epg, (50, 50, 50), 1000, jac=True),
    "C3J_fisp_pulse_jac_64_atoms_1000_vars": cfg_fisp_pulse_jac,
    "C3J_optim_mrf_1_atom_800_vars_400_TR": cfg_optim_mrf,
    "C4_gre_diffusion_40k": cfg_gre_diffusion,
    "C5_mt_bssfp_20k": cfg_mt_bssfp,
    "C5J_mt_bssfp_pulse_jac_1616_atoms_200_vars": cfg_mt_bssfp_pulse_jac,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--lanes", type=int, default=0, help="force lanes per atom (tuning)")
    ap.add_argument("--atoms-per-cta", type=int, default=0, help="force atoms per CTA (tuning)")
    args = ap.parse_args()
    import torch

    from epgpy_b200 import engine, epg, lowering

    for name, fn in CONFIGS.items():
        if args.only and not any(o in name for o in args.only):
            continue
        t0 = time.perf_counter()
        seq, opts, jac = fn(epg)
        opts = dict(opts)
        init = opts.pop("init", None)
        extra = {"propagate_nondiff": True} if opts.pop("propagate_nondiff", False) else {}
        probe = [None, epg.Jacobian(jac)] if jac else None
        low = lowering.lower(seq, init=init, probe=probe, options=opts, dtype=args.dtype, **extra)
        t_host = time.perf_counter() - t0
        plan = engine.Plan(low)
        if args.lanes or args.atoms_per_cta:
            plan.set_variant(lanes_per_atom=args.lanes, atoms_per_cta=args.atoms_per_cta)
        cfg = plan.config()
        dev = torch.cuda.current_device()
        sig, jc = plan.run(dev)  # warm-up (allocates)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(dev, signal=sig, jacobian=jc)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        best = min(ms)
        peak = engine.fma_peak(dev, args.dtype, 0.2)
        # flops of the tape (base state counted once: a lower bound when the variables are tiled and the base recomputed)
        tf = cfg["flops_per_atom"] * low.natoms / (best * 1e-3) / 1e12
        print(json.dumps({
            "config": name, "dtype": args.dtype, "atoms": low.natoms, "npool": low.npool, "nvar": low.nvar, "nadc": low.nadc,
            "max_order": low.max_order, "ms": best, "atoms_per_s": low.natoms / (best * 1e-3),
            "state_updates_per_s": cfg["updates_per_atom"] * low.natoms / (best * 1e-3),
            "tflops_executed": tf, "fma_frac": tf / peak if peak else None, "host_lowering_s": t_host, "kernel": cfg}))
        del sig, jc, plan
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
