"""Generate golden vectors by running the UNMODIFIED reference (py-baudin/epgpy).

Run in the build container only (the reference lives read-only at /root/reference
and does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/<case>.npz with `signal` (and `jacobian`) for every case of
tests/cases.py, plus `primitives.npz` with operator-level known answers
(coefficient arrays and single-operator state updates) used to pin the oracle's
individual functions.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))  # tests/
sys.path.insert(0, os.environ.get("EPGPY_REFERENCE", "/root/reference"))

import epgpy  # noqa: E402  (the reference)
from epgpy import diffusion, evolution, exchange, statematrix, transition  # noqa: E402

import cases  # noqa: E402


def main():
    ns = cases.namespace(epgpy)
    for name, fn in cases.CASES.items():
        case = fn(ns)
        sig, jac = cases.run_api(ns, case)
        out = {"signal": sig, "shape": np.asarray(epgpy.core.getshape(case["seq"]))}
        if jac is not None:
            out["jacobian"] = jac
        times = epgpy.core.get_adc_times(case["seq"])
        try:
            out["times"] = np.asarray(times, dtype=float)
        except (ValueError, TypeError):
            pass
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(f"{name:20s} signal {sig.shape} jac {None if jac is None else jac.shape}")

    # ---- equal-size grid axes: per-atom SCALAR runs of the reference (its vectorised run is wrong there)
    case = cases.fisp_equal_axes(ns)
    T1, T2, B1 = case["axes"]
    sig = jac = None
    for i, t1 in enumerate(T1):
        for j, t2 in enumerate(T2):
            for k, b1 in enumerate(B1):
                s1, j1 = epgpy.simulate(case["build"](float(t1), float(t2), float(b1)), probe=[None, epgpy.core.Jacobian(case["jac"])])
                s1, j1 = np.asarray(s1), np.asarray(j1)
                if sig is None:
                    sig = np.zeros((s1.shape[0], len(T1), len(T2), len(B1)), dtype=complex)
                    jac = np.zeros(sig.shape + (j1.shape[-1],), dtype=complex)
                sig[:, i, j, k], jac[:, i, j, k] = s1[:, 0], j1[:, 0]
    vec = np.asarray(epgpy.simulate(case["build"](T1, T2[None, :], B1[None, None, :]), probe=None))
    np.savez_compressed(os.path.join(HERE, "fisp_equal_axes.npz"), signal=sig, jacobian=jac,
                        reference_vectorised_rel_diff=np.abs(vec - sig).max() / np.abs(sig).max())
    print(f"fisp_equal_axes      signal {sig.shape} jac {jac.shape}; the reference's own vectorised run differs by "
          f"{np.abs(vec - sig).max() / np.abs(sig).max():.3f} (relative)")

    # ---- order-2 derivatives: Hessian probes
    for name, fn in cases.HESSIAN_CASES.items():
        sig, hes = cases.run_hessian(ns, fn(ns))
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), signal=sig, hessian=hes)
        print(f"{name:20s} signal {sig.shape} hessian {hes.shape}")

    # ---- Fourier probes (DFT / Imaging) and gradient / time-accumulation shifts: one array per probe event
    for name, fn in cases.FOURIER_CASES.items():
        vals = cases.run_probes(ns, fn(ns))
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **{f"probe{i}": np.asarray(v) for i, v in enumerate(vals)})
        print(f"{name:20s}", len(vals), "probe events", sorted({np.asarray(v).shape for v in vals}))

    # ---- `probe=` expressions
    case = cases.probe_expr(ns)
    vals = epgpy.simulate(case["seq"], probe=case["probe"])
    np.savez_compressed(os.path.join(HERE, "probe_expr.npz"), **{f"probe{i}": np.asarray(v) for i, v in enumerate(vals)})
    print("probe_expr          ", [np.asarray(v).shape for v in vals])

    # ---- operator-level primitives
    rng = np.random.RandomState(1)
    alpha = rng.uniform(-180, 180, (4, 1))
    phi = rng.uniform(-180, 180, (1, 3))
    tau = rng.uniform(1, 20, (3,))
    T1 = rng.uniform(200, 2000, (1, 2))
    T2 = rng.uniform(20, 200, (1, 1, 4))
    g = rng.uniform(-0.1, 0.1, (3, 1, 1))
    prim = dict(alpha=alpha, phi=phi, tau=tau, T1=T1, T2=T2, g=g)
    prim["rf"] = transition.rotation_operator(alpha, phi)
    prim["rf_dalpha"] = transition.rotation_d_alpha(alpha, phi)
    prim["rf_dphi"] = transition.rotation_d_phi(alpha, phi)
    arr, arr0 = evolution.relaxation_operator(tau, T1, T2, g)
    prim["relax_arr"], prim["relax_arr0"] = arr, arr0
    for p, f in (("tau", evolution.relaxation_d_tau), ("T1", evolution.relaxation_d_T1),
                 ("T2", evolution.relaxation_d_T2), ("g", evolution.relaxation_d_g)):
        d, d0 = f(tau, T1, T2, g)
        prim[f"relax_d_{p}"] = d
        if d0 is not None:
            prim[f"relax_d0_{p}"] = d0
    parr, _ = evolution.precession_operator(tau, g)
    prim["prec_arr"] = parr
    # (scalar tau: with two arrays the reference's precession_d_* broadcast right-aligned)
    prim["prec_d_tau"] = evolution.precession_d_tau(3.7, g)[0]
    prim["prec_d_g"] = evolution.precession_d_g(3.7, g)[0]
    earr, earr0 = evolution.evolution_operator(0.1 + 0.2j, 0.3, 0.05)
    prim["evol_arr"], prim["evol_arr0"] = earr, earr0
    # diffusion
    prim["bmat_1"] = diffusion.compute_bmatrix(1.5, [[1e3, 2e3, -5e2], [0, 1e3, 2e3]])
    prim["bmat_2"] = diffusion.compute_bmatrix(1.5, [[1e3, 2e3, -5e2], [0, 1e3, 2e3]], [[2e3, 2.5e3, -1e3], [1e3, 1.5e3, 1.5e3]])
    Dten = np.array([[2.0, 0.1, 0], [0.1, 1.0, 0.2], [0, 0.2, 0.5]]) * 1e-3
    DL, DT = diffusion.diffusion_operator(prim["bmat_1"], prim["bmat_2"], Dten)
    prim["Dten"], prim["diff_DL"], prim["diff_DT"] = Dten, DL, DT
    DL, DT = diffusion.diffusion_operator(prim["bmat_1"], prim["bmat_2"], 1.3e-3)
    prim["diff_DL_iso"], prim["diff_DT_iso"] = DL, DT
    # exchange
    kmat = exchange.exchange_matrix(4.3e-3, densities=[0.883, 0.117])
    prim["kmat"] = kmat
    xop = exchange.X(5.0, kmat, T1=[779.0, 779.0], T2=[45.0, 12e-3], g=[np.linspace(-0.1, 0.1, 5)])
    prim["xmat"] = xop.mat
    prim["xaxis"] = np.asarray(xop.axis)
    prim["kmat3"] = exchange.exchange_matrix([1e-3, 2e-3], ncomp=3)
    # single-operator state updates on a random valid state (docs/basics.md:199-216 style)
    n = 3
    half = rng.randn(2, n + 1, 3) + 1j * rng.randn(2, n + 1, 3)
    half[:, 0, 2] = half[:, 0, 2].real
    full = np.zeros((2, 2 * n + 1, 3), dtype=complex)
    full[:, n:, 0] = half[:, :, 0]
    full[:, n:, 1] = half[:, :, 1]
    full[:, n:, 2] = half[:, :, 2]
    full[:, :n, 0] = half[:, :0:-1, 1].conj()
    full[:, :n, 1] = half[:, :0:-1, 0].conj()
    full[:, :n, 2] = half[:, :0:-1, 2].conj()
    full[:, n, 1] = full[:, n, 0].conj()
    prim["state0"] = full
    sm = statematrix.StateMatrix(full, kvalue=800.0)
    for label, op in (
        ("T", epgpy.core.T([30.0, 140.0], 25.0)),
        ("E", epgpy.core.E(7.0, 600.0, [40.0, 90.0], 0.03)),
        ("S+1", epgpy.core.S(1)),
        ("S-2", epgpy.core.S(-2)),
        ("D", epgpy.core.D(4.0, 2e-3)),
        ("Dk", epgpy.core.D(4.0, 2e-3, k=1)),
        ("SPOILER", epgpy.core.SPOILER),
    ):
        prim[f"state_{label}"] = op(sm).states
    smn = statematrix.StateMatrix(full, max_nstate=3)
    prim["state_S+1_nmax3"] = epgpy.core.S(1)(smn).states
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **prim)
    print("primitives", len(prim))


if __name__ == "__main__":
    main()
