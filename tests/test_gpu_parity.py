"""Parity of the CUDA path (through the C ABI) with the reference's golden vectors and the oracle.
Runs on the B200 box: pytest -m gpu"""

import numpy as np
import pytest

import cases
import oracle_api
from util import RTOL32, RTOL64, product_namespace, rel_err, run_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def epg():
    return product_namespace()


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_fp64_matches_reference(name, golden, epg):
    ref = golden(name)
    case = cases.CASES[name](epg)
    sig, jac = run_case(epg.simulate, epg, case)
    assert sig.dtype == np.complex128
    assert rel_err(sig, ref["signal"]) < RTOL64
    if "jacobian" in ref.files:
        assert rel_err(jac, ref["jacobian"]) < RTOL64


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_fp32_matches_reference(name, golden, epg):
    ref = golden(name)
    case = cases.CASES[name](epg)
    sig, jac = run_case(epg.simulate, epg, case, dtype="float32")
    assert sig.dtype == np.complex64
    assert rel_err(sig, ref["signal"]) < RTOL32
    if "jacobian" in ref.files:
        # derivative columns of very different magnitude share one array: compare column-wise
        for i in range(jac.shape[-1]):
            col = ref["jacobian"][..., i]
            if np.abs(col).max() > 1e-9 * np.abs(ref["jacobian"]).max():
                assert rel_err(jac[..., i], col) < 5 * RTOL32
            else:  # a derivative that vanishes by cancellation of O(2 pi tau) terms: absolute check
                assert np.abs(jac[..., i] - col).max() < 1e-3 * np.abs(ref["jacobian"]).max()


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_equal_axes_grid_matches_per_atom_reference_runs(dtype, golden, epg):
    """grid axes of EQUAL size, like the headline 100 x 100 x 100 dictionary: the golden file holds n^3 per-atom scalar
    runs of the unmodified reference (its own vectorised run is wrong there)"""
    ref = golden("fisp_equal_axes")
    case = cases.fisp_equal_axes(epg)
    T1, T2, B1 = case["axes"]
    seq = case["build"](T1, T2[None, :], B1[None, None, :])
    tol = RTOL64 if dtype == "float64" else RTOL32
    sig = epg.simulate(seq, dtype=dtype)  # forward: the real-valued whole-TR kernel of the headline
    assert sig.shape == ref["signal"].shape and rel_err(sig, ref["signal"]) < tol
    sig, jac = epg.simulate(seq, probe=[None, epg.Jacobian(case["jac"])], dtype=dtype)
    assert rel_err(sig, ref["signal"]) < tol
    for i in range(jac.shape[-1]):
        assert rel_err(jac[..., i], ref["jacobian"][..., i]) < (tol if dtype == "float64" else 5 * tol)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(cases.HESSIAN_CASES))
def test_order2_hessian_matches_reference(name, dtype, golden, epg):
    """order-2 forward mode on the device (pair tiles of the shared-memory kernel) against the reference's Hessian probe"""
    ref = golden(name)
    sig, hes = cases.run_hessian(epg, cases.HESSIAN_CASES[name](epg), dtype=dtype)
    tol = RTOL64 if dtype == "float64" else RTOL32
    assert rel_err(sig, ref["signal"]) < tol
    want = ref["hessian"]
    for i in range(want.shape[-2]):
        for j in range(want.shape[-1]):
            col = want[..., i, j]
            if dtype == "float64" or np.abs(col).max() > 1e-9 * np.abs(want).max():
                scale = max(np.abs(col).max(), 1e-300)
                assert np.abs(hes[..., i, j] - col).max() / scale < (tol if dtype == "float64" else 20 * tol) or np.abs(col).max() < 1e-14


def test_probe_expressions(golden, epg):
    ref = golden("probe_expr")
    case = cases.probe_expr(epg)
    vals = epg.simulate(case["seq"], probe=case["probe"])
    assert len(vals) == 3
    for i, v in enumerate(vals):
        assert rel_err(np.asarray(v), ref[f"probe{i}"]) < RTOL64


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(cases.FOURIER_CASES))
def test_fourier_probes(name, dtype, golden, epg):
    """DFT / Imaging probes through the engine: the ring kernel reads every transverse configuration of the lattice
    (EPGX_FLAG_SLOT rows), the host applies the probe's weights (probe.py:168-219, utils.py:12-115)"""
    ref = golden(name)
    case = cases.FOURIER_CASES[name](epg)
    vals = epg.simulate(case["seq"], asarray=False, dtype=dtype, **case["options"])
    assert len(vals) == len(ref.files)
    tol = RTOL64 if dtype == "float64" else RTOL32  # (measured FP32: 1.0e-6 and 2.5e-7)
    for i, v in enumerate(vals):
        want = ref[f"probe{i}"]
        assert np.shape(v) == want.shape
        assert np.abs(np.asarray(v) - want).max() <= tol * max(np.abs(want).max(), 1e-30)


def _run_variant(epg, case, dtype="f64", **variant):
    from epgpy_b200 import engine, functions, lowering

    opts = dict(case.get("options") or {})
    init = epg.StateMatrix(density=case["density"]) if case.get("density") is not None else None
    if case.get("init") is not None:
        init = np.array(case["init"])
    probe = [None, epg.Jacobian(case["jac"])] if case.get("jac") else None
    low = lowering.lower(case["seq"], init=init, probe=probe, options=opts, dtype=dtype)
    plan = engine.Plan(low)
    plan.set_variant(**variant)
    cfg = plan.config()
    parts, _ = functions.run_lowered(low, plan=plan)
    vals = functions._assemble(low, parts)
    return vals, cfg


@pytest.mark.parametrize("lanes,atoms", [(1, 1), (1, 32), (2, 3), (8, 16), (32, 2), (64, 1), (128, 2), (256, 1)])
@pytest.mark.parametrize("name", ["fisp_unbounded", "fisp_bounded", "misc_ops", "spgr_exchange", "fisp_jac_global",
                                  "gre_diffusion_tensor", "init_states", "init_states_jac", "init_states_jac_complex",
                                  "gre_lattice_2d", "gre_lattice_3d_cropped", "lattice_jac"])
def test_ring_kernel_variants(name, lanes, atoms, golden, epg):
    """ring kernel: every lanes-per-atom / atoms-per-CTA mapping gives the same answer (ragged tails included)"""
    ref = golden(name)
    vals, cfg = _run_variant(epg, cases.CASES[name](epg), kernel=1, lanes_per_atom=lanes, atoms_per_cta=atoms)
    assert cfg["lanes_per_atom"] == lanes and cfg["kernel"] == 0
    assert rel_err(vals[0], ref["signal"]) < RTOL64
    if "jacobian" in ref.files:
        assert rel_err(vals[1], ref["jacobian"]) < RTOL64


FORWARD = ["readme_mse", "mse_grid", "fisp_unbounded", "fisp_bounded", "bssfp_offres", "gre_diffusion",
           "gre_diffusion_1d", "hyperecho", "misc_ops", "adc_reduce", "gre_diffusion_tensor", "init_states",
           "init_states_real", "init_states_cropped"]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("lanes,atoms", [(1, 5), (2, 3), (4, 32), (8, 16), (16, 1), (32, 2), (64, 1), (64, 3), (128, 2), (256, 1)])
@pytest.mark.parametrize("name", FORWARD)
def test_reg_kernel_variants(name, lanes, atoms, dtype, golden, epg):
    """register kernel: sub-warp groups, one warp and several warps per atom, every slot count"""
    ref = golden(name)
    try:
        vals, cfg = _run_variant(epg, cases.CASES[name](epg), dtype=dtype, kernel=2, lanes_per_atom=lanes, atoms_per_cta=atoms)
    except MemoryError:
        pytest.skip("more orders than lanes x slots of any instance")
    assert cfg["kernel"] == 1 and cfg["lanes_per_atom"] == lanes
    assert rel_err(vals[0], ref["signal"]) < (RTOL64 if dtype == "f64" else RTOL32)


def test_reg_and_ring_agree_on_large_orders(epg):
    """400-TR FISP (200 live orders): register kernel (1, 2, 4 warps per atom) vs ring kernel"""
    case = cases.fisp(epg, 400, sizes=(3, 2, 2))
    ring, _ = _run_variant(epg, case, kernel=1)
    for lanes in (32, 64, 128):
        reg, cfg = _run_variant(epg, case, kernel=2, lanes_per_atom=lanes)
        assert cfg["kernel"] == 1
        assert rel_err(reg[0], ring[0]) < 1e-12
    bounded = cases.fisp(epg, 400, sizes=(3, 2, 2), max_nstate=37)
    ring, _ = _run_variant(epg, bounded, kernel=1)
    for lanes in (4, 8, 32):
        reg, _ = _run_variant(epg, bounded, kernel=2, lanes_per_atom=lanes)
        assert rel_err(reg[0], ring[0]) < 1e-12


def _real_sequence(epg, ntr=70):
    """+-90 degree pulses, relaxation, diffusion, spoiler, PD, shifts of both signs: a real-valued graph"""
    T1 = np.linspace(300, 2000, 5)
    T2 = np.linspace(30, 200, 4)[None, :]
    B1 = np.array([0.8, 1.0, 1.15])[None, None, :]
    seq = [epg.PD([1.0, 0.5, 2.0, 1.5, 0.7]), epg.T(180, 90), epg.E(15, T1, T2)]
    for i in range(ntr):
        ph = 90 if i % 3 else 270
        seq += [epg.T((10 + i % 40) * B1, ph), epg.E(2.5, T1, T2), epg.Adc(phase=-ph + 90.0), epg.E(6 + (i % 5), T1, T2),
                epg.S(1 if i % 11 else -1)]
        if i % 7 == 3:
            seq += [epg.D(4.0, 1.2e-3, k=1 if i % 11 else -1)]
        if i == 40:
            seq += [epg.SPOILER]
    return seq


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("lanes,atoms", [(0, 0), (1, 7), (2, 16), (4, 32), (8, 3), (16, 8), (32, 4)])
@pytest.mark.parametrize("max_nstate", [None, 9])
def test_real_kernel_variants(lanes, atoms, dtype, max_nstate, epg):
    """the real-valued register kernel (three reals per order) against the ring kernel and the oracle"""
    case = {"seq": _real_sequence(epg), "options": {"kvalue": 2500.0, **({"max_nstate": max_nstate} if max_nstate else {})}}
    ring, _ = _run_variant(epg, case, kernel=1)
    try:
        got, cfg = _run_variant(epg, case, dtype=dtype, kernel=3, lanes_per_atom=lanes, atoms_per_cta=atoms)
    except MemoryError:
        pytest.skip("more orders than lanes x slots of any instance")
    assert cfg["kernel"] == 2
    assert rel_err(got[0], ring[0]) < (1e-12 if dtype == "f64" else RTOL32)
    ref = oracle_api.O.simulate(_real_sequence(oracle_api.epg), kvalue=2500.0, max_nstate=max_nstate)
    assert rel_err(ring[0], ref) < RTOL64


def _tr_runs_sequence(epg, runs, zero_flip_at=(), sizes=(3, 2, 5)):
    """FISP-like train cut into RUNS of whole-TR records by records that are not TRs (a spoiled inversion, a pause
    without read-out, a reset): the stream builder aligns every run of >= 4 plain TR pairs to a tape window, keeps the
    first pair of an odd run generic and pads partial windows (csrc/epgx.cu).  zero_flip_at: TR indices with a 0 degree
    pulse (u = 0: the window must run unscaled, csrc/epgx_real.cuh)"""
    T1 = np.linspace(300, 3000, sizes[0])
    T2 = np.linspace(20, 300, sizes[1])[None, :]
    B1 = np.linspace(0.7, 1.2, sizes[2])[None, None, :]
    rng = np.random.RandomState(3)
    seq, i = [epg.T(180, 0), epg.E(20, T1, T2)], 0
    for r, n in enumerate(runs):
        for _ in range(n):
            fa = 0.0 if i in zero_flip_at else 10 + 50 * abs(np.sin(i * np.pi / 37))
            tr = rng.uniform(11, 16)
            seq.append([epg.T(fa * B1, 90), epg.E(3, T1, T2), epg.ADC, epg.E(tr - 3, T1, T2), epg.S(1)])
            i += 1
        if r % 3 == 0:
            seq += [epg.E(40, T1, T2), epg.SPOILER, epg.T(180, 0), epg.E(15, T1, T2)]
        elif r % 3 == 1:
            seq += [epg.T(30 * B1, 90), epg.E(5, T1, T2), epg.S(1)]  # a TR without read-out: not a whole-TR record
        else:
            seq += [epg.RESET, epg.T(90 * B1, 90), epg.E(4, T1, T2)]
    return seq


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("lanes", [0, 2, 8, 32])
@pytest.mark.parametrize("runs,zeros,max_nstate", [((1, 3, 4, 5, 33, 2, 64, 7, 38), (), None), ((70, 9), (12, 13, 50, 75), None),
                                                   ((5, 67, 4), (40,), 6), ((32,), (), None), ((31,), (), None)])
def test_real_kernel_whole_tr_runs(runs, zeros, max_nstate, lanes, dtype, epg):
    """runs of whole-TR records of every length class (below the run threshold, odd, one window, several windows, partial
    last window), separated by generic records, zero-flip pulses inside scaled windows, bounded and unbounded states:
    real kernel (all lane counts) against the ring kernel and the oracle"""
    opts = {"max_nstate": max_nstate} if max_nstate else {}
    case = {"seq": _tr_runs_sequence(epg, runs, zeros), "options": opts}
    ring, _ = _run_variant(epg, case, kernel=1)
    try:
        got, cfg = _run_variant(epg, case, dtype=dtype, kernel=3, lanes_per_atom=lanes)
    except MemoryError:
        pytest.skip("more orders than lanes x slots of any instance")
    assert cfg["kernel"] == 2
    assert rel_err(got[0], ring[0]) < (1e-12 if dtype == "f64" else RTOL32)
    ref = oracle_api.O.simulate(_tr_runs_sequence(oracle_api.epg, runs, zeros), **opts)
    assert rel_err(ring[0], ref) < RTOL64
    assert rel_err(got[0], ref) < (RTOL64 if dtype == "f64" else RTOL32)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("lanes", [0, 1, 4, 32])
@pytest.mark.parametrize("name", ["init_states_real", "init_states_cropped"])
def test_real_kernel_initial_states(name, lanes, dtype, golden, epg):
    """initial state with populated orders n > 0 in the real-valued register kernel"""
    ref = golden(name)
    got, cfg = _run_variant(epg, cases.CASES[name](epg), dtype=dtype, kernel=3, lanes_per_atom=lanes)
    assert cfg["kernel"] == 2
    assert rel_err(got[0], ref["signal"]) < (RTOL64 if dtype == "f64" else RTOL32)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("kernel,lanes", [(4, 0), (4, 2), (4, 32), (4, 64), (5, 0)])
def test_realjac_kernels_initial_states(kernel, lanes, dtype, golden, epg):
    """initial state with populated orders in the real-valued derivative kernels (orders over warps, warp per state set)"""
    ref = golden("init_states_jac")
    got, cfg = _run_variant(epg, cases.init_states_jac(epg), dtype=dtype, kernel=kernel, lanes_per_atom=lanes)
    assert cfg["kernel"] == kernel - 1
    tol = RTOL64 if dtype == "f64" else RTOL32
    assert rel_err(got[0], ref["signal"]) < tol
    for i in range(ref["jacobian"].shape[-1]):
        assert rel_err(got[1][..., i], ref["jacobian"][..., i]) < (tol if dtype == "f64" else 5 * tol)


def _real_jac_sequence(epg, ntr=50):
    """real-valued graph with derivatives w.r.t. B1 (per-atom coefficient), T1, T2 and one pulse angle"""
    T1 = np.linspace(300, 2000, 4)
    T2 = np.linspace(30, 200, 3)[None, :]
    B1 = np.array([0.8, 1.0, 1.15])[None, None, :]
    o1 = ["T1", "T2"]
    seq = [epg.T(180, 90), epg.E(15, T1, T2, order1=o1)]
    for i in range(ntr):
        ph = 90 if i % 3 else 270
        fa = 10.0 + i % 40
        od = {"B1": {"alpha": fa}}
        if i == 7:
            od["a7"] = {"alpha": B1}
        seq += [epg.T(fa * B1, ph, order1=od), epg.E(2.5, T1, T2, order1=o1), epg.Adc(phase=-ph + 90.0),
                epg.E(6 + (i % 5), T1, T2, order1={"T1": {"T1": 1}, "T2": {"T2": 1}, "tau": {"tau": 0.5}}), epg.S(1 if i % 11 else -1)]
        if i % 7 == 3:
            seq += [epg.D(4.0, 1.2e-3, k=1 if i % 11 else -1)]
        if i == 30:
            seq += [epg.SPOILER]
    return seq


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("lanes,atoms", [(0, 0), (1, 5), (2, 16), (8, 3), (32, 4), (64, 1), (64, 3), (128, 2), (256, 1)])
@pytest.mark.parametrize("max_nstate", [None, 9])
def test_realjac_kernel_variants(lanes, atoms, dtype, max_nstate, epg):
    """real-valued register kernel with resident partial states (sub-warp, one warp, several warps per atom)
    against the ring kernel and the oracle, D / SPOILER propagated to the partials"""
    variables = ["B1", "T1", "T2", "tau", "a7"]  # 5 variables -> two tiles of three
    case = {"seq": _real_jac_sequence(epg), "jac": variables,
            "options": {"kvalue": 2500.0, **({"max_nstate": max_nstate} if max_nstate else {})}}
    from epgpy_b200 import engine, functions, lowering

    def run(kernel, dt, **variant):
        low = lowering.lower(case["seq"], probe=[None, epg.Jacobian(variables)], options=dict(case["options"]), dtype=dt,
                             propagate_nondiff=True)
        plan = engine.Plan(low)
        plan.set_variant(kernel=kernel, **variant)
        parts, _ = functions.run_lowered(low, plan=plan)
        return functions._assemble(low, parts), plan.config()

    ring, _ = run(1, "f64")
    try:
        got, cfg = run(4, dtype, lanes_per_atom=lanes, atoms_per_cta=atoms)
    except MemoryError:
        pytest.skip("more orders than lanes x slots of any instance")
    assert cfg["kernel"] == 3 and cfg["var_tiles"] == 2
    tol = 1e-12 if dtype == "f64" else RTOL32
    assert rel_err(got[0], ring[0]) < tol
    for i in range(len(variables)):
        assert rel_err(got[1][..., i], ring[1][..., i]) < (tol if dtype == "f64" else 5 * RTOL32)
    rs, rj = oracle_api.O.simulate(_real_jac_sequence(oracle_api.epg), jacobian=variables, kvalue=2500.0,
                                   max_nstate=max_nstate, propagate_nondiff=True)
    assert rel_err(ring[0], rs) < RTOL64 and rel_err(ring[1], rj) < RTOL64


def _trj_sequence(epg, variables, ntr=70):
    """FISP-like train whose TRs collapse to whole-TR derivative groups (EPGX_OP_TRJ), with a few TRs that do not
    (phase-compensated ADC, diffusion, spoiler, negative shift, a proton-density change in between)"""
    T1 = np.linspace(300, 2000, 4)
    T2 = np.linspace(30, 200, 3)[None, :]
    B1 = np.array([0.8, 1.0, 1.15])[None, None, :]
    o1 = {v: {v: 1} for v in ("T1", "T2") if v in variables}
    o2 = dict(o1, **({"tau": {"tau": 0.5}} if "tau" in variables else {}))
    seq = [epg.T(180, 90), epg.E(15, T1, T2, order1=o1)]
    for i in range(ntr):
        ph = 90 if i % 3 else 270
        fa = 10.0 + i % 40
        od = {"B1": {"alpha": fa}} if "B1" in variables else {}
        adc = epg.Adc(phase=40.0) if i == 17 else epg.ADC
        seq += [epg.T(fa * B1, ph, order1=od), epg.E(2.5, T1, T2, order1=o1), adc, epg.E(6 + (i % 5), T1, T2, order1=o2),
                epg.S(-1 if i == 29 else 1)]
        if i == 23:
            seq += [epg.D(4.0, 1.2e-3, k=1)]
        if i == 37:
            seq += [epg.PD(0.7)]
        if i == 44:
            seq += [epg.SPOILER]
    return seq


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("lanes,atoms", [(0, 0), (2, 16), (8, 3), (32, 4), (64, 3), (128, 2), (256, 1)])
@pytest.mark.parametrize("variables,max_nstate", [(["B1", "T1", "T2"], None), (["B1", "T1", "T2"], 9), (["T2", "B1"], None),
                                                  (["tau", "T1"], None), (["B1"], 20)])
def test_realjac_whole_tr_groups(lanes, atoms, dtype, variables, max_nstate, epg):
    """derivative tapes of at most three variables: TRs of the form [inj] E [inj] T [inj] E ADC + shift run as fused
    five-coefficient groups in the real-valued derivative kernel; against the ring kernel and the oracle"""
    from epgpy_b200 import engine, functions, lowering
    options = {"kvalue": 2500.0, **({"max_nstate": max_nstate} if max_nstate else {})}

    def run(kernel, dt, **variant):
        low = lowering.lower(_trj_sequence(epg, variables), probe=[None, epg.Jacobian(variables)], options=dict(options), dtype=dt,
                             propagate_nondiff=True)
        plan = engine.Plan(low)
        plan.set_variant(kernel=kernel, **variant)
        parts, _ = functions.run_lowered(low, plan=plan)
        return functions._assemble(low, parts), plan.config()

    ring, _ = run(1, "f64")
    try:
        got, cfg = run(4, dtype, lanes_per_atom=lanes, atoms_per_cta=atoms)
    except MemoryError:
        pytest.skip("more orders than lanes x slots of any instance")
    assert cfg["kernel"] == 3 and cfg["var_tiles"] == 1
    if lanes == 0:  # the same tape through the warp-per-state-set kernel
        alt, acfg = run(5, dtype)
        assert acfg["kernel"] == 4 and acfg["threads_per_cta"] == 32 * (1 + len(variables))
        assert rel_err(alt[0], got[0]) < (1e-11 if dtype == "f64" else RTOL32)
        assert rel_err(alt[1], got[1]) < (1e-11 if dtype == "f64" else 5 * RTOL32)
    tol = 1e-11 if dtype == "f64" else RTOL32
    assert rel_err(got[0], ring[0]) < tol
    for i in range(len(variables)):
        assert rel_err(got[1][..., i], ring[1][..., i]) < (tol if dtype == "f64" else 5 * RTOL32)
    rs, rj = oracle_api.O.simulate(_trj_sequence(oracle_api.epg, variables), jacobian=variables, kvalue=2500.0,
                                   max_nstate=max_nstate, propagate_nondiff=True)
    assert rel_err(ring[0], rs) < RTOL64 and rel_err(ring[1], rj) < RTOL64


def test_realjac_kernel_is_chosen_for_the_fisp_jacobian(golden, epg):
    from epgpy_b200 import engine, lowering

    case = cases.fisp_jac_global(epg)
    low = lowering.lower(case["seq"], probe=[None, epg.Jacobian(case["jac"])])
    assert engine.Plan(low).config()["kernel"] == 3
    case = cases.mse_jac(epg)  # T(150, 0): complex couplings -> ring kernel
    low = lowering.lower(case["seq"], probe=[None, epg.Jacobian(case["jac"])])
    assert engine.Plan(low).config()["kernel"] == 0


def test_real_kernel_is_chosen_for_fisp_and_refused_otherwise(epg):
    from epgpy_b200 import engine, lowering

    plan = engine.Plan(lowering.lower(cases.fisp_unbounded(epg)["seq"]))
    assert plan.config()["kernel"] == 2
    plan = engine.Plan(lowering.lower(cases.readme_mse(epg)["seq"]))  # T(120, 0) couples real and imaginary parts
    assert plan.config()["kernel"] == 1
    with pytest.raises(NotImplementedError):
        plan.set_variant(kernel=3)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("diffusion,phase,spoil", [(True, True, 117.0), (False, True, 50.0), (True, False, 0.0), (False, False, 90.0)])
def test_complex_whole_tr_windows(diffusion, phase, spoil, dtype, epg):
    """200-TR gradient-echo trains ([D] . E.T.E . ADC(phase) . S): several pure whole-TR windows of the complex
    register kernel, against the ring kernel (FP64) and the oracle"""
    def build(e, ntr=200):
        T1 = np.array([600.0, 1200.0, 1500.0])
        T2 = np.array([[40.0, 80.0]])
        seq = []
        for n in range(ntr):
            ph = spoil * n * (n + 1) / 2
            seq += [e.T(12 + n % 30, ph), e.E(2, T1, T2), e.Adc(phase=-ph) if phase else e.ADC, e.E(7, T1, T2), e.S(1)]
            if diffusion:
                seq += [e.D(9.0, 1.5e-3, k=1)]
        return seq

    case = {"seq": build(epg), "options": {"kvalue": 800.0}}
    ring, _ = _run_variant(epg, case, kernel=1)
    got, cfg = _run_variant(epg, case, dtype=dtype, kernel=2, lanes_per_atom=32)
    assert cfg["kernel"] == 1
    assert rel_err(got[0], ring[0]) < (1e-12 if dtype == "f64" else RTOL32)
    ref = oracle_api.O.simulate(build(oracle_api.epg), kvalue=800.0)
    assert rel_err(ring[0], ref) < RTOL64


@pytest.mark.parametrize("vars_per_pass", [1, 3])
@pytest.mark.parametrize("name", ["fisp_jac_pulses", "jac_all_params", "mse_jac"])
def test_variable_tiling(name, vars_per_pass, golden, epg):
    ref = golden(name)
    vals, cfg = _run_variant(epg, cases.CASES[name](epg), kernel=1, vars_per_pass=vars_per_pass)
    assert cfg["vars_per_pass"] == vars_per_pass and cfg["kernel"] == 0
    assert rel_err(vals[0], ref["signal"]) < RTOL64 and rel_err(vals[1], ref["jacobian"]) < RTOL64


def test_c_abi_host_call(golden, epg):
    """epgx_simulate_host: plain pointers in, plain pointers out; atom sub-ranges"""
    from epgpy_b200 import engine, lowering

    ref = golden("fisp_bounded")["signal"]
    low = lowering.lower(cases.fisp_bounded(epg)["seq"], options={"max_nstate": 10})
    plan = engine.Plan(low)
    out = np.zeros((low.nadc, low.natoms, 1), dtype=np.complex128)
    plan.run_host(0, 0, low.natoms, out)
    assert rel_err(out.reshape(ref.shape), ref) < RTOL64
    # a ragged sub-range of atoms
    part = np.zeros((low.nadc, 7, 1), dtype=np.complex128)
    plan.run_host(0, 5, 7, part)
    assert np.array_equal(part, out[:, 5:12])


def test_partials_through_nondiff_ops_gpu(epg):
    """D / X / SPOILER applied to the partial states (propagate_nondiff=True) vs the oracle"""
    def seq(e):
        T2 = np.array([40.0, 80.0])
        out = [e.T(90, 90)]
        for i in range(6):
            out += [e.S(1), e.D(3.0, 1.5e-3, k=1), e.E(3, 900.0, T2), e.T(35, 10.0 * i, order1={"a": "alpha"}),
                    e.SPOILER if i == 3 else e.NULL if hasattr(e, "NULL") else e.Wait(0), e.ADC]
        return out

    sig, jac = epg.simulate(seq(epg), probe=[None, epg.Jacobian(["a"])], kvalue=3000.0, propagate_nondiff=True)
    rs, rj = oracle_api.O.simulate(seq(oracle_api.epg), jacobian=["a"], kvalue=3000.0, propagate_nondiff=True)
    assert rel_err(sig, rs) < RTOL64 and rel_err(jac, rj) < RTOL64


@pytest.mark.parametrize("vars_per_pass", [0, 1])
def test_mt_bssfp_pulse_jacobian_gpu(vars_per_pass, epg):
    """BASELINE configs[4]: two-pool MT bSSFP + per-pulse flip-angle Jacobian (ring kernel, exchange, tiled variables)
    against central finite differences of the oracle's forward signal and the forward signal itself"""
    from epgpy_b200 import engine, functions, lowering
    from test_lowering_cpu import _mt_jacobian_by_finite_differences
    ntr, noff = 12, 5
    case = cases.bssfp_mt_pulse_jac(epg, ntr, noff)
    low = lowering.lower(case["seq"], probe=[None, epg.Jacobian(case["jac"])], init=epg.StateMatrix(density=case["density"]),
                         propagate_nondiff=True)
    plan = engine.Plan(low)
    if vars_per_pass:
        plan.set_variant(kernel=1, vars_per_pass=vars_per_pass)
    parts, _ = functions.run_lowered(low, plan=plan)
    sig, jac = functions._assemble(low, parts)
    ref = oracle_api.O.simulate(cases.bssfp_mt_pulse_jac(oracle_api.epg, ntr, noff, diff=False)["seq"], density=case["density"])
    assert rel_err(sig, np.asarray(ref)) < RTOL64
    assert rel_err(jac.sum(axis=1), _mt_jacobian_by_finite_differences(ntr, noff)) < 1e-7


def test_larger_grid_vs_oracle(epg):
    """FISP at a size the oracle still finishes in seconds: 12 x 10 x 8 atoms, 120 TRs, unbounded"""
    case = cases.fisp(epg, 120, sizes=(12, 10, 8))
    sig = np.asarray(epg.simulate(case["seq"]))
    ref, _ = oracle_api.run(cases.fisp(oracle_api.epg, 120, sizes=(12, 10, 8)))
    assert rel_err(sig, ref) < RTOL64
    sig32 = np.asarray(epg.simulate(case["seq"], dtype="float32"))
    assert rel_err(sig32, ref) < RTOL32


def test_linearity_and_density_scaling(epg):
    """size-independent properties at a large grid (50k atoms): the signal is linear in the proton
    density, and PD(2) == 2 x PD(1); pruning of unobservable orders does not change any sample"""
    from epgpy_b200 import engine, functions, lowering

    T1 = np.linspace(300, 3000, 50)
    T2 = np.linspace(20, 300, 40)[None, :]
    B1 = np.linspace(0.7, 1.2, 25)[None, None, :]
    fa, tr = cases._fisp_schedule(100)

    def seq(pd):
        s = [epg.PD(pd), epg.T(180, 0), epg.E(20, T1, T2)]
        for i in range(100):
            s += [epg.T(fa[i] * B1, 90), epg.E(3, T1, T2), epg.ADC, epg.E(tr[i] - 3, T1, T2), epg.S(1)]
        return s

    a = epg.simulate(seq(1.0), max_nstate=16)
    b = epg.simulate(seq(2.5), max_nstate=16)
    assert a.shape == (100, 50, 40, 25)
    assert rel_err(b, 2.5 * a) < 1e-13
    low = lowering.lower(seq(1.0), options={"max_nstate": 16}, prune_unobservable=False)
    parts, _ = functions.run_lowered(low)
    c = functions._assemble(low, parts)[0]
    assert rel_err(c, a) < 1e-14


def test_simulate_over_several_devices_of_one_process(golden, epg):
    """simulate(device=[0, 1]): contiguous atom slabs on two devices of one process, no inter-device traffic"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ref = golden("mse_grid")["signal"]
    sig = epg.simulate(cases.mse_grid(epg)["seq"], device=[0, 1])
    assert rel_err(sig, ref) < RTOL64
    case = cases.fisp_jac_global(epg)
    sig, jac = epg.simulate(case["seq"], probe=[None, epg.Jacobian(case["jac"])], device=[1, 0])
    assert rel_err(sig, golden("fisp_jac_global")["signal"]) < RTOL64
    assert rel_err(jac, golden("fisp_jac_global")["jacobian"]) < RTOL64


# ------------------------------------------------------------------------------------------------ #
# final-state read-back: op(sm), simulate(init=previous state)
# ------------------------------------------------------------------------------------------------ #


def test_operator_call_returns_the_reference_state(golden, epg):
    """`op(sm)` (epgpy/operator.py:96-104) through epgx_simulate_state against the reference's own states
    (primitives.npz: state_<op> = op(StateMatrix(state0, kvalue=800)).states)"""
    p = golden("primitives")
    sm = epg.StateMatrix(p["state0"], kvalue=800.0)
    ops = {"T": epg.T([30.0, 140.0], 25.0), "E": epg.E(7.0, 600.0, [40.0, 90.0], 0.03), "S+1": epg.S(1), "S-2": epg.S(-2),
           "D": epg.D(4.0, 2e-3), "Dk": epg.D(4.0, 2e-3, k=1), "SPOILER": epg.SPOILER}
    for label, op in ops.items():
        out = op(sm)
        want = p[f"state_{label}"]
        assert out.states.shape == want.shape, label
        assert rel_err(out.states, want) < RTOL64, label
        assert np.array_equal(sm.states, p["state0"]), "the input state matrix must not change"
    smn = epg.StateMatrix(p["state0"], max_nstate=3)
    assert rel_err(epg.S(1)(smn).states, p["state_S+1_nmax3"]) < RTOL64
    # in place
    sm2 = epg.StateMatrix(p["state0"], kvalue=800.0)
    assert ops["T"](sm2, inplace=True) is sm2 and rel_err(sm2.states, p["state_T"]) < RTOL64


@pytest.mark.parametrize("name,cut", [("fisp_unbounded", 5 * 23 + 2), ("misc_ops", 9), ("spgr_exchange", 4 * 11), ("three_pool_exchange", 4 * 7)])
def test_simulate_resumes_from_a_read_back_state(name, cut, golden, epg):
    """run the head of a sequence with functions.apply_operators, read the state back, resume with
    simulate(init=state) (the reference's 'resumable by hand' use, functions.py:133-144): the tail's read-outs equal
    those of the one-piece run"""
    from epgpy_b200 import functions
    from epgpy_b200.lowering import flatten_sequence

    case = cases.CASES[name](epg)
    seq = flatten_sequence(case["seq"])
    opts = dict(case.get("options") or {})
    init = epg.StateMatrix(density=case["density"], **opts) if case.get("density") is not None else epg.StateMatrix([0, 0, 1], **opts)
    head, tail = seq[:cut], seq[cut:]
    nhead = sum(isinstance(op, type(epg.ADC)) for op in head)
    whole = golden(name)["signal"]
    # broadcast the initial state to the grid of the whole sequence, as simulate does (functions.py:136-141)
    grid = epg.getshape(seq)
    sm0 = epg.StateMatrix(init.states, density=init.density, shape=grid, **opts)
    mid = functions.apply_operators([op for op in head if not isinstance(op, type(epg.ADC))], sm0)
    assert mid.shape == tuple(grid)
    got = epg.simulate(tail, init=mid)
    assert rel_err(np.asarray(got), whole[nhead:]) < RTOL64


# ------------------------------------------------------------------------------------------------ #
# per-pulse variables: one thread per state set (epgx_pulsejac.cuh)
# ------------------------------------------------------------------------------------------------ #


def _pulse_sequence(epg, ntr, every=1, extras=True):
    """FISP-like train with one flip-angle variable per pulse (examples/differentiation/optim_mrf.py:96-149), plus a few
    TRs that are not of the plain form (negative shift, diffusion, spoiler, phase-compensated ADC, PD change)"""
    fa, tr = cases._fisp_schedule(ntr)
    T1 = np.linspace(300, 3000, 3)
    T2 = np.linspace(20, 300, 4)[None, :]
    names = []
    seq = [epg.T(180, 90), epg.E(20, T1, T2)]
    for i in range(ntr):
        kw = {}
        if i % every == 0:
            names.append(f"a{i:04d}")
            kw = dict(order1={names[-1]: "alpha"})
        adc = epg.Adc(phase=30.0) if (extras and i == 9) else epg.ADC
        seq += [epg.T(fa[i], 90 if i % 3 else 270, **kw), epg.E(3, T1, T2), adc, epg.E(tr[i] - 3, T1, T2),
                epg.S(-1 if (extras and i == 13) else 1)]
        if extras and i == 17:
            seq += [epg.D(4.0, 1.2e-3, k=1)]
        if extras and i == 21:
            seq += [epg.SPOILER]
        if extras and i == 25:
            seq += [epg.PD(0.8, reset=False)]
    return seq, names


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("ntr,max_nstate,pre_inject", [(40, 10, True), (40, 3, True), (40, 15, False), (150, 7, True)])
def test_pulsejac_kernel(ntr, max_nstate, pre_inject, dtype, epg):
    """thread-per-state-set kernel against the shared-memory kernel and the oracle: signal + Jacobian with one variable
    per pulse, D / SPOILER / negative shifts / PD in between, with and without derivative pre-injection"""
    from epgpy_b200 import engine, functions, lowering

    seq, names = _pulse_sequence(epg, ntr)
    variables = ["magnitude"] + names

    def run(kernel, dt):
        low = lowering.lower(seq, probe=[None, epg.Jacobian(variables)], options={"max_nstate": max_nstate, "kvalue": 2500.0}, dtype=dt,
                             propagate_nondiff=True, pre_inject=pre_inject)
        plan = engine.Plan(low)
        if kernel:
            plan.set_variant(kernel=kernel)
        res, _ = functions.run_lowered(low, plan=plan)
        return functions._assemble(low, res), plan.config()

    ring, _ = run(1, "f64")
    got, cfg = run(0, dtype)  # the automatic choice
    assert cfg["kernel"] == 5 and cfg["var_tiles"] == (len(names) + 1 + 127) // 128
    tol = 1e-12 if dtype == "f64" else RTOL32
    assert rel_err(got[0], ring[0]) < tol
    for i in range(len(variables)):
        col = ring[1][..., i]
        assert np.abs(got[1][..., i] - col).max() <= (tol if dtype == "f64" else 5 * tol) * max(np.abs(ring[1]).max(), 1e-300)
    oseq, _ = _pulse_sequence(oracle_api.epg, ntr)
    rs, rj = oracle_api.O.simulate(oseq, jacobian=variables, kvalue=2500.0, max_nstate=max_nstate, propagate_nondiff=True)
    assert rel_err(ring[0], rs) < RTOL64 and rel_err(ring[1], rj) < RTOL64


def test_pulsejac_kernel_is_refused_for_repeated_variables(epg):
    from epgpy_b200 import engine, lowering

    case = cases.fisp_jac_global(epg)  # B1, T1, T2: injected at every operator
    plan = engine.Plan(lowering.lower(case["seq"], probe=[None, epg.Jacobian(case["jac"])], options={"max_nstate": 8}))
    assert plan.config()["kernel"] != 5
    with pytest.raises(NotImplementedError):
        plan.set_variant(kernel=6)


@pytest.mark.parametrize("name", ["fisp_bounded", "gre_diffusion_1d", "spgr_exchange", "mse_jac", "misc_ops"])
def test_lattice_path_reproduces_the_1d_path(name, golden, epg):
    """the general lattice machinery (gather maps, slot of k = 0, per-slot diffusion tables) forced on 1-d sequences must
    give the 1-d answers: golden vectors of the reference"""
    ref = golden(name)
    case = cases.CASES[name](epg)
    sig, jac = run_case(epg.simulate, epg, case, lattice=True)
    assert rel_err(sig, ref["signal"]) < RTOL64
    if "jacobian" in ref.files:
        assert rel_err(jac, ref["jacobian"]) < RTOL64


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_real_valued_signal_rows_path(dtype, epg):
    """real-valued plans ship only the real parts over PCIe and widen them to complex on the host
    (epgx_simulate_real + epgx_expand_real): bit-identical to the complex rows of the plain path, ragged chunks and a
    column offset included"""
    from epgpy_b200 import engine, functions, lowering

    case = cases.fisp(epg, 200, sizes=(45, 41, 31))  # 57 k atoms x 200 rows x 8 B (FP32) = 91 MB: above the small-result cut
    low = lowering.lower(case["seq"], dtype=dtype)
    plan = engine.Plan(low)
    assert plan.real_signal()
    a, _ = functions.run_lowered(low, plan=plan, real_output=False)
    b, _ = functions.run_lowered(low, plan=plan)
    assert b.parts[0][3] is None, "the real-rows path was not taken"
    assert np.array_equal(a.sig_host, b.sig_host) and not np.any(b.sig_host.imag)
    c, _ = functions.run_lowered(low, plan=plan, atom_range=(1234, 30011), nchunk=7)
    assert np.array_equal(c.sig_host, a.sig_host[:, 1234:1234 + 30011])
    # a complex plan refuses
    low2 = lowering.lower(cases.readme_mse(epg)["seq"], dtype=dtype)
    assert not engine.Plan(low2).real_signal()
