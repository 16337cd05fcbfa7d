"""Host-side checks of the merged record stream the register kernels execute (epgx_plan_stream; no GPU needed):
whole-TR grouping of forward and derivative FISP tapes (csrc/epgx.cu, csrc/epgx_common.cuh)."""

import collections

import numpy as np
import pytest

import cases
from util import product_namespace

OP_SEG, OP_TR, OP_TRC, OP_TRJ, OP_CONT, OP_NOP = 64, 65, 66, 67, 13, 0


@pytest.fixture(scope="module")
def epg():
    return product_namespace()


def _stream(epg, case, **kw):
    from epgpy_b200 import engine, lowering
    probe = [None, epg.Jacobian(case["jac"])] if case.get("jac") else None
    low = lowering.lower(case["seq"], probe=probe, options=dict(case.get("options", {})), **kw)
    plan = engine.Plan(low)
    return low, plan, plan.stream()


def test_forward_fisp_collapses_to_whole_tr_records(epg):
    low, plan, st = _stream(epg, cases.fisp_unbounded(epg))
    count = collections.Counter(st["code"].tolist())
    assert plan.config()["kernel"] == 2
    assert count[OP_TR] >= low.nadc - 2  # every TR but the first and the last is one record pair


def test_fisp_jacobian_collapses_to_whole_tr_groups(epg):
    case = cases.fisp_jac_global(epg, ntr=40)
    low, plan, st = _stream(epg, case)
    assert plan.config()["kernel"] == 3
    heads = np.flatnonzero(st["code"] == OP_TRJ)
    assert len(heads) == low.nadc - 2  # generic records: the first TR (with the inversion pulse) and the last one (trailing E)
    # groups of five records, aligned to multiples of five inside a 64-record window, never split by a window
    assert np.all(heads % 64 % 5 == 0) and np.all(heads % 64 <= 59)
    for h in heads:
        assert np.all(st["code"][h + 1:h + 5] == OP_CONT)
        grp = st[h:h + 5]
        assert grp[0]["flags"] & 0x182 == 0x182  # PRE | POST | PARTIALS
        assert grp[1]["flags"] == 2  # unit shift +1, no segment flags
        # B1 is injected at the pulse, T1 and T2 before both relaxation operators; one slot per variable
        pre, pulse, post = int(grp[2]["flags"]), int(grp[3]["flags"]), int(grp[4]["flags"])
        assert pre == post and pre | pulse == 7 and pre & pulse == 0 and bin(pulse).count("1") == 1
    # windows holding groups are flagged for the coefficient assembly pass
    for w0 in range(0, len(st), 64):
        has = np.any(st["code"][w0:w0 + 64] == OP_TRJ)
        assert bool(st["flags"][w0] & 0x2000) == bool(has)
    # signal and Jacobian rows ride in CONT'
    rows = [(int(st[h + 1]["aux"]), int(st[h + 1]["rsv1"])) for h in heads]
    assert rows == [(i, i) for i in range(1, low.nadc - 1)]


def test_more_than_three_variables_keep_generic_records(epg):
    case = cases.fisp_jac_pulses(epg)
    low, plan, st = _stream(epg, case)
    assert low.nvar > 3 and not np.any(st["code"] == OP_TRJ)


def test_complex_injections_are_not_grouped(epg):
    case = cases.mse_jac(epg)  # T(150, 0): complex couplings, ring kernel
    low, plan, st = _stream(epg, case)
    assert plan.config()["kernel"] == 0 and not np.any(st["code"] == OP_TRJ)


def test_kernel_choice_for_derivative_tapes(epg):
    from epgpy_b200 import engine, lowering
    # few orders: the orders-over-lanes kernel; many orders and whole-TR groups: one warp per state set
    _, small, _ = _stream(epg, cases.fisp_jac_global(epg, ntr=40))
    assert small.config()["kernel"] == 3
    low, big, st = _stream(epg, cases.fisp_jac_global(epg, ntr=400))
    cfg = big.config()
    assert low.max_order + 1 > 128 and cfg["kernel"] == 4 and cfg["threads_per_cta"] == 128 and cfg["atoms_per_cta"] == 1
    # pure windows: 12 plain groups + 4 NOP records
    pure = [w0 for w0 in range(0, len(st) - 63, 64) if st["flags"][w0] & 0x1000]
    assert len(pure) >= (low.nadc - 2) // 12 - 2
    for w0 in pure:
        assert np.all(st["code"][w0:w0 + 60:5] == OP_TRJ) and np.all(st["flags"][w0 + 1:w0 + 60:5] == 2)
        assert np.all(st["code"][w0 + 60:w0 + 64] == OP_NOP)
    big.set_variant(kernel=4)
    assert big.config()["kernel"] == 3
    big.set_variant(kernel=5)
    assert big.config()["kernel"] == 4


@pytest.mark.timeout(120)
def test_variant_choice_terminates_on_tiny_grids(epg):
    """host-side launch-shape heuristics (small grids get smaller CTAs) for every kernel / lanes / atoms request"""
    from epgpy_b200 import engine, lowering
    for case in (cases.readme_mse(epg), cases.fisp_jac_global(epg, ntr=12), cases.spgr_exchange(epg, ntr=6)):
        probe = [None, epg.Jacobian(case["jac"])] if case.get("jac") else None
        init = epg.StateMatrix(density=case["density"]) if case.get("density") is not None else None
        low = lowering.lower(case["seq"], probe=probe, init=init, options=dict(case.get("options", {})))
        for kernel in range(6):
            for lanes in (0, 1, 2, 32, 64, 128, 256):
                for atoms in (0, 1, 3, 16):
                    plan = engine.Plan(low)
                    try:
                        plan.set_variant(kernel=kernel, lanes_per_atom=lanes, atoms_per_cta=atoms)
                    except (NotImplementedError, MemoryError):
                        continue
                    cfg = plan.config()
                    assert 1 <= cfg["atoms_per_cta"] and cfg["threads_per_cta"] <= 256


def _both_ways(epg, seq, jac=None, options=None, init=None, **kw):
    """the tape through the numpy interpreter, segment by segment and through the merged stream"""
    import tape_interp
    from epgpy_b200 import engine, lowering
    low = lowering.lower(seq, probe=[None, epg.Jacobian(jac)] if jac else None, init=init, options=dict(options or {}), **kw)
    st = engine.Plan(low).stream()
    return low, st, tape_interp.run(low), tape_interp.run(low, stream=st)


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_merged_stream_reproduces_the_tape(name, epg):
    """every golden case: the stream the register kernels execute (segment markers, whole-TR records and groups with
    their fused five-coefficient algebra, hopped E records) gives what the plain tape gives"""
    case = cases.CASES[name](epg)
    init = epg.StateMatrix(density=case["density"]) if case.get("density") is not None else None
    if name in ("gre_lattice_2d", "gre_lattice_3d_cropped", "lattice_jac", "gre_lattice_float", "gre_gradient_time"):
        pytest.skip("lattice tapes run in the shared-memory kernel only: the register kernels' stream is not used")
    if case.get("init") is not None:
        init = np.array(case["init"])
    low, st, (s0, j0), (s1, j1) = _both_ways(epg, case["seq"], case.get("jac"), case.get("options"), init)
    assert np.abs(s1 - s0).max() <= 1e-13 * max(1.0, np.abs(s0).max())
    if low.nvar:
        assert np.abs(j1 - j0).max() <= 1e-12 * max(1.0, np.abs(j0).max())


@pytest.mark.parametrize("variables,max_nstate", [(["B1", "T1", "T2"], None), (["B1", "T1", "T2"], 9), (["T2", "B1"], None),
                                                  (["tau", "T1"], None), (["B1"], 20)])
def test_whole_tr_groups_on_the_host(variables, max_nstate, epg):
    """the derivative train of tests/test_gpu_parity.py (whole-TR groups mixed with TRs that are not): stream == tape"""
    import test_gpu_parity as tg
    options = {"kvalue": 2500.0, **({"max_nstate": max_nstate} if max_nstate else {})}
    low, st, (s0, j0), (s1, j1) = _both_ways(epg, tg._trj_sequence(epg, variables), variables, options, propagate_nondiff=True)
    assert np.count_nonzero(st["code"] == OP_TRJ) >= 60
    assert np.abs(s1 - s0).max() <= 1e-13 * max(1.0, np.abs(s0).max())
    for v in range(low.nvar):
        assert np.abs(j1[:, v] - j0[:, v]).max() <= 1e-12 * max(1e-30, np.abs(j0[:, v]).max())


def test_whole_tr_runs_are_aligned_to_tape_windows(epg):
    """real-valued plans: every run of >= 4 plain whole-TR pairs starts at a window boundary behind a skip mark, windows
    hold an even number of pairs (count in rsv1 of the window's first record, flag 0x8000), the first pair of an odd run
    stays in front of the padding, and every TR keeps its place in the order of the tape (csrc/epgx.cu)"""
    from epgpy_b200 import engine, lowering
    import test_gpu_parity as tgp

    runs = (1, 3, 4, 5, 33, 2, 64, 7, 38)
    low = lowering.lower(tgp._tr_runs_sequence(epg, runs))
    plan = engine.Plan(low)
    st = plan.stream()
    assert plan.config()["kernel"] == 2
    code, flags = st["code"], st["flags"]
    fast = 0
    for w0 in range(0, len(st), 64):
        w = st[w0:w0 + 64]
        if flags[w0] & 0x8000:
            m = int(w["rsv1"][0])
            assert 2 <= m <= 32 and m % 2 == 0
            assert np.all(w["code"][0:2 * m:2] == OP_TR) and np.all(w["code"][1:2 * m:2] == OP_CONT)
            assert np.all(w["code"][2 * m:] == OP_NOP)
            if w0 > 0:  # the window in front ends with padding that opens with a skip mark, or is a full fast window
                prev = st[w0 - 64:w0]
                if not (flags[w0 - 64] & 0x8000 and int(prev["rsv1"][0]) == 32):
                    pad = np.flatnonzero((prev["code"] == OP_NOP) & (prev["aux"] == 1))
                    assert len(pad) and np.all(prev["code"][pad[-1]:] == OP_NOP)
            fast += m
    # the long runs (33, 64, 7, 38 TRs; a run loses its first / last TR to the generic path when the records around it
    # fuse with them) take the fast path, and no run of four plain pairs is left in a generic window
    assert fast >= 32 + 62 + 6 + 36
    streak = 0
    for w0 in range(0, len(st), 64):
        if flags[w0] & 0x8000:
            streak = 0
            continue
        i = w0
        while i < min(w0 + 64, len(st)):
            if code[i] == OP_TR and int(flags[i + 1]) & ~(2 << 2) == 2:
                streak += 1
                i += 2
            else:
                if code[i] != OP_NOP:
                    streak = 0
                i += 1
            assert streak < 4
    # read-out rows are visited in increasing order (no TR was moved across another)
    rows = [int(st[i + 1]["aux"]) for i in np.flatnonzero(code == OP_TR)]
    assert rows == sorted(rows) and len(rows) == len(set(rows))
