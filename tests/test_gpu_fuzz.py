"""Random sequences on the GPU: the automatically chosen kernel, every forced kernel that accepts the tape, both
precisions -- against the oracle."""

import numpy as np
import pytest

import fuzz
import oracle_api
from util import product_namespace, rel_err

pytestmark = pytest.mark.gpu


def _run(epg, seq, opts, jac, dtype, kernel, lanes=0):
    from epgpy_b200 import engine, functions, lowering

    low = lowering.lower(seq, probe=[None, epg.Jacobian(jac)] if jac else None, options=dict(opts), dtype=dtype,
                         propagate_nondiff=True)
    plan = engine.Plan(low)
    plan.set_variant(kernel=kernel, lanes_per_atom=lanes)
    parts, _ = functions.run_lowered(low, plan=plan)
    return functions._assemble(low, parts), plan.config()["kernel"]


@pytest.mark.parametrize("seed", range(40))
def test_random_forward(seed):
    epg = product_namespace()
    real_only = seed % 3 == 0
    seq, opts, _ = fuzz.random_case(epg, seed, real_only=real_only)
    ref_seq, _, _ = fuzz.random_case(oracle_api.epg, seed, real_only=real_only)
    ref = oracle_api.O.simulate(ref_seq, kvalue=opts["kvalue"], max_nstate=opts.get("max_nstate"))
    scale = max(1.0, np.abs(ref).max())
    used = set()
    for kernel in (0, 1, 2, 3):  # auto, ring, reg, real
        for dtype, tol in (("f64", 1e-11), ("f32", 2e-5)):
            for lanes in (0, 1, 32, 64):
                try:
                    got, k = _run(epg, seq, opts, None, dtype, kernel, lanes)
                except (NotImplementedError, MemoryError):
                    continue
                used.add(k)
                assert np.abs(got[0] - ref).max() < tol * scale, (kernel, dtype, lanes)
    assert 0 in used and 1 in used and (2 in used) == real_only_eligible(seq, epg)


def real_only_eligible(seq, epg):
    from epgpy_b200 import engine, lowering

    return engine.Plan(lowering.lower(seq)).config()["kernel"] == 2


@pytest.mark.parametrize("seed", range(100, 125))
def test_random_jacobian(seed):
    epg = product_namespace()
    real_only = seed % 2 == 0
    seq, opts, jac = fuzz.random_case(epg, seed, real_only=real_only, with_jac=True)
    ref_seq, _, _ = fuzz.random_case(oracle_api.epg, seed, real_only=real_only, with_jac=True)
    rs, rj = oracle_api.O.simulate(ref_seq, kvalue=opts["kvalue"], max_nstate=opts.get("max_nstate"), jacobian=jac,
                                   propagate_nondiff=True)
    ss, sj = max(1.0, np.abs(rs).max()), max(1.0, np.abs(rj).max())
    for kernel in (0, 1, 4, 5):  # auto, ring, realjac (orders over warps), setjac (one warp per state set)
        # FP32 bar of the derivative columns: 5e-5 = 5 x the 1e-5 signal bar, as in test_fp32_matches_reference (all
        # variables of a random sequence share one scale here, hence the same factor for the whole array)
        for dtype, tol in (("f64", 1e-10), ("f32", 5e-5)):
            for lanes in (0, 2, 32, 128):
                try:
                    got, k = _run(epg, seq, opts, jac, dtype, kernel, lanes)
                except (NotImplementedError, MemoryError):
                    continue
                assert np.abs(got[0] - rs).max() < tol * ss, (kernel, dtype, lanes)
                assert np.abs(got[1] - rj).max() < tol * sj, (kernel, dtype, lanes)
