"""Random sequences through the lowering + tape interpreter against the oracle (CPU)."""

import numpy as np
import pytest

import fuzz
import oracle_api
import tape_interp
from util import product_namespace, rel_err


@pytest.mark.parametrize("seed", range(40))
def test_random_sequences_forward(seed):
    epg = product_namespace()
    seq, opts, _ = fuzz.random_case(epg, seed, real_only=seed % 3 == 0)
    ref_seq, _, _ = fuzz.random_case(oracle_api.epg, seed, real_only=seed % 3 == 0)
    ref = oracle_api.O.simulate(ref_seq, kvalue=opts["kvalue"], max_nstate=opts.get("max_nstate"))
    for prune in (True, False):
        got = tape_interp.simulate(None, seq, prune_unobservable=prune, **opts)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("seed", range(100, 125))
def test_random_sequences_jacobian(seed):
    epg = product_namespace()
    seq, opts, jac = fuzz.random_case(epg, seed, real_only=seed % 2 == 0, with_jac=True)
    ref_seq, _, _ = fuzz.random_case(oracle_api.epg, seed, real_only=seed % 2 == 0, with_jac=True)
    rs, rj = oracle_api.O.simulate(ref_seq, kvalue=opts["kvalue"], max_nstate=opts.get("max_nstate"), jacobian=jac,
                                   propagate_nondiff=True)
    s, j = tape_interp.simulate(None, seq, probe=[None, epg.Jacobian(jac)], propagate_nondiff=True, **opts)
    assert np.abs(s - rs).max() < 1e-11 * max(1.0, np.abs(rs).max())
    assert np.abs(j - rj).max() < 1e-10 * max(1.0, np.abs(rj).max())
