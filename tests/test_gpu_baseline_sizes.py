"""The five BASELINE.json configurations at (or near) their full sizes on the GPU, checked against the
oracle on sub-grids the oracle finishes in seconds, plus size-independent properties."""

import numpy as np
import pytest

import cases
import oracle_api
from util import RTOL32, RTOL64, product_namespace, rel_err

pytestmark = pytest.mark.gpu
O = oracle_api.O


@pytest.fixture(scope="module")
def epg():
    return product_namespace()


def test_c2_mse_t2_mapping_dictionary_full_grid(epg):
    """configs[1]: 17 echoes, T2 x B1 x T1 = 200 x 100 x 10 atoms, FP64; every 20th/10th/3rd atom vs the oracle"""
    def build(e, T2, B1, T1):
        B1 = np.asarray(B1)[None, :]
        T1 = np.asarray(T1)[None, None, :]
        exc, rfc, rlx, sh = e.T(90 * B1, 90), e.T(180 * B1, 0), e.E(4.75, T1, np.asarray(T2)), e.S(1)
        return [exc] + [[sh, rlx, rfc, sh, rlx, e.ADC]] * 17

    T2, B1, T1 = np.linspace(20, 300, 200), np.linspace(0.5, 1.2, 100), np.linspace(500, 3000, 10)
    sig = epg.simulate(build(epg, T2, B1, T1))
    assert sig.shape == (17, 200, 100, 10)
    ref = O.simulate(build(oracle_api.epg, T2[::20], B1[::10], T1[::3]))
    assert rel_err(sig[:, ::20, ::10, ::3], ref) < RTOL64
    sig32 = epg.simulate(build(epg, T2, B1, T1), dtype="float32")
    assert rel_err(sig32[:, ::20, ::10, ::3], ref) < RTOL32
    # CPMG property: with B1 = 1 and T1 >> TE the echo train is exp(-n TE / T2)
    j = int(np.argmin(np.abs(B1 - 1.0)))
    T2x = epg.simulate(build(epg, T2, [1.0], [1e9]))[:, :, 0, 0]
    assert np.allclose(np.abs(T2x), np.exp(-9.5 * np.arange(1, 18)[:, None] / T2[None, :]), rtol=1e-9)


def test_c3_fisp_dictionary_1000_tr(epg):
    """configs[2]: 1000 TRs; a 40 x 40 x 25 = 40 k atom slab of the 1 M grid on the device, a 3 x 3 x 2 sub-grid
    against the oracle (unbounded states: the oracle carries 2001 rows), FP64 and FP32"""
    import bench

    T1, T2, B1 = bench.grid_axes((100, 100, 100))
    T1, T2, B1 = T1[::3][:34], T2[::3][:34], B1[::4]
    seq = bench.fisp_sequence(epg, T1, T2, B1, 1000)
    sig = epg.simulate(seq)
    assert sig.shape == (1000, 34, 34, 25)
    sub = (slice(None), slice(2, 34, 12), slice(5, 34, 12), slice(3, 25, 15))
    ref = O.simulate(bench.fisp_sequence(oracle_api.epg, T1[sub[1]], T2[sub[2]], B1[sub[3]], 1000))
    assert rel_err(sig[sub], ref) < RTOL64
    sig32 = epg.simulate(seq, dtype="float32")
    assert rel_err(sig32[sub], ref) < RTOL32
    assert np.abs(sig.imag).max() == 0  # a real-valued phase graph
    # bounded variant (max_nstate = 32) against the oracle's own truncation
    sigb = epg.simulate(seq, max_nstate=32)
    refb = O.simulate(bench.fisp_sequence(oracle_api.epg, T1[sub[1]], T2[sub[2]], B1[sub[3]], 1000), max_nstate=32)
    assert rel_err(sigb[sub], refb) < RTOL64


def test_c3_fisp_jacobian_1000_tr(epg):
    """configs[2] "plus its flip-angle Jacobian" (SURVEY 8d, M3J(i)): 1000 TRs, variables B1 (the flip-angle scale), T1,
    T2; a 12 x 10 x 8 slab on the device through the warp-per-state-set kernel (83 pure whole-TR windows), a
    2 x 2 x 2 sub-grid against the oracle, FP64 and FP32, and the orders-over-warps kernel on the same tape"""
    import bench
    from epgpy_b200 import engine, functions, lowering

    T1, T2, B1 = bench.grid_axes((100, 100, 100))
    T1, T2, B1 = T1[::9][:12], T2[::10][:10], B1[::13][:8]
    names = ["B1", "T1", "T2"]
    seq = bench.fisp_sequence(epg, T1, T2, B1, 1000, jac=True)
    sub = (slice(None), slice(1, 12, 8), slice(2, 10, 6), slice(0, 8, 5))
    rs, rj = O.simulate(bench.fisp_sequence(oracle_api.epg, T1[sub[1]], T2[sub[2]], B1[sub[3]], 1000, jac=True), jacobian=names)
    rs, rj = np.asarray(rs), np.asarray(rj)
    for dtype, tol in (("f64", RTOL64), ("f32", RTOL32)):
        low = lowering.lower(seq, probe=[None, epg.Jacobian(names)], dtype=dtype)
        for variant, kernel in ((0, 4), (4, 3)):  # automatic choice: one warp per state set; forced: orders over warps
            plan = engine.Plan(low)
            if variant:
                plan.set_variant(kernel=variant)
            assert plan.config()["kernel"] == kernel
            parts, _ = functions.run_lowered(low, plan=plan)
            sig, jac = functions._assemble(low, parts)
            assert sig.shape == (1000, 12, 10, 8) and jac.shape == (1000, 12, 10, 8, 3)
            assert rel_err(sig[sub], rs) < tol
            for i in range(3):  # columns of very different magnitude: compare one by one
                assert rel_err(jac[sub][..., i], rj[..., i]) < (tol if dtype == "f64" else 5 * tol)


def test_c4_rf_spoiled_gre_3d_gradients_diffusion_500_tr(epg):
    """configs[3]: quadratic RF phase, 3-d (collinear) gradient shifts, isotropic diffusion, 500 TRs"""
    def build(e, T1, T2, ntr=500):
        T2 = np.asarray(T2)[None, :]
        seq, kv = [], [2, 1, -1]
        for n in range(ntr):
            ph = 117.0 * n * (n + 1) / 2
            seq.append([e.T(15, ph), e.E(2, T1, T2), e.Adc(phase=-ph), e.E(8, T1, T2), e.S(kv), e.D(10, 2e-3, k=kv)])
        return seq

    T1, T2 = np.linspace(400, 2000, 40), np.linspace(30, 200, 30)
    sig = epg.simulate(build(epg, T1, T2), kvalue=500.0)
    assert sig.shape == (500, 40, 30)
    ref = O.simulate(build(oracle_api.epg, T1[::13], T2[::10]), kvalue=500.0, kvec=[2, 1, -1])
    assert rel_err(sig[:, ::13, ::10], ref) < RTOL64
    sig20 = epg.simulate(build(epg, T1, T2), kvalue=500.0, max_nstate=20)  # examples/gradient/random_spoiling.py:50
    ref20 = O.simulate(build(oracle_api.epg, T1[::13], T2[::10]), kvalue=500.0, kvec=[2, 1, -1], max_nstate=20)
    assert rel_err(sig20[:, ::13, ::10], ref20) < RTOL64


def test_c5_two_pool_mt_bssfp_500_tr_and_pulse_jacobian(epg):
    """configs[4]: EPG-X two-pool MT bSSFP (gre_exchange.py model2), 500 TRs, 101 off-resonances x 6 flip
    angles; forward vs the oracle, and the per-pulse flip-angle Jacobian (exact chain rule through X) vs
    central finite differences of the oracle's forward signal"""
    T1, T2, khi, f = [779.0, 779.0], [45.0, 12e-3], 4.3e-3, [1 - 0.117, 0.117]
    TR = 5.0

    def build(e, offres, FA, ntr=500, jac_pulses=(), dalpha=None):
        kmat = e.exchange_matrix(khi, densities=f)
        FA = np.asarray(FA, dtype=float)[None, None, :]
        sat = e.R(rL=[0, 0.0316])
        exg = e.X(TR, kmat, T1=T1, T2=T2, g=[np.asarray(offres)])
        seq = []
        pool = np.array([1.0, 0.0])[:, None, None]  # the RF pulse acts on the free pool only
        for i in range(ntr):
            a = FA + (dalpha[1] if dalpha is not None and dalpha[0] == i else 0.0)
            phase = 0.0 if i % 2 == 0 else 180.0
            if i in jac_pulses:
                seq += [e.T(a * pool, phase, order1={f"a{i}": {"alpha": pool}}), sat, exg]
            else:
                seq += [e.T(a * pool, phase) @ sat, exg]
        seq += [e.T(FA * np.array([1.0, 0.0])[:, None, None], 0.0) @ sat, e.Adc(reduce=0)]
        return seq

    offres = 1 / TR * np.linspace(-0.5, 0.5, 101)
    FA = np.array([5.0, 10.0, 15.0, 20.0, 30.0, 45.0])
    sig = epg.simulate(build(epg, offres, FA), init=epg.StateMatrix(density=f))
    assert sig.shape == (1, 101, 6)
    ref = O.simulate(build(oracle_api.epg, offres[::25], FA[::2]), density=f)
    assert rel_err(sig[:, ::25, ::2], ref) < RTOL64

    # Jacobian w.r.t. the flip angle of three pulses of a 40-TR train (partials propagated through X)
    pulses, ntr = (3, 17, 38), 40
    s, jac = epg.simulate(build(epg, offres[::10], FA[::2], ntr, pulses), init=epg.StateMatrix(density=f),
                          probe=[None, epg.Jacobian([f"a{i}" for i in pulses])], propagate_nondiff=True)
    h = 1e-4
    for c, i in enumerate(pulses):
        up = O.simulate(build(oracle_api.epg, offres[::10], FA[::2], ntr, dalpha=(i, +h)), density=f)
        dn = O.simulate(build(oracle_api.epg, offres[::10], FA[::2], ntr, dalpha=(i, -h)), density=f)
        fd = (up - dn) / (2 * h)
        # the Jacobian probe reads the un-reduced F0 of both pools; the ADC sums over the pool axis
        assert rel_err(jac[..., c].sum(axis=1), fd) < 1e-6
