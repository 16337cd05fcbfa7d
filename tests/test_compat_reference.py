"""Drop-in check against the REAL reference (where it is importable: /root/reference in the build container,
baseline/_ref -- the pip --target install made by __graft_entry__.build() -- on the GPU box).  Sequences are built
with the reference's own operator classes, converted with `compat.from_reference`, run through the lowering -- by the
tape interpreter in the CPU suite, by the CUDA ENGINE in `-m gpu` -- and compared with the reference's own `simulate`."""

import os
import sys

import numpy as np
import pytest

import cases
import tape_interp
from util import RTOL64, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("EPGPY_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")]


@pytest.fixture(scope="module")
def ref():
    for root in CANDIDATES:
        if root and os.path.isdir(os.path.join(root, "epgpy")):
            sys.path.insert(0, root)
            try:
                import epgpy
            finally:
                sys.path.remove(root)
            return epgpy
    pytest.skip("the reference package is not available on this machine")


def _engine_simulate(engine):
    """`simulate` of the path under test: the numpy tape interpreter (CPU suite) or the CUDA engine (-m gpu)"""
    if engine == "interp":
        return lambda seq, **kw: tape_interp.simulate(None, seq, **kw)
    import epgpy_b200

    return epgpy_b200.epg.simulate


ENGINES = ["interp", pytest.param("cuda", marks=pytest.mark.gpu)]


NAMES = ["readme_mse", "mse_grid", "fisp_bounded", "fisp_jac_global", "mse_jac", "jac_all_params", "gre_diffusion_1d",
         "gre_diffusion_tensor", "bssfp_mt", "spgr_exchange", "hyperecho", "misc_ops", "adc_reduce", "gre_gradient_time"]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", NAMES)
def test_reference_objects_run_on_the_engine_path(name, engine, ref):
    import epgpy_b200
    from epgpy_b200 import compat

    simulate = _engine_simulate(engine)

    rns = cases.namespace(ref)
    case = cases.CASES[name](rns)                      # operators of the reference
    want_sig, want_jac = cases.run_api(rns, case)      # the reference's own simulate
    epg = cases.namespace(epgpy_b200)
    seq = compat.from_reference(case["seq"])
    opts = dict(case.get("options") or {})
    if case.get("density") is not None:
        opts["init"] = epg.StateMatrix(density=case["density"])
    if case.get("jac"):
        sig, jac = simulate(seq, probe=[None, epg.Jacobian(case["jac"])], **opts)
        assert rel_err(jac, want_jac) < RTOL64
    else:
        sig = simulate(seq, **opts)
    assert rel_err(np.asarray(sig), want_sig) < RTOL64
    got_t, want_t = epg.get_adc_times(seq), ref.core.get_adc_times(case["seq"])
    assert len(got_t) == len(want_t)
    for a, b in zip(got_t, want_t):
        assert np.allclose(np.asarray(a, dtype=float), np.asarray(b, dtype=float), rtol=1e-13, atol=0)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", sorted(cases.FOURIER_CASES))
def test_reference_fourier_probes_run_on_the_engine_path(name, engine, ref):
    """sequences built with the reference's own G / C / DFT / Imaging objects, converted and run on the engine path"""
    from epgpy_b200 import compat

    rns = cases.namespace(ref)
    case = cases.FOURIER_CASES[name](rns)
    seq = compat.from_reference(case["seq"])  # (before the reference runs: its Imaging._acquire POPS modulation / weights
    want = cases.run_probes(rns, case)        # from the probe's options, probe.py:205-210)
    got = _engine_simulate(engine)(seq, asarray=False, **case["options"])
    assert len(got) == len(want)
    for g, w in zip(got, want):
        w = np.asarray(w)
        assert np.shape(g) == w.shape and np.abs(np.asarray(g) - w).max() <= 1e-10 * max(np.abs(w).max(), 1e-30)


def _random_lattice_sequence(R, seed):
    """random gradient / time-accumulation train of the REFERENCE's operators with spoilers, resets, and DFT / Imaging /
    Adc probes in one to three dimensions (wavenumbers on the grid: k * kvalue = multiples of kgrid)"""
    rng = np.random.RandomState(seed)
    g0 = 0.1 / (2 * np.pi * 42576.0 * 1e-3)
    T1, T2 = np.array([500.0, 1200.0]), np.array([40.0, 80.0, 160.0])[None, :]
    use_c, kdim = rng.rand() < 0.6, int(rng.choice([1, 2, 3]))
    x = rng.uniform(-2, 2, (4, kdim))
    seq = []
    for _ in range(rng.randint(4, 9)):
        seq.append(R.T(rng.uniform(10, 170), rng.uniform(-180, 180)))
        for _ in range(rng.randint(1, 3)):
            gv = [int(v) * g0 for v in rng.randint(-3, 4, kdim)]
            if any(gv):
                seq.append(R.G(1.0, gv))
            if use_c and rng.rand() < 0.7:
                seq.append(R.C(float(rng.choice([0.5, 1.0, 1.5]))))
        seq.append(R.E(rng.uniform(1, 8), T1, T2))
        r = rng.rand()
        if r < 0.15:
            seq.append(R.SPOILER)
        elif r < 0.22:
            seq.append(R.RESET)
        r = rng.rand()
        if r < 0.4:
            seq.append(R.DFT(x))
        elif r < 0.7:
            mod = complex(rng.uniform(0, 0.5), rng.uniform(-0.1, 0.1)) if use_c else None
            seq.append(R.Imaging(x, voxel_size=float(rng.uniform(0.2, 1.5)), modulation=mod, reduce=False))
        else:
            seq.append(R.Adc(phase=float(rng.uniform(-90, 90))))
    return seq, dict(kvalue=2.5, tvalue=2.0, kgrid=0.25)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("seed", range(12))
def test_random_gradient_trains_with_fourier_probes(seed, engine, ref):
    """fuzz of the configuration lattice with accumulated time and of the probes over several configurations against the
    unmodified reference (shift-merge on the grid, statematrix.F0 with a time coordinate, probe.DFT / Imaging)"""
    from epgpy_b200 import compat

    seq, opts = _random_lattice_sequence(ref.core, seed)
    conv = compat.from_reference(seq)
    want = ref.simulate(seq, asarray=False, **opts)
    got = _engine_simulate(engine)(conv, asarray=False, **opts)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        w = np.asarray(w)
        assert np.shape(g) == w.shape
        assert np.abs(np.asarray(g) - w).max() <= 1e-10 * max(np.abs(w).max(), 1e-3)


@pytest.mark.parametrize("engine", ENGINES)
def test_slice_selective_pulse_like_the_reference(engine, ref):
    """`rfpulse.encode_phase` (rfpulse.py:321-345) and the small `utils` helpers it uses: excitation profile of a windowed
    sinc across a slice, with and without rephasing, against the reference's own helper + simulate"""
    import epgpy_b200
    from epgpy import rfpulse as rrf
    from epgpy_b200 import rfpulse as brf, utils as butils

    simulate = _engine_simulate(engine)
    t = np.linspace(-3, 3, 31)
    values = np.sinc(t) * np.hanning(31)
    for fov, kw in ((20.0, dict(rewind=True)), (20.0, dict(rewind=0.4, npoint=21)), (np.linspace(-8, 8, 13), dict(expand=False))):
        rs = rrf.encode_phase(rrf.RFPulse(values, 3.0, alpha=70, phi=15.0), 12.0, fov, **kw)
        bs = brf.encode_phase(brf.RFPulse(values, 3.0, alpha=70, phi=15.0), 12.0, fov, **kw)
        want = np.asarray(ref.simulate(list(rs) + [ref.core.ADC, ref.core.Adc("Z0")]))
        got = np.asarray(simulate(list(bs) + [epgpy_b200.epg.ADC, epgpy_b200.epg.Adc("Z0")]))
        assert got.shape == want.shape and rel_err(got, want) < RTOL64
    with pytest.raises(TypeError):
        brf.encode_phase(epgpy_b200.epg.T(30, 0), 12.0, 20.0)
    x = np.linspace(-3, 3, 7)
    assert np.allclose(butils.freq_to_space(9.0, butils.space_to_freq(9.0, x)), x)
    assert np.allclose(butils.space_to_freq(9.0, x), ref.utils.space_to_freq(9.0, x))
    assert np.allclose(butils.spatial_range(30.0, 11), ref.utils.spatial_range(30.0, 11))
    sm = np.zeros((2, 5, 3), dtype=complex)
    sm[:, 2, 2] = 1
    sm[:, 3, 0], sm[:, 1, 1] = 0.3 + 0.1j, 0.3 - 0.1j
    assert butils.check_states(sm) and np.allclose(butils.get_norm(sm), ref.utils.get_norm(sm))


def test_identity_is_preserved(ref):
    from epgpy_b200 import compat

    e = ref.core.E(5, 1000, 30)
    seq = [ref.core.T(90, 90), [ref.core.S(1), e, ref.core.T(150, 0), ref.core.S(1), e, ref.core.ADC]]
    out = compat.from_reference(seq)
    assert out[1][1] is out[1][4] and type(out[1][1]).__name__ == "E"


@pytest.mark.parametrize("engine", ENGINES)
def test_reference_sequence_layer_runs_unchanged_on_the_engine_path(engine, ref, monkeypatch):
    """the reference's symbolic `Sequence` layer (signal / jacobian / crlb) only calls
    `functions.simulate`; with that one call routed through `compat.from_reference` + the lowering, it runs
    unchanged (reference test/test_sequence.py:6-61 restated)."""
    import epgpy_b200
    from epgpy_b200 import compat
    from epgpy.sequence import Sequence, Variable, operators

    T2, B1 = Variable("T2"), Variable("B1")
    necho = 5
    exc, rfc = operators.T(90, 90), operators.T(180 * B1, 0)
    spl, rlx, adc = operators.S(1, duration=5), operators.E(5, 1400, T2), operators.ADC
    seq = Sequence([exc] + [spl, rlx, rfc, spl, rlx, adc] * necho)

    want_sig = seq.signal(T2=30, B1=0.8)
    want = seq.jacobian(["T2", "B1"], T2=[30, 40, 50], B1=0.8)
    want_crlb = seq.crlb(["T2", "B1"])(T2=30, B1=0.8)

    calls = []
    simulate = _engine_simulate(engine)

    def routed(sequence, **kw):
        calls.append(len(sequence))
        probe = kw.pop("probe", None)
        if probe is not None:
            probe = [compat.from_reference(p) if p is not None else None for p in (probe if isinstance(probe, (list, tuple)) else [probe])]
        kw.pop("asarray", None)
        return simulate(compat.from_reference(sequence), probe=probe, **kw)

    import epgpy.sequence as refseq

    monkeypatch.setattr(refseq._functions, "simulate", routed)
    got_sig = seq.signal(T2=30, B1=0.8)
    got = seq.jacobian(["T2", "B1"], T2=[30, 40, 50], B1=0.8)
    got_crlb = seq.crlb(["T2", "B1"])(T2=30, B1=0.8)
    assert calls, "the sequence layer did not go through the routed simulate"
    assert rel_err(got_sig, want_sig) < RTOL64
    assert rel_err(got[0], want[0]) < RTOL64 and rel_err(got[1], want[1]) < RTOL64
    assert np.allclose(got_crlb, want_crlb, rtol=1e-8)


@pytest.mark.parametrize("engine", ENGINES)
def test_reference_hessian_and_crlb_gradient_run_on_the_engine_path(engine, ref, monkeypatch):
    """the reference's `Sequence.hessian` and `Sequence.crlb(..., gradient=...)` -- the inner loop of
    its sequence optimisation (examples/differentiation/optim_mrf.py:127-149) -- with their one `simulate` call routed
    through `compat.from_reference` + the lowering: order-2 partial states on the engine, CRLB and its gradient within
    1e-8 of the reference's own."""
    from epgpy_b200 import compat
    from epgpy.sequence import Sequence, Variable, operators

    T1, T2 = Variable("T1"), Variable("T2")
    ntr = 12
    alphas = [Variable(f"alpha_{i:02d}") for i in range(ntr)]
    ops_ = [operators.T(180, 0), operators.E(20, T1, T2)]
    for i in range(ntr):
        ops_ += [operators.T(alphas[i], 90), operators.E(5, T1, T2), operators.ADC, operators.E(7, T1, T2), operators.S(1)]
    seq = Sequence(ops_)
    values = {"T1": 900.0, "T2": 70.0, **{f"alpha_{i:02d}": 15.0 + 3 * i for i in range(ntr)}}
    avars = [f"alpha_{i:02d}" for i in range(ntr)]

    want_h = seq.hessian(["T1", "T2"], avars, options={"max_nstate": 10})(values)
    want_c = seq.crlb(["T1", "T2"], gradient=avars, options={"max_nstate": 10})(values)

    calls = []
    simulate = _engine_simulate(engine)

    def routed(sequence, **kw):
        calls.append(len(sequence))
        probe = kw.pop("probe", None)
        if probe is not None:
            probe = [compat.from_reference(p) if p is not None else None for p in (probe if isinstance(probe, (list, tuple)) else [probe])]
        kw.pop("asarray", None)
        return simulate(compat.from_reference(sequence), probe=probe, **kw)

    import epgpy.sequence as refseq

    monkeypatch.setattr(refseq._functions, "simulate", routed)
    got_h = seq.hessian(["T1", "T2"], avars, options={"max_nstate": 10})(values)
    got_c = seq.crlb(["T1", "T2"], gradient=avars, options={"max_nstate": 10})(values)
    assert calls
    for g, w in zip(got_h, want_h):
        assert rel_err(np.asarray(g), np.asarray(w)) < RTOL64
    assert np.allclose(got_c[0], want_c[0], rtol=1e-8) and np.allclose(got_c[1], want_c[1], rtol=1e-8, atol=1e-8 * np.abs(want_c[1]).max())
