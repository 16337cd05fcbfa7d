"""Shared workload definitions for the parity suite.

Every case is written ONCE against the reference's public operator API
(`epg.T/E/S/D/X/...`, nested lists, `@`), and is built three times:
  * with the unmodified reference (`tests/golden/make_golden.py`, build container only),
  * with the CPU oracle through the thin adapter `tests/oracle_api.py`,
  * with the product `epgpy_b200` (drop-in: same constructors).

A case function takes the API namespace and returns a dict:
  seq      nested list of operators
  options  simulate(**options) keywords (max_nstate, kvalue)
  density  None | list   -> init=StateMatrix(density=...)
  jac      None | list of variable names -> probe=[ADC, Jacobian(jac)]
  kvec     None | base shift vector (oracle only: collinear n-d shifts)
The sizes are small enough for the reference/oracle to finish in seconds; the
BASELINE.json configs are reproduced at reduced grid size (same operators).
"""

import numpy as np


def _fisp_schedule(ntr, seed=0):
    rng = np.random.RandomState(seed)
    fa = 10 + 50 * np.abs(np.sin(np.arange(ntr) * np.pi / 200))
    tr = rng.uniform(11, 16, ntr)
    return fa, tr


def readme_mse(epg):
    """BASELINE configs[0] verbatim (reference README.md:52-74)"""
    FA, ESP, Necho, T1, T2 = 120, 10, 20, 150, [30, 40, 50]
    exc, rfc = epg.T(90, 90), epg.T(FA, 0)
    rlx = epg.E(ESP / 2, T1, T2)
    shift = epg.S(1, duration=ESP / 2)
    seq = [exc] + [[shift, rlx, rfc, shift, rlx, epg.ADC]] * Necho
    return dict(seq=seq)


def mse_grid(epg, nT2=6, nB1=5, nT1=3):
    """BASELINE configs[1] at reduced grid"""
    T2 = np.linspace(20, 300, nT2)
    B1 = np.linspace(0.5, 1.2, nB1)[None, :]
    T1 = np.linspace(500, 3000, nT1)[None, None, :]
    exc = epg.T(90 * B1, 90)
    rfc = epg.T(180 * B1, 0)
    rlx = epg.E(4.75, T1, T2)
    sh = epg.S(1)
    seq = [exc] + [[sh, rlx, rfc, sh, rlx, epg.ADC]] * 17
    return dict(seq=seq)


def fisp(epg, ntr=60, sizes=(3, 4, 2), max_nstate=None, g=None):
    """BASELINE configs[2] at reduced size.
    NOTE the axis sizes are pairwise different on purpose: the reference's in-place matmul
    (epgpy/opmatrix.py:208-221) silently mis-aligns an operator axis i with grid axis i-1 when
    their sizes coincide (e.g. 3x3x3 or 100x100x100), giving wrong signals; see DESIGN.md."""
    fa, tr = _fisp_schedule(ntr)
    T1 = np.linspace(300, 3000, sizes[0])
    T2 = np.linspace(20, 300, sizes[1])[None, :]
    B1 = np.linspace(0.7, 1.2, sizes[2])[None, None, :]
    kw = {} if g is None else {"g": g}
    seq = [epg.T(180, 0), epg.E(20, T1, T2, **kw)]
    for i in range(ntr):
        seq.append([epg.T(fa[i] * B1, 90), epg.E(3, T1, T2, **kw), epg.ADC,
                    epg.E(tr[i] - 3, T1, T2, **kw), epg.S(1)])
    opts = {} if max_nstate is None else {"max_nstate": max_nstate}
    return dict(seq=seq, options=opts)


def fisp_unbounded(epg):
    return fisp(epg, 60)


def fisp_bounded(epg):
    return fisp(epg, 80, max_nstate=10)


def bssfp_offres(epg, ntr=50):
    """bSSFP variant of M3: no shift, off-resonance axis (n = 0 all along)"""
    T1, T2 = 1000.0, np.array([40.0, 80.0, 120.0])
    g = np.linspace(-0.1, 0.1, 11)[None, :]
    rf1, rf2 = epg.T(30, 0), epg.T(30, 180)
    rlx = epg.E(5, T1, T2, g)
    seq = [[rf1, rlx], [rf2, rlx]] * (ntr // 2) + [[rf1, epg.ADC]]
    return dict(seq=seq)


def fisp_jac_global(epg, ntr=30):
    """M3J(i): global variables B1, T1, T2"""
    fa, tr = _fisp_schedule(ntr)
    T1 = np.linspace(300, 3000, 2)
    T2 = np.linspace(20, 300, 3)[None, :]
    B1 = np.linspace(0.7, 1.2, 2)[None, None, :]
    o1 = ["T1", "T2"]
    seq = [epg.T(180, 0), epg.E(20, T1, T2, order1=o1)]
    for i in range(ntr):
        seq.append([epg.T(fa[i] * B1, 90, order1={"B1": {"alpha": fa[i]}}), epg.E(3, T1, T2, order1=o1), epg.ADC,
                    epg.E(tr[i] - 3, T1, T2, order1=o1), epg.S(1)])
    return dict(seq=seq, jac=["B1", "T1", "T2"])


def fisp_jac_pulses(epg, ntr=16, max_nstate=10):
    """M3J(ii): per-pulse flip-angle variables, bounded states (optim_mrf.py:96)"""
    fa, tr = _fisp_schedule(ntr)
    T1 = np.linspace(300, 3000, 2)
    T2 = np.linspace(20, 300, 3)[None, :]
    names = [f"alpha_{i:03d}" for i in range(ntr)]
    seq = [epg.T(180, 90), epg.E(20, T1, T2)]
    for i in range(ntr):
        seq.append([epg.T(fa[i], 90, order1={names[i]: "alpha"}), epg.E(3, T1, T2), epg.ADC,
                    epg.E(tr[i] - 3, T1, T2), epg.S(1)])
    return dict(seq=seq, jac=["magnitude"] + names, options={"max_nstate": max_nstate})


def mse_jac(epg):
    """reference test/test_diff.py:282-331 chain, as a simulate() call"""
    exc = epg.T(90, 90)
    ref = epg.T(150, 0, order1="alpha")
    relax = epg.E(5, 1e3, 35, order1="T2")
    grad = epg.S(1)
    seq = [exc] + [grad, relax, ref, grad, relax, epg.ADC] * 5
    return dict(seq=seq, jac=["T2", "alpha"])


def jac_all_params(epg):
    """phi / tau / T1 / g derivatives with array coefficients (diff.py:556-568)"""
    T2 = np.array([30.0, 50.0, 80.0])
    g = np.array([[0.0, 0.02]])
    c = np.array([[1.0], [2.0], [3.0]])  # coefficient arrays need the full grid ndim in the reference
    rf = epg.T(40, [[15.0, 60.0]], order1={"a": {"alpha": c}, "p": {"phi": 1.0}})
    rlx = epg.E(6.0, 800.0, T2, g, order1={"tau": {"tau": 1}, "T1": {"T1": 1}, "g": {"g": 1}, "a": {"T2": 0.5}})
    ph = epg.Phi(20.0, order1={"p": {"phi": 2.0}})
    pr = epg.P(2.0, g, order1={"g": "g", "tau": "tau"})
    seq = [epg.T(90, 90)] + [rf, rlx, epg.S(1), ph, pr, epg.ADC] * 6
    return dict(seq=seq, jac=["a", "p", "tau", "T1", "g", "missing"])


def gre_diffusion(epg, ntr=40):
    """BASELINE configs[3] at reduced size: RF spoiling, 3-d shift, isotropic D"""
    T1 = np.array([600.0, 1200.0])
    T2 = np.array([[40.0, 80.0, 120.0]])
    seq = []
    kv = [2, 1, -1]
    for n in range(ntr):
        ph = 117.0 * n * (n + 1) / 2
        seq.append([epg.T(15, ph), epg.E(2, T1, T2), epg.Adc(phase=-ph), epg.E(8, T1, T2),
                    epg.S(kv), epg.D(10, 2e-3, k=kv)])
    return dict(seq=seq, options={"kvalue": 500.0}, kvec=kv)


def gre_diffusion_1d(epg, ntr=30):
    """1-d integer shifts, D with and without the ramp term, tensor-free; max_nstate"""
    T2 = np.array([40.0, 80.0])
    seq = [epg.T(90, 90)]
    for n in range(ntr):
        seq.append([epg.S(1), epg.D(3.0, 1.5e-3, k=1), epg.E(3, 900.0, T2), epg.D(2.0, 1.5e-3),
                    epg.T(35, 0), epg.ADC])
    return dict(seq=seq, options={"kvalue": 3000.0, "max_nstate": 12})


def _mt_model():
    T1, T2, khi, f = [779.0, 779.0], [45.0, 12e-3], 4.3e-3, [1 - 0.117, 0.117]
    return T1, T2, khi, f


def bssfp_mt(epg, ntr=40, noff=7):
    """BASELINE configs[4] forward (examples/exchange/gre_exchange.py:163-175, model2)"""
    T1, T2, khi, f = _mt_model()
    kmat = epg.exchange_matrix(khi, densities=f)
    FA, TR = 10, 5
    offres = 1 / TR * np.linspace(-0.5, 0.5, noff)
    sat = epg.R(rL=[0, 0.0316])
    rf1 = epg.T([FA, 0], 0) @ sat
    rf2 = epg.T([FA, 0], 180) @ sat
    exg = epg.X(TR, kmat, T1=T1, T2=T2, g=[offres])
    adc = epg.Adc(reduce=0)
    seq = [[rf1, exg], [rf2, exg]] * (ntr // 2) + [[rf1, adc]]
    return dict(seq=seq, density=f)


def bssfp_mt_pulse_jac(epg, ntr=8, noff=5, dalpha=None, diff=True):
    """BASELINE configs[4] with its Jacobian: two-pool MT bSSFP, one order-1 variable per pulse (flip angle of both
    pools).  Not in CASES: the reference does not propagate partial states through X, so the oracle for
    this Jacobian is the central finite difference of the forward signal (`dalpha`: per-pulse perturbation, degrees)."""
    T1, T2, khi, f = _mt_model()
    kmat = epg.exchange_matrix(khi, densities=f)
    FA, TR = 10.0, 5
    offres = 1 / TR * np.linspace(-0.5, 0.5, noff)
    sat = epg.R(rL=[0, 0.0316])
    exg = epg.X(TR, kmat, T1=T1, T2=T2, g=[offres])
    seq = []
    for i in range(ntr):
        d = 0.0 if dalpha is None else float(dalpha[i])
        kw = dict(order1={f"a{i}": "alpha"}) if diff else {}
        seq += [epg.T([FA + d, d], 0 if i % 2 == 0 else 180, **kw) @ sat, exg, epg.Adc(reduce=0)]
    return dict(seq=seq, density=f, jac=[f"a{i}" for i in range(ntr)] if diff else None)


def spgr_exchange(epg, ntr=30):
    """two-pool water exchange SPGR with RF spoiling + shift (gre_exchange.py:73-87, model1)"""
    T1, T2, khi, f = [1000.0, 500.0], [100.0, 20.0], 2e-3, [0.8, 0.2]
    kmat = epg.exchange_matrix(khi, densities=f)
    PH = np.array([50.0, 117.0, 150.0])  # 3 values: a 2-value axis hits a reference in-place matmul misalignment
    exg = epg.X(5, kmat, T1=T1, T2=T2)
    adc = epg.Adc(reduce=0)
    seq = [[epg.T(10, [i * (i + 1) / 2 * PH]), adc, exg, epg.S(1)] for i in range(ntr)]
    return dict(seq=seq, density=f, options={"max_nstate": 8})


def three_pool_exchange(epg, ntr=24):
    """THREE exchanging compartments (exchange.py:14-81 is generic in N; exchange_matrix(..., ncomp=3)): spoiled
    gradient echo with shifts, off-resonance axis, read-out summed over the pools"""
    T1, T2, f = [1000.0, 700.0, 400.0], [90.0, 40.0, 8.0], [0.6, 0.3, 0.1]
    kmat = epg.exchange_matrix(3e-3, ncomp=3, densities=f)
    g = [np.linspace(-0.02, 0.02, 4)]
    exg = epg.X(6, kmat, T1=T1, T2=T2, g=g)
    adc = epg.Adc(reduce=0)
    PH = np.array([50.0, 117.0, 150.0, 84.0])  # the pulse carries the second grid axis from the first operator on
    seq = [[epg.T(12 + i % 5, [i * (i + 1) / 2 * PH]), adc, exg, epg.S(1)] for i in range(ntr)]
    return dict(seq=seq, density=f, options={"max_nstate": 6})


def hyperecho(epg, npulse=15):
    """reference test/test_core.py:9-32 (shorter); probe both F0 and Z0 through two ADCs"""
    grad = epg.S(1)
    se1 = [grad, epg.T(10, 0), grad, epg.ADC, epg.Adc("Z0")]
    se2 = [grad, epg.T(-10, 0), grad, epg.ADC, epg.Adc("Z0")]
    seq = [epg.T(90, 90)] + se1 * npulse + [grad, epg.T(180, 0), grad] + se2 * npulse
    return dict(seq=seq)


def misc_ops(epg):
    """Phi, P, R, SPOILER, RESET, PD, Wait, negative and multi-step shifts, Adc phase/weights/reduce"""
    T2 = np.array([30.0, 60.0, 90.0])
    g = np.array([[-0.05, 0.0, 0.02, 0.07]])
    seq = [
        epg.PD([1.0, 0.5, 2.0]), epg.T(70, 30), epg.S(2), epg.E(4, 700.0, T2, g), epg.Phi([[10.0, 20, 30, 40]]),
        epg.Adc(phase=25.0), epg.T(120, [[0.0, 45, 90, 135]]), epg.S(-1), epg.P(3.0, g), epg.ADC,
        epg.R(rT=0.1 + 0.3j, rL=0.05, r0=0.02), epg.S(-2), epg.Wait(2.0), epg.T(45, 10), epg.Adc("Z0"),
        epg.SPOILER, epg.T(33, 77), epg.S(1), epg.E(5, 500.0, 50.0), epg.T(150, 0), epg.S(1), epg.ADC,
        epg.RESET, epg.T(20, 90), epg.Adc(phase=[10.0, 20.0, 30.0]),
    ]
    return dict(seq=seq)


def adc_reduce(epg):
    """weights / reduce / phase on the read-out (reference test/test_functions.py:69-76)"""
    T2 = np.array([[30.0], [40.0], [50.0]])
    g = np.array([[-0.1, -0.05, 0, 0.05, 1]])
    rlx = epg.E(10, 1000, T2, g=g)
    adc = epg.Adc(reduce=1, weights=[[1, 2, 3, 4, 5]], phase=[5.0, 10.0, 15.0])
    grad = epg.S(1, duration=10)
    seq = [epg.T(90, 90), grad, rlx, epg.T(180, 0), grad, rlx, adc, grad, rlx, epg.T(170, 0), grad, rlx, adc]
    return dict(seq=seq)


def gre_diffusion_tensor(epg, ntr=24):
    """anisotropic diffusion TENSOR with collinear 3-d shifts (diffusion.py:140-145), ramp term included"""
    T1 = np.array([600.0, 1200.0])
    T2 = np.array([[40.0, 80.0, 120.0]])
    Dten = np.array([[2.0, 0.1, 0.0], [0.1, 1.0, 0.2], [0.0, 0.2, 0.5]]) * 1e-3
    seq, kv = [], [2, 1, -1]
    for n in range(ntr):
        ph = 117.0 * n * (n + 1) / 2
        seq.append([epg.T(15, ph), epg.E(2, T1, T2), epg.Adc(phase=-ph), epg.E(8, T1, T2), epg.D(3.0, Dten),
                    epg.S(kv), epg.D(10, Dten, k=kv)])
    return dict(seq=seq, options={"kvalue": 500.0}, kvec=kv)


def _valid_state(n, seed, real=False):
    """a random (2n+1) x 3 state obeying F-(k) = conj F+(-k), Z(k) = conj Z(-k) (statematrix.py:418-421)"""
    rng = np.random.RandomState(seed)
    half = rng.randn(n + 1, 3) + (0 if real else 1j) * rng.randn(n + 1, 3)
    half[0, 2] = half[0, 2].real
    full = np.zeros((2 * n + 1, 3), dtype=complex)
    full[n:, :] = half
    full[:n, 0] = half[:0:-1, 1].conj()
    full[:n, 1] = half[:0:-1, 0].conj()
    full[:n, 2] = half[:0:-1, 2].conj()
    full[n, 1] = full[n, 0].conj()
    return 0.3 * full


def init_states(epg, ntr=14):
    """initial state with n = 3 populated orders and complex entries (functions.py:113-144): complex kernels"""
    T2 = np.array([40.0, 80.0, 120.0])
    g = np.array([[0.0, 0.03]])
    seq = []
    for i in range(ntr):
        seq.append([epg.T(25 + i, 30.0 * i), epg.E(4, 900.0, T2, g), epg.ADC, epg.Adc("Z0"), epg.S(1 if i % 4 else -2)])
    return dict(seq=seq, init=_valid_state(3, 11))


def init_states_real(epg, ntr=14):
    """real initial state with n = 3 populated orders under +-90 degree pulses: the real-valued kernels;
    max_nstate below the initial order count crops it at the first shift (statematrix.resize, shift.py:98)"""
    T2 = np.array([40.0, 80.0, 120.0])
    seq = []
    for i in range(ntr):
        seq.append([epg.T(25 + i, 90 if i % 3 else 270), epg.E(4, 900.0, T2), epg.ADC, epg.Adc("Z0"), epg.S(1 if i % 4 else -1)])
    return dict(seq=seq, init=_valid_state(3, 12, real=True))


def init_states_cropped(epg):
    """max_nstate BELOW the order count of the initial state: the first shift crops it (shift.py:98, statematrix.resize)"""
    case = init_states_real(epg)
    case["options"] = {"max_nstate": 2}
    return case


def init_states_jac(epg, ntr=10):
    """initial state with populated orders + order-1 variables: the partial states start from zero (diff.py:103-109)"""
    T2 = np.array([40.0, 80.0, 120.0])
    seq = []
    for i in range(ntr):
        seq.append([epg.T(25 + i, 90, order1={"B1": {"alpha": 25.0 + i}}), epg.E(4, 900.0, T2, order1=["T2"]), epg.ADC, epg.S(1)])
    return dict(seq=seq, init=_valid_state(2, 13, real=True), jac=["magnitude", "B1", "T2"])


def init_states_jac_complex(epg, ntr=10):
    T2 = np.array([40.0, 80.0, 120.0])
    seq = []
    for i in range(ntr):
        seq.append([epg.T(25 + i, 20.0 * i, order1={"B1": {"alpha": 25.0 + i}}), epg.E(4, 900.0, T2, 0.02, order1=["T2"]), epg.ADC,
                    epg.S(1)])
    return dict(seq=seq, init=_valid_state(2, 14), jac=["B1", "T2"])


CASES = {
    "readme_mse": readme_mse,
    "mse_grid": mse_grid,
    "fisp_unbounded": fisp_unbounded,
    "fisp_bounded": fisp_bounded,
    "bssfp_offres": bssfp_offres,
    "fisp_jac_global": fisp_jac_global,
    "fisp_jac_pulses": fisp_jac_pulses,
    "mse_jac": mse_jac,
    "jac_all_params": jac_all_params,
    "gre_diffusion": gre_diffusion,
    "gre_diffusion_1d": gre_diffusion_1d,
    "bssfp_mt": bssfp_mt,
    "spgr_exchange": spgr_exchange,
    "three_pool_exchange": three_pool_exchange,
    "hyperecho": hyperecho,
    "misc_ops": misc_ops,
    "adc_reduce": adc_reduce,
    "gre_diffusion_tensor": gre_diffusion_tensor,
    "init_states": init_states,
    "init_states_real": init_states_real,
    "init_states_cropped": init_states_cropped,
    "init_states_jac": init_states_jac,
    "init_states_jac_complex": init_states_jac_complex,
}


def fisp_equal_axes(epg, ntr=30, n=3):
    """FISP + (B1, T1, T2) Jacobian on an n x n x n grid: EQUAL axis sizes, like the headline 100 x 100 x 100
    dictionary.  The reference's own vectorised run is wrong on such grids, so the golden file is
    assembled from n^3 per-atom SCALAR runs of the unmodified reference (tests/golden/make_golden.py)."""
    fa, tr = _fisp_schedule(ntr)
    T1 = np.linspace(300, 3000, n)
    T2 = np.linspace(20, 300, n)
    B1 = np.linspace(0.7, 1.2, n)
    return dict(axes=(T1, T2, B1), jac=["B1", "T1", "T2"], build=lambda T1, T2, B1: _fisp_jac_seq(epg, fa, tr, T1, T2, B1))


def _fisp_jac_seq(epg, fa, tr, T1, T2, B1):
    o1 = ["T1", "T2"]
    seq = [epg.T(180, 0), epg.E(20, T1, T2, order1=o1)]
    for i in range(len(fa)):
        seq.append([epg.T(fa[i] * B1, 90, order1={"B1": {"alpha": fa[i]}}), epg.E(3, T1, T2, order1=o1), epg.ADC,
                    epg.E(tr[i] - 3, T1, T2, order1=o1), epg.S(1)])
    return seq


def hessian_ssfp(epg, ntr=6):
    """order-2 forward mode (epgpy/diff.py:290-378; reference test/test_diff.py:334-468): every differentiable
    operator carries order2=True, the Hessian probe picks pairs of variables and, through 'magnitude', first derivatives"""
    T2 = np.array([30.0, 50.0, 80.0])
    B1 = np.array([[0.8, 1.1]])
    seq = [epg.T(90, 90)]
    for i in range(ntr):
        seq += [epg.T((20 + 3 * i) * B1, 90 + 10 * i, order1=True, order2=True), epg.E(5, 1e3, T2, 0.01, order1=True, order2=True),
                epg.S(1), epg.ADC]
    return dict(seq=seq, hessian=(["alpha", "T2", "magnitude", "g"], ["T2", "phi", "tau", "alpha", "T1"]))


def hessian_fisp(epg, ntr=20):
    """FISP with global variables (B1 through a chain-rule coefficient, T1, T2): the Hessian of the dictionary atoms"""
    fa, tr = _fisp_schedule(ntr)
    T1 = np.linspace(300, 3000, 2)
    T2 = np.linspace(20, 300, 3)[None, :]
    B1 = np.linspace(0.7, 1.2, 2)[None, None, :]
    # (pairs spelled out: the reference's list-of-names spelling of `order2` raises, diff.py:214)
    pe = [("T1", "T1"), ("T1", "T2"), ("T2", "T2"), ("B1", "T1"), ("B1", "T2")]
    pt = [("B1", "B1"), ("B1", "T1"), ("B1", "T2")]
    seq = [epg.T(180, 0), epg.E(20, T1, T2, order1=["T1", "T2"], order2=pe)]
    for i in range(ntr):
        seq.append([epg.T(fa[i] * B1, 90, order1={"B1": {"alpha": fa[i]}}, order2=pt),
                    epg.E(3, T1, T2, order1=["T1", "T2"], order2=pe), epg.ADC,
                    epg.E(tr[i] - 3, T1, T2, order1=["T1", "T2"], order2=pe), epg.S(1)])
    v = ["B1", "T1", "T2"]
    return dict(seq=seq, hessian=(v, v), options={"max_nstate": 12})


HESSIAN_CASES = {"hessian_ssfp": hessian_ssfp, "hessian_fisp": hessian_fisp}


def run_hessian(epg, case, simulate=None, **extra):
    """(signal, hessian) of a HESSIAN_CASES entry with a reference-compatible API"""
    simulate = simulate or epg.simulate
    opts = dict(case.get("options") or {})
    opts.update(extra)
    sig, hes = simulate(case["seq"], probe=[None, epg.Hessian(*case["hessian"])], **opts)
    return np.asarray(sig), np.asarray(hes)


def slice_profile(epg, nsample=41, nfreq=15):
    """shaped RF pulse (rfpulse.py:37-197): a windowed sinc of 41 samples as a T / E chain over a frequency axis --
    excitation profile (F0 and Z0) of a 60 degree pulse followed by a rephasing precession, with relaxation"""
    t = np.linspace(-3, 3, nsample)
    values = np.sinc(t) * np.hanning(nsample) * np.exp(1j * 0.3 * t)
    freqs = np.linspace(-2.0, 2.0, nfreq)
    pulse = epg.RFPulse(values, 4.0, alpha=60, phi=20.0, T1=900.0, T2=[[60.0], [120.0]], g=[freqs])
    return dict(seq=[pulse, epg.P(-2.0, [freqs]), epg.ADC, epg.Adc("Z0")])


def gre_lattice_2d(epg, ntr=16):
    """general integer n-d shifts (the reference's `shift-nd` method, shift.py:103-117, 297-364): NON-collinear 2-d
    gradient moments that come back to k = 0, with diffusion on the 2-d wavenumbers"""
    T2 = np.array([40.0, 80.0])
    ks = [[1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [0, -2], [1, 0]]
    seq = []
    for i in range(ntr):
        k = ks[i % 8]
        seq += [epg.T(30 + 5 * i, 20.0 * i), epg.E(5, 800.0, T2), epg.ADC, epg.Adc("Z0"), epg.S(k), epg.D(4.0, 2e-3, k=k)]
    return dict(seq=seq, options={"kvalue": 300.0})


def gre_lattice_3d_cropped(epg, ntr=14):
    """3-d lattice cropped at max_nstate = 2, a spoiler and a reset in between, B1 axis"""
    B1 = np.array([[0.8, 1.0, 1.2]])
    T2 = np.array([40.0, 80.0])
    ks = [[1, 0, 0], [0, 1, -1], [-1, 0, 1], [0, -1, 0], [1, 1, 0], [-1, 0, 0]]
    seq = []
    for i in range(ntr):
        k = ks[i % 6]
        seq += [epg.T((25 + 4 * i) * B1, 35.0 * i), epg.E(6, 700.0, T2, 0.01), epg.ADC, epg.S(k)]
        if i == 6:
            seq += [epg.SPOILER]
        if i == 10:
            seq += [epg.RESET]
    return dict(seq=seq, options={"max_nstate": 2})


def lattice_jac(epg, ntr=10):
    """derivatives on a 2-d lattice: the partial states move through the same gather maps"""
    T2 = np.array([40.0, 80.0, 120.0])
    ks = [[1, 0], [0, 1], [-1, 0], [0, -1], [1, -1]]
    seq = []
    for i in range(ntr):
        seq += [epg.T(30 + 5 * i, 90.0, order1={"B1": {"alpha": 30.0 + 5 * i}}), epg.E(5, 800.0, T2, order1=["T2"]), epg.ADC,
                epg.S(ks[i % 5])]
    return dict(seq=seq, jac=["magnitude", "B1", "T2"])


def gre_lattice_float(epg, ntr=16):
    """FLOAT shifts on multiples of `kgrid` (the reference's shift-merge method, shift.py:119-145, 367-444): exactly
    the integer lattice in grid units"""
    case = gre_lattice_2d(epg, ntr)
    seq = []
    for op in lowering_flatten(case["seq"]):
        name = type(op).__name__
        if name == "S" or (isinstance(op, dict) and op.get("kind") == "S"):
            k = op.k if not isinstance(op, dict) else op["k"]
            seq.append(epg.S([0.5 * float(x) for x in np.asarray(k).reshape(-1)]))
        elif name == "D" or (isinstance(op, dict) and op.get("kind") == "D"):
            k = op.k if not isinstance(op, dict) else op["k"]
            seq.append(epg.D(4.0, 2e-3, k=[0.5 * float(x) for x in np.asarray(k).reshape(-1)]))
        else:
            seq.append(op)
    return dict(seq=seq, options={"kvalue": 600.0, "kgrid": 0.25})


def gre_gradient_time(epg, necho=8):
    """gradient lobes `G` (float wavenumbers, two spatial components) and time accumulation `C` (fourth coordinate) in a
    multi spin echo with imperfect refocusing, on a grid that divides k * kvalue and t * tvalue but NOT k itself (the
    reference quantises wavenumbers, i.e. shifts times ktvalue: shift.py:136-141, statematrix.py:203-211).  F0 is the
    state refocused in space AND time (shift.py:163-210): the gradient-recalled read-out between two spin echoes sees
    only the pathways whose accumulated time is zero as well."""
    T1 = np.array([600.0, 1100.0, 1600.0])
    T2 = np.array([40.0, 90.0])[None, :]
    g0 = 0.1 / (2 * np.pi * 42576.0 * 1e-3)   # gradient (mT/m) whose 1 ms lobe shifts k by 0.1 rad/m
    seq = [epg.T(90, 90)]
    for i in range(necho):
        seq += [epg.G(1.0, [2 * g0, g0]), epg.C(1.5), epg.E(4, T1, T2), epg.T(150 - 4 * i, 0), epg.G(1.0, [2 * g0, g0]), epg.C(1.5),
                epg.E(4, T1, T2), epg.ADC,
                epg.G(0.5, [-8 * g0, 4 * g0]), epg.C(0.5), epg.D(6.0, 1.5e-3), epg.G(0.5, [8 * g0, -4 * g0]), epg.C(0.5),
                epg.E(2, T1, T2), epg.Adc(phase=15.0 * i)]
    return dict(seq=seq, options={"kvalue": 2.5, "tvalue": 2.0, "kgrid": 0.25})


def lowering_flatten(seq):
    out = []
    for item in seq:
        if isinstance(item, (list, tuple)):
            out.extend(lowering_flatten(item))
        else:
            out.append(item)
    return out


CASES["slice_profile"] = slice_profile
CASES["gre_lattice_float"] = gre_lattice_float
CASES["gre_gradient_time"] = gre_gradient_time
CASES["gre_lattice_2d"] = gre_lattice_2d
CASES["gre_lattice_3d_cropped"] = gre_lattice_3d_cropped
CASES["lattice_jac"] = lattice_jac


def probe_expr(epg):
    """`probe=` expressions over F0 / Z0 superseding the in-sequence ADCs (probe.py:7-66; functions.py:118-127)"""
    case = misc_ops(epg)
    case["probe"] = ["abs(F0)", "Z0.real + 2 * F0", "F0"]
    return case


# ---- Fourier probes (probe.py:168-219): DFT / Imaging read every transverse configuration


def dft_gre_1d(epg, ntr=12):
    """RF-spoiled GRE with integer unit shifts; `DFT` probes at 21 positions after every pulse (magnetisation profile
    across the voxel) next to the plain ADC"""
    T1 = np.array([500.0, 1200.0])
    T2 = np.array([40.0, 80.0, 160.0])[None, :]
    x = np.linspace(-0.5e-3, 0.5e-3, 21)  # m
    seq = []
    for i in range(ntr):
        seq += [epg.T(30, 117 * i * (i + 1) / 2), epg.E(3, T1, T2), epg.DFT(x), epg.ADC, epg.E(7, T1, T2), epg.S(1)]
    return dict(seq=seq, options={"kvalue": 2 * np.pi * 1e3})


def imaging_gradients_2d(epg, necho=5):
    """gradient lobes in two dimensions + accumulated time; `Imaging` probes: box voxels, T2' and off-resonance
    modulation (complex), positions on a 5 x 4 grid, not reduced; a second probe reduced over everything"""
    T1 = np.array([600.0, 1100.0, 1600.0])
    T2 = np.array([40.0, 90.0])[None, :]
    g0 = 0.1 / (2 * np.pi * 42576.0 * 1e-3)
    pos = np.stack(np.meshgrid(np.linspace(-2.0, 2.0, 5), np.linspace(-1.0, 1.5, 4), indexing="ij"), axis=-1)  # (5, 4, 2), m
    seq = [epg.T(90, 90)]
    for i in range(necho):
        seq += [epg.G(1.0, [2 * g0, g0]), epg.C(1.5), epg.E(4, T1, T2), epg.T(150 - 4 * i, 0), epg.G(1.0, [2 * g0, g0]), epg.C(1.5),
                epg.E(4, T1, T2), epg.Imaging(pos, voxel_size=0.8, modulation=0.3 + 0.05j, reduce=False),
                epg.G(0.5, [-8 * g0, 4 * g0]), epg.C(0.5), epg.G(0.5, [8 * g0, -4 * g0]), epg.C(0.5),
                epg.E(2, T1, T2), epg.Imaging(pos[:, 0], voxel_shape="point", modulation=0.2)]
    return dict(seq=seq, options={"kvalue": 2.5, "tvalue": 2.0, "kgrid": 0.25})


FOURIER_CASES = {"dft_gre_1d": dft_gre_1d, "imaging_gradients_2d": imaging_gradients_2d}


def run_probes(epg, case):
    """values of every probe of the sequence, in order of appearance, as a list of arrays / lists"""
    return epg.simulate(case["seq"], asarray=False, **case["options"])


def namespace(pkg):
    """`epg`-like namespace of a reference-compatible package (+ exchange_matrix helper)"""
    import types

    core = pkg.core
    ns = types.SimpleNamespace(**{k: getattr(core, k) for k in dir(core) if not k.startswith("_")})
    ns.exchange_matrix = pkg.exchange.exchange_matrix
    import importlib

    ns.RFPulse = importlib.import_module(pkg.__name__ + ".rfpulse").RFPulse
    return ns


def run_api(epg, case):
    """run a case with a reference-compatible API (the reference itself or epgpy_b200)"""
    opts = dict(case.get("options") or {})
    if case.get("density") is not None:
        opts["init"] = epg.StateMatrix(density=case["density"])
    if case.get("init") is not None:
        opts["init"] = np.array(case["init"])
    if case.get("jac"):
        sig, jac = epg.simulate(case["seq"], probe=[None, epg.Jacobian(case["jac"])], **opts)
        return np.asarray(sig), np.asarray(jac)
    return np.asarray(epg.simulate(case["seq"], **opts)), None
