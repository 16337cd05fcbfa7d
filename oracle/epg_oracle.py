"""CPU oracle for the EPG operator-chain hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain-numpy restatement of the algorithm of the reference
(py-baudin/epgpy 3.2.dev6, read-only at /root/reference) for the path
`epg.simulate` over T / Phi / E / P / R / S / D / X / ADC plus the order-1
forward-mode derivatives.  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product package (`epgpy_b200`) never does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function of
this file against fixtures in `tests/golden/*.npz` that were produced by
importing the unmodified reference in the build container
(`tests/golden/make_golden.py`, committed), and against the closed-form
known answers of the reference's own unit tests.

Like the reference, the state is kept in FULL storage: a complex128 array
`states[*grid, 2n+1, 3]`, columns (F+, F-, Z), row n is order k=0
(reference: epgpy/statematrix.py:55,388-422).  Parameter arrays broadcast
LEFT-ALIGNED against the grid: axis 0 of a parameter is axis 0 of the grid
(reference: epgpy/common.py:273-334, `append=True`).

Every function cites the reference lines it follows.
"""

from __future__ import annotations

import numpy as np

DEG = np.pi / 180.0


# --------------------------------------------------------------------------- #
# sequence description (tiny, oracle-private; NOT the product's operator API)
# --------------------------------------------------------------------------- #


class Op(dict):
    """one step of a sequence: {'kind': str, ...parameters...}"""

    __getattr__ = dict.get

    def __matmul__(self, other):
        return combine(self, other)


def _arr(x):
    return None if x is None else np.asarray(x)


def T(alpha, phi, order1=None, order2=None, duration=0.0):
    """RF pulse (epgpy/transition.py:7-65). order1: {var: {'alpha'|'phi': coeff}};
    order2: {(var1, var2): {param: second derivative of the parameter w.r.t. the pair}} (diff.py:225-227), usually empty"""
    return Op(kind="T", alpha=_arr(alpha), phi=_arr(phi), order1=_parse_order1(order1, ("alpha", "phi")),
              order2=_parse_order2(order2), duration=duration)


def Phi(phi, order1=None, duration=0.0):
    """phase offset (epgpy/transition.py:79-108)"""
    return Op(kind="Phi", phi=_arr(phi), order1=_parse_order1(order1, ("phi",)), duration=duration)


def E(tau, T1, T2, g=0.0, order1=None, order2=None, duration=0.0):
    """relaxation/precession/recovery (epgpy/evolution.py:69-153)"""
    if duration is True:
        duration = tau
    return Op(kind="E", tau=_arr(tau), T1=_arr(T1), T2=_arr(T2), g=_arr(g),
              order1=_parse_order1(order1, ("tau", "T1", "T2", "g")), order2=_parse_order2(order2), duration=duration)


def P(tau, g, order1=None, duration=0.0):
    """precession only (epgpy/evolution.py:156-213)"""
    if duration is True:
        duration = tau
    return Op(kind="P", tau=_arr(tau), g=_arr(g), order1=_parse_order1(order1, ("tau", "g")), duration=duration)


def R(rT=0.0, rL=0.0, r0=None, order1=None, duration=0.0):
    """raw-rate evolution (epgpy/evolution.py:9-66)"""
    return Op(kind="R", rT=_arr(rT), rL=_arr(rL), r0=_arr(r0),
              order1=_parse_order1(order1, ("rT", "rL", "r0")), duration=duration)


def S(k, duration=0.0):
    """integer 1-d shift (epgpy/shift.py:14-101, shift-1d method only);
    `k` may be an int, or an integer vector: the sequence's vectors must all be
    integer multiples of one base vector (collinear), which reduces exactly to 1-d."""
    return Op(kind="S", k=k, duration=duration)


def D(tau, D, k=None, duration=0.0):
    """diffusion (epgpy/diffusion.py:14-79)"""
    if duration is True:
        duration = tau
    return Op(kind="D", tau=_arr(tau), D=_arr(D), k=k, duration=duration)


def X(tau, khi, axis=-1, T1=None, T2=None, g=None, duration=0.0):
    """exchange (epgpy/exchange.py:11-120)"""
    if duration is True:
        duration = tau
    return Op(kind="X", tau=_arr(tau), khi=khi, axis=axis, T1=T1, T2=T2, g=g, duration=duration)


def ADC(attr="F0", phase=None, reduce=None, weights=None, duration=0.0):
    """read-out (epgpy/probe.py:82-165)"""
    return Op(kind="ADC", attr=attr, phase=_arr(phase), reduce=reduce, weights=_arr(weights), duration=duration)


def SPOILER():
    """epgpy/operator.py:281-286"""
    return Op(kind="SPOILER", duration=0.0)


def RESET():
    """epgpy/operator.py:297-304"""
    return Op(kind="RESET", duration=0.0)


def PD(pd, reset=True):
    """epgpy/operator.py:315-341"""
    return Op(kind="PD", pd=_arr(pd), reset=reset, duration=0.0)


def WAIT(duration):
    """epgpy/operator.py:259-265"""
    return Op(kind="WAIT", duration=duration)


def MATRIX(mat, mat0=None, dmats=None, order1=None, duration=0.0):
    """generic 3x3 operator with affine term (epgpy/opmatrix.py:10-63)"""
    dmats = dmats or {}
    return Op(kind="MATRIX", mat=np.asarray(mat, dtype=complex), mat0=None if mat0 is None else np.asarray(mat0, dtype=complex),
              dmats=dmats, order1=_parse_order1(order1, tuple(dmats)), duration=duration)


def _parse_order2(order2):
    """{(var1, var2): {param: coeff}} with sorted pairs (epgpy/diff.py:225-227, 534-540); other spellings of the
    keyword only SELECT pairs in the reference and carry no coefficient"""
    if not isinstance(order2, dict):
        return {}
    return {tuple(sorted(pair)): dict(c) for pair, c in order2.items() if c}


def _parse_order1(order1, parameters):
    """epgpy/diff.py:153-195 (order-1 part only)"""
    if not order1:
        return {}
    if order1 is True:
        return {p: {p: 1} for p in parameters}
    if isinstance(order1, str):
        order1 = [order1]
    if isinstance(order1, (list, tuple)):
        return {p: {p: 1} for p in order1}
    if all(isinstance(v, str) for v in order1.values()):
        return {var: {order1[var]: 1} for var in order1}
    if all(isinstance(v, dict) for v in order1.values()):
        return {var: dict(order1[var]) for var in order1}
    raise ValueError(f"Invalid parameter 'order1' value: {order1}")


# --------------------------------------------------------------------------- #
# broadcasting helpers (left-aligned)
# --------------------------------------------------------------------------- #


def _shape(x):
    return () if x is None else np.shape(x)


def broadcast_left(*shapes):
    """epgpy/common.py:290-303 with append=True"""
    ndim = max([len(s) for s in shapes] + [1])
    out = [1] * ndim
    for s in shapes:
        for i, d in enumerate(s):
            if d == 1:
                continue
            if out[i] not in (1, d):
                raise ValueError(f"Incompatible shapes: {shapes}")
            out[i] = d
    return tuple(out)


def _left(x, ndim, tail=0):
    """append singleton axes so that `x` (shape = lead + tail axes) has `ndim` leading axes"""
    x = np.asarray(x)
    lead = x.ndim - tail
    if lead == 0:
        x = x.reshape((1,) + x.shape)
        lead = 1
    idx = x.shape[:lead] + (1,) * (ndim - lead) + x.shape[lead:]
    return x.reshape(idx)


def op_shape(op):
    """the operator's own parameter shape (reference: Operator.shape of each class)"""
    kind = op.kind
    if kind == "T":
        return broadcast_left(_shape(op.alpha), _shape(op.phi))
    if kind == "Phi":
        return broadcast_left(_shape(op.phi))
    if kind == "E":
        return broadcast_left(_shape(op.tau), _shape(op.T1), _shape(op.T2), _shape(op.g))
    if kind == "P":
        return broadcast_left(_shape(op.tau), _shape(op.g))
    if kind == "R":
        return broadcast_left(_shape(op.rT), _shape(op.rL), _shape(op.r0))
    if kind == "PD":
        return broadcast_left(_shape(op.pd))
    if kind == "D":
        return broadcast_left(_shape(op.tau))
    if kind == "MATRIX":
        return broadcast_left(op.mat.shape[:-2])
    if kind == "X":
        mats, ax = exchange_matrices(op)
        return tuple(d for i, d in enumerate(mats.shape[:-1]) if i != ax + 1)
    return (1,)


# --------------------------------------------------------------------------- #
# coefficient builders
# --------------------------------------------------------------------------- #


def rot_x(alpha):
    """epgpy/transition.py:120-137"""
    a = DEG * np.atleast_1d(alpha)
    m = np.empty(a.shape + (3, 3), dtype=complex)
    c2, s2, s, c = np.cos(a / 2) ** 2, np.sin(a / 2) ** 2, np.sin(a), np.cos(a)
    m[..., 0, 0], m[..., 0, 1], m[..., 0, 2] = c2, s2, -1j * s
    m[..., 1, 0], m[..., 1, 1], m[..., 1, 2] = s2, c2, 1j * s
    m[..., 2, 0], m[..., 2, 1], m[..., 2, 2] = -0.5j * s, 0.5j * s, c
    return m


def rot_x_d(alpha):
    """epgpy/transition.py:172-186 (per degree)"""
    a = DEG * np.atleast_1d(alpha)
    m = np.empty(a.shape + (3, 3), dtype=complex)
    s, c = np.sin(a), np.cos(a)
    m[..., 0, 0], m[..., 0, 1], m[..., 0, 2] = -0.5 * s, 0.5 * s, -1j * c
    m[..., 1, 0], m[..., 1, 1], m[..., 1, 2] = 0.5 * s, -0.5 * s, 1j * c
    m[..., 2, 0], m[..., 2, 1], m[..., 2, 2] = -0.5j * c, 0.5j * c, -s
    return m * DEG


def rot_z(phi):
    """epgpy/transition.py:140-151"""
    p = DEG * np.atleast_1d(phi)
    m = np.zeros(p.shape + (3, 3), dtype=complex)
    m[..., 0, 0], m[..., 1, 1], m[..., 2, 2] = np.exp(1j * p), np.exp(-1j * p), 1
    return m


def rot_z_d(phi):
    """epgpy/transition.py:189-196 (per degree)"""
    p = DEG * np.atleast_1d(phi)
    m = np.zeros(p.shape + (3, 3), dtype=complex)
    m[..., 0, 0], m[..., 1, 1] = 1j * np.exp(1j * p), -1j * np.exp(-1j * p)
    return m * DEG


def _pair(a, b):
    """left-aligned expansion of two parameter arrays (epgpy/common.py:306-334)"""
    nd = max(np.ndim(a), np.ndim(b), 1)
    return _left(a, nd), _left(b, nd)


def rf_matrix(alpha, phi):
    """T = Rz(phi) Rx(alpha) Rz(-phi)   (epgpy/transition.py:114-117)"""
    alpha, phi = _pair(alpha, phi)
    return rot_z(phi) @ rot_x(alpha) @ rot_z(-phi)


def rf_matrix_dalpha(alpha, phi):
    """epgpy/transition.py:160-162"""
    alpha, phi = _pair(alpha, phi)
    return rot_z(phi) @ rot_x_d(alpha) @ rot_z(-phi)


def rf_matrix_dphi(alpha, phi):
    """epgpy/transition.py:165-169"""
    alpha, phi = _pair(alpha, phi)
    return rot_z_d(phi) @ rot_x(alpha) @ rot_z(-phi) - rot_z(phi) @ rot_x(alpha) @ rot_z_d(-phi)


def rot_x_d2(alpha):
    """epgpy/transition.py:223-237 (per degree^2)"""
    a = DEG * np.atleast_1d(alpha)
    m = np.empty(a.shape + (3, 3), dtype=complex)
    s, c = np.sin(a), np.cos(a)
    m[..., 0, 0], m[..., 0, 1], m[..., 0, 2] = -0.5 * c, 0.5 * c, 1j * s
    m[..., 1, 0], m[..., 1, 1], m[..., 1, 2] = 0.5 * c, -0.5 * c, -1j * s
    m[..., 2, 0], m[..., 2, 1], m[..., 2, 2] = 0.5j * s, -0.5j * s, -c
    return m * DEG**2


def rot_z_d2(phi):
    """epgpy/transition.py:240-247 (per degree^2)"""
    p = DEG * np.atleast_1d(phi)
    m = np.zeros(p.shape + (3, 3), dtype=complex)
    m[..., 0, 0], m[..., 1, 1] = -np.exp(1j * p), -np.exp(-1j * p)
    return m * DEG**2


def rf_matrix_d2(alpha, phi, p1, p2):
    """second derivatives of the RF pulse (epgpy/transition.py:203-220)"""
    alpha, phi = _pair(alpha, phi)
    pair = tuple(sorted((p1, p2)))
    if pair == ("alpha", "alpha"):
        return rot_z(phi) @ rot_x_d2(alpha) @ rot_z(-phi)
    if pair == ("alpha", "phi"):
        return rot_z_d(phi) @ rot_x_d(alpha) @ rot_z(-phi) - rot_z(phi) @ rot_x_d(alpha) @ rot_z_d(-phi)
    if pair == ("phi", "phi"):
        return (rot_z_d2(phi) @ rot_x(alpha) @ rot_z(-phi) + rot_z(phi) @ rot_x(alpha) @ rot_z_d2(-phi)
                - 2 * rot_z_d(phi) @ rot_x(alpha) @ rot_z_d(-phi))
    raise ValueError(pair)


def relax_d2(p1, p2, tau, T1, T2, g=0.0):
    """second derivatives of E: (d2arr, d2arr0) or None when the pair vanishes (epgpy/evolution.py:405-488)"""
    tau, T1, T2, g = _e_params(tau, T1, T2, g)
    rT = tau * (1 / T2 + 2j * np.pi * g)
    rL = tau / T1
    pair = tuple(sorted((p1, p2)))
    arr, arr0 = evolution_arrays(rT, rL, rL)
    fM = fZ = None
    if pair == ("tau", "tau"):
        fM, fZ = (rT / tau) ** 2, 1 / T1**2
    elif pair == ("T1", "T1"):
        fZ = tau**2 / T1**4 - 2 * tau / T1**3
    elif pair == ("T2", "T2"):
        fM = tau**2 / T2**4 - 2 * tau / T2**3
    elif pair == ("g", "g"):
        fM = (-2j * np.pi * tau) ** 2
    elif pair == ("T1", "tau"):
        fZ = (1 - rL) / T1**2
    elif pair == ("T2", "tau"):
        fM = (1 - rT) / T2**2
    elif pair == ("g", "tau"):
        fM = -2j * np.pi * (1 - rT)
    elif pair == ("T2", "g"):
        fM = -2j * np.pi * (tau / T2) ** 2
    else:
        return None
    arr[..., 1] = 0 if fM is None else arr[..., 1] * fM
    arr[..., 0] = arr[..., 1].conj()
    arr[..., 2] = 0 if fZ is None else arr[..., 2] * fZ
    arr0[..., 2] = -arr[..., 2]
    return arr, (arr0 if fZ is not None else None)


def evolution_arrays(rT, rL, r0=None):
    """arr = [conj e^{-rT}, e^{-rT}, e^{-rL}], arr0 = [0,0,1-e^{-r0}]  (epgpy/evolution.py:220-242)"""
    nd = max(np.ndim(rT), np.ndim(rL), np.ndim(r0) if r0 is not None else 0, 1)
    rT, rL = _left(rT, nd), _left(rL, nd)
    shape = broadcast_left(rT.shape, rL.shape, () if r0 is None else _left(r0, nd).shape)
    arr = np.zeros(shape + (3,), dtype=complex)
    arr[..., 1] = np.exp(-rT)
    arr[..., 0] = arr[..., 1].conj()
    arr[..., 2] = np.exp(-rL)
    arr0 = None
    if r0 is not None:
        arr0 = np.zeros(shape + (3,), dtype=complex)
        arr0[..., 2] = 1 - np.exp(-_left(r0, nd))
    return arr, arr0


def _e_params(tau, T1, T2, g):
    nd = max(np.ndim(tau), np.ndim(T1), np.ndim(T2), np.ndim(g), 1)
    return tuple(_left(np.asarray(x, dtype=float), nd) for x in (tau, T1, T2, g))


def relax_arrays(tau, T1, T2, g=0.0):
    """epgpy/evolution.py:251-256"""
    tau, T1, T2, g = _e_params(tau, T1, T2, g)
    rT = tau * (1 / T2 + 2j * np.pi * g)
    rL = tau / T1
    return evolution_arrays(rT, rL, rL)


def relax_d(param, tau, T1, T2, g=0.0):
    """first derivatives of E: (darr, darr0)  (epgpy/evolution.py:360-399)"""
    tau, T1, T2, g = _e_params(tau, T1, T2, g)
    rT = tau * (1 / T2 + 2j * np.pi * g)
    rL = tau / T1
    if param == "tau":
        arr, arr0 = evolution_arrays(rT, rL, rL)
        arr[..., 1] *= -rT / tau
        arr[..., 0] = arr[..., 1].conj()
        arr[..., 2] *= -1 / T1
        arr0[..., 2] = -arr[..., 2]
        return arr, arr0
    if param == "T1":
        arr, arr0 = evolution_arrays(0 * rT, rL, rL)
        arr[..., :2] = 0
        arr[..., 2] *= tau / T1**2
        arr0[..., 2] = -arr[..., 2]
        return arr, arr0
    if param == "T2":
        arr, _ = evolution_arrays(rT, 0 * rL, None)
        arr[..., 0] *= tau / T2**2
        arr[..., 1] *= tau / T2**2
        arr[..., 2] = 0
        return arr, None
    if param == "g":
        arr, _ = evolution_arrays(rT, 0 * rL, None)
        arr[..., 1] *= -2j * np.pi * tau
        arr[..., 0] = arr[..., 1].conj()
        arr[..., 2] = 0
        return arr, None
    raise ValueError(param)


def precession_arrays(tau, g):
    """epgpy/evolution.py:245-248"""
    tau, g = _pair(np.asarray(tau, dtype=float), np.asarray(g, dtype=float))
    return evolution_arrays(2j * np.pi * g * tau, 0 * tau * g, None)


def precession_d(param, tau, g):
    """epgpy/evolution.py:313-328"""
    tau, g = _pair(np.asarray(tau, dtype=float), np.asarray(g, dtype=float))
    arr, _ = evolution_arrays(2j * np.pi * g * tau, 0 * tau * g, None)
    arr[..., 1] *= -2j * np.pi * (g if param == "tau" else tau)
    arr[..., 0] = arr[..., 1].conj()
    arr[..., 2] = 0
    return arr, None


def evolution_d(param, rT, rL, r0=None):
    """epgpy/evolution.py:263-280"""
    if param == "rT":
        arr, _ = evolution_arrays(rT, 0 * np.asarray(rL))
        arr[..., 2] = 0
        return -arr, None
    if param == "rL":
        arr, _ = evolution_arrays(0 * np.asarray(rT), rL)
        arr[..., :2] = 0
        return -arr, None
    if param == "r0":
        arr, arr0 = evolution_arrays(0 * np.asarray(rT), 0 * np.asarray(rL), r0)
        arr[:] = 0
        arr0[..., 2] -= 1
        return arr, -arr0
    raise ValueError(param)


def expm(mat):
    """matrix exponential through an eigen-decomposition (epgpy/exchange.py:262-282)"""
    nrm = np.linalg.norm(mat)
    if np.isclose(nrm, 0):
        return np.eye(mat.shape[-1]).reshape(mat.shape)
    tr = lambda m: np.moveaxis(m, -1, -2)
    if np.allclose(mat, tr(mat).conj()):
        ev, vec = np.linalg.eigh(mat / nrm)
    else:
        ev, vec = np.linalg.eig(mat / nrm)
    ex = np.expm1(ev * nrm) + 1
    return tr(np.linalg.solve(tr(vec), ex[..., None] * tr(vec)))


def kinetic_matrix(k, axis=-1, ncomp=2, densities=None):
    """scalar exchange rate -> N x N kinetic matrix (epgpy/exchange.py:127-151)"""
    k = np.asarray(k, dtype=float)
    if np.any(k < 0):
        raise ValueError("Cannot have negative echange rate")
    if axis > k.ndim:
        k = np.expand_dims(k, tuple(range(k.ndim, axis)))
    axis = (k.ndim + axis + 1) if axis < 0 else axis
    kron = np.eye(ncomp) + (np.eye(ncomp) - 1) / (ncomp - 1)
    if densities is not None:
        kron = kron / densities
    return np.moveaxis(k[..., None, None] * kron, -2, axis)


def exchange_matrices(op):
    """mat[..., N(axis), N, ..., 3] = [mT, conj mT, mL]   (epgpy/exchange.py:14-66,154-203)"""
    khi, axis = op.khi, op.axis
    if np.ndim(khi) == 0:
        khi = kinetic_matrix(khi, axis=axis, ncomp=2)
    khi = np.asarray(khi, dtype=float)
    axis = int(khi.ndim + axis - 1) if axis < 0 else int(axis)
    tau = np.asarray(op.tau, dtype=float)
    T1 = np.asarray(np.inf if op.T1 is None else op.T1, dtype=float)
    T2 = np.asarray(np.inf if op.T2 is None else op.T2, dtype=float)
    g = np.asarray(0.0 if op.g is None else op.g, dtype=float)
    ncomp = khi.shape[-1]
    eye = np.eye(ncomp)
    minshape = khi.shape[:-1]
    shape = np.broadcast_shapes(*[s[::-1] for s in (tau.shape, T1.shape, T2.shape, g.shape, minshape)])[::-1]
    nd = len(shape)
    tau, T1, T2, g = [np.expand_dims(a, tuple(range(a.ndim, nd))) for a in (tau, T1, T2, g)]
    T1, T2, g = [np.broadcast_to(a, shape) for a in (T1, T2, g)]
    khi = np.expand_dims(khi, tuple(range(nd - len(minshape))))
    tau, T1, T2, g = [np.moveaxis(a, axis, -1) for a in (tau, T1, T2, g)]
    xT = -khi + (-1 / T2 + 2j * np.pi * g)[..., None] * eye
    xL = -khi + (-1 / T1)[..., None] * eye
    mT = expm(xT * tau[..., None])
    mL = expm(xL * tau[..., None])
    mT = np.moveaxis(mT, (-2, -1), (axis, axis + 1))
    mL = np.moveaxis(mL, (-2, -1), (axis, axis + 1))
    return np.stack([mT, mT.conj(), mL], axis=-1), axis


# --------------------------------------------------------------------------- #
# state primitives (full storage)
# --------------------------------------------------------------------------- #


def new_states(grid, init=None, density=1.0):
    """initial states / equilibrium  (epgpy/statematrix.py:12-80,379-422; functions.py:133-141)"""
    nd = len(grid)
    dens = np.broadcast_to(_left(np.asarray(density, dtype=float), nd), grid)
    eq = np.zeros(tuple(grid) + (1, 3), dtype=complex)
    eq[..., 0, 2] = dens
    if init is None:
        st = eq.copy()
    else:
        init = np.asarray(init, dtype=complex)
        if init.ndim == 1:
            init = init.reshape(1, 3)
        if init.ndim == 2:
            if init.shape[0] % 2 != 1 or init.shape[1] != 3:
                raise ValueError("init must be (2n+1) x 3")
            st = np.broadcast_to(init, tuple(grid) + init.shape).copy()
        else:
            st = np.broadcast_to(_left(init, nd, tail=2), tuple(grid) + init.shape[-2:]).copy()
        n = (st.shape[-2] - 1) // 2
        eq = resize(eq, n)
    return st, eq


def nstate(states):
    return (states.shape[-2] - 1) // 2


def resize(states, n):
    """symmetric zero-pad / crop to n orders (epgpy/statematrix.py:293-297,793-804)"""
    cur = nstate(states)
    if n == cur:
        return states
    if n > cur:
        pad = [(0, 0)] * (states.ndim - 2) + [(n - cur, n - cur), (0, 0)]
        return np.pad(states, pad)
    d = cur - n
    return states[..., d:-d, :].copy()


def apply_matrix(states, mat, eq=None, mat0=None, inplace=False):
    """states[g,s,:] <- M[g] . states[g,s,:] (+ mat0 . equilibrium)   (epgpy/opmatrix.py:199-221)
    Written as the row-vector product states[g] (s x 3) @ M[g]^T (3 x 3): one small GEMM per atom instead of
    one 3x3 matvec per order.  inplace=True overwrites `states` like the reference (`out=states`, opmatrix.py:208-221)."""
    nd = states.ndim - 2
    mt = np.swapaxes(_left(mat, nd, tail=2), -1, -2)
    if inplace and np.broadcast_shapes(mt.shape[:-2], states.shape[:-2]) == states.shape[:-2]:
        out = np.matmul(states, mt, out=states)
    else:
        out = np.matmul(states, mt)
    if mat0 is not None:
        m0t = np.swapaxes(_left(mat0, nd, tail=2), -1, -2)
        n = nstate(states)
        # the equilibrium is non-zero at k = 0 only (statematrix.py:94-100): one row instead of the whole state
        out[..., n, :] += np.matmul(np.broadcast_to(eq, states.shape)[..., n:n + 1, :], m0t)[..., 0, :]
    return out


def apply_diag(states, arr, eq=None, arr0=None, inplace=False):
    """states *= arr ; states += arr0 * equilibrium   (epgpy/opscalar.py:213-232)
    inplace=True multiplies in place like the reference (`states *= arr`, opscalar.py:222-232)."""
    nd = states.ndim - 2
    a = _left(arr, nd, tail=1)[..., None, :]
    if inplace and np.broadcast_shapes(a.shape, states.shape) == states.shape:
        states *= a
        out = states
    else:
        out = states * a
    if arr0 is not None:
        n = nstate(out)
        neq = nstate(eq)
        # the equilibrium is non-zero at k = 0 only: add its row instead of a whole-state pass
        out[..., n, :] += _left(arr0, nd, tail=1) * eq[..., neq, :]
    return out


def shift_int(states, k, inplace=False):
    """1-d integer shift on an already resized array (epgpy/shift.py:283-292); the reference shifts in place
    (overlapping slice assignments, which numpy buffers)"""
    out = states if inplace else states.copy()
    n = k
    if n > 0:
        out[..., n:, 0] = states[..., :-n, 0]
        out[..., :-n, 1] = states[..., n:, 1]
        out[..., :n, 0] = 0
        out[..., -n:, 1] = 0
    else:
        out[..., :n, 0] = states[..., -n:, 0]
        out[..., -n:, 1] = states[..., :n, 1]
        out[..., n:, 0] = 0
        out[..., :-n, 1] = 0
    return out


def shift_nd(states, coords, kvec, nmax=None, tol=1e-8):
    """general integer n-d shift on a lattice of configurations (epgpy/shift.py:297-364, not in place):
    F+ moves from k to k + dk, Z stays, F-(k) = conj F+(-k) is rebuilt from the symmetry; the union of old and moved
    points is re-sorted (symmetric order: point j mirrors point N - 1 - j), cropped at nmax and pruned of rows that are
    zero within `tol` for every atom.  states [..., N, 3], coords [N, d] -> (states', coords')"""
    coords = np.asarray(coords, dtype=int)
    kvec = np.asarray(kvec, dtype=int).reshape(1, -1)
    n1 = coords.shape[0]
    uniq, inv = np.unique(np.concatenate([coords, coords + kvec, coords - kvec]), axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    idx_l, idx_t = inv[:n1], inv[n1:2 * n1]
    keep_l = keep_t = slice(None)
    if nmax is not None:
        keep = np.all(np.abs(uniq) <= nmax, axis=-1)
        if not np.all(keep):
            uniq = uniq[keep]
            remap = -np.ones(keep.size, dtype=int)
            remap[keep] = np.arange(uniq.shape[0])
            idx_t, idx_l = remap[idx_t], remap[idx_l]
            keep_t, keep_l = idx_t >= 0, idx_l >= 0
    new = np.zeros(states.shape[:-2] + (uniq.shape[0], 3), dtype=states.dtype)
    new[..., idx_l[keep_l], 2] = states[..., keep_l, 2]
    new[..., idx_t[keep_t], 0] = states[..., keep_t, 0]
    new[..., 1] = new[..., ::-1, 0].conj()
    nonzero = ~np.all(np.isclose(new, 0, atol=tol), axis=tuple(range(new.ndim - 2)) + (-1,))
    nonzero[(uniq.shape[0] - 1) // 2] = True
    return new[..., nonzero, :], uniq[nonzero]


def bmatrix(tau, k1, k2=None):
    """b-matrix of a linear change k1 -> k2 (epgpy/diffusion.py:86-123). tau in ms, k in rad/m"""
    outer = lambda a, b: a[..., None] * b[..., None, :]
    tau = np.asarray(tau, dtype=float) * 1e-3
    k1 = np.atleast_2d(k1) * 1e-3
    b = outer(k1, k1) * tau
    if k2 is None:
        return b
    k2 = np.atleast_2d(k2) * 1e-3
    kd = k2 - k1
    if np.allclose(kd, 0):
        return b
    return b + tau * (0.5 * outer(k1, kd) + 0.5 * outer(kd, k1) + outer(kd, kd) / 3)


def diffusion_factors(bL, bT, Dcoef):
    """exp(-Tr(b D))   (epgpy/diffusion.py:126-147)"""
    if np.ndim(Dcoef) == 0:
        i = np.arange(bT.shape[-1])
        return np.exp(-bL[..., i, i].sum(-1) * Dcoef), np.exp(-bT[..., i, i].sum(-1) * Dcoef)
    Dcoef = np.asarray(Dcoef, dtype=float)
    return np.exp(-(bL * Dcoef).sum((-2, -1))), np.exp(-(bT * Dcoef).sum((-2, -1)))


# --------------------------------------------------------------------------- #
# driver
# --------------------------------------------------------------------------- #


def flatten(seq):
    """epgpy/functions.py:355-369"""
    if isinstance(seq, Op):
        return [seq]
    out = []
    for item in seq:
        if isinstance(item, Op):
            out.append(item)
        elif isinstance(item, (list, tuple)):
            out.extend(flatten(item))
        else:
            raise ValueError(f"Invalid operator: {item}")
    return out


def get_shape(seq):
    """epgpy/functions.py:14-17"""
    return broadcast_left(*[op_shape(op) for op in flatten(seq)])


def adc_times(seq):
    """epgpy/functions.py:38-47"""
    t, out = 0, []
    for op in flatten(seq):
        t = t + op.duration
        if op.kind == "ADC":
            out.append(t)
    return out


class _Sim:
    """mutable simulation context: base state + order-1 partial states"""

    def __init__(self, grid, init, density, max_nstate, kvalue, kvec):
        self.grid = tuple(grid)
        self.states, self.eq = new_states(grid, init, density)
        self.partials = {}
        self.partials2 = {}  # (var1, var2) sorted -> second-order partial state
        self.pairs = []      # the pairs to propagate (those the Hessian asks for)
        self.max_nstate = max_nstate
        self.kvalue = kvalue
        self.kvec = kvec  # base shift vector (collinear n-d shifts) or None
        self.coords = None  # [N, d] integer lattice points of the stored rows (general n-d shifts), else None
        self.kgrid = None   # float shifts (shift-merge, shift.py:119-145) on multiples of this grid: lattice units
        self.tvalue = 1.0   # scale of the fourth (accumulated time) coordinate (statematrix.py:203-211)

    # -- helpers
    def all_states(self):
        yield None, self.states
        for v in self.partials:
            yield v, self.partials[v]
        for pair in self.partials2:
            yield pair, self.partials2[pair]

    def set(self, v, s):
        if v is None:
            self.states = s
        elif isinstance(v, tuple):
            self.partials2[v] = s
        else:
            self.partials[v] = s

    def wavenumbers(self):
        """k of every stored order, rad/m (epgpy/statematrix.py:177-186)"""
        if self.coords is not None:
            c3 = self.coords[:, :3].astype(float)  # (a fourth coordinate is accumulated time: statematrix.py:177-186)
            # integer lattice: shift counts x kvalue; float lattice: grid units of the WAVENUMBERS k * kvalue
            return c3 * self.kvalue if self.kgrid is None else c3 * self.kgrid
        n = nstate(self.states)
        m = np.arange(-n, n + 1, dtype=float)
        if self.kvec is None:
            return m[:, None] * self.kvalue
        return m[:, None] * np.asarray(self.kvec, dtype=float)[None, :] * self.kvalue


def _linear_coeffs(op):
    """(kind, arr|mat, arr0|mat0, {param: (d, d0)}) of a differentiable operator"""
    k = op.kind
    if k == "T":
        d = {}
        params = {p for v in op.order1 for p in op.order1[v]}
        if "alpha" in params:
            d["alpha"] = (rf_matrix_dalpha(op.alpha, op.phi), None)
        if "phi" in params:
            d["phi"] = (rf_matrix_dphi(op.alpha, op.phi), None)
        return "mat", rf_matrix(op.alpha, op.phi), None, d
    if k == "Phi":
        d = {"phi": (rot_z_d(op.phi), None)} if op.order1 else {}
        return "mat", rot_z(op.phi), None, d
    if k == "E":
        arr, arr0 = relax_arrays(op.tau, op.T1, op.T2, op.g)
        params = {p for v in op.order1 for p in op.order1[v]}
        return "diag", arr, arr0, {p: relax_d(p, op.tau, op.T1, op.T2, op.g) for p in params}
    if k == "P":
        arr, arr0 = precession_arrays(op.tau, op.g)
        params = {p for v in op.order1 for p in op.order1[v]}
        return "diag", arr, arr0, {p: precession_d(p, op.tau, op.g) for p in params}
    if k == "R":
        arr, arr0 = evolution_arrays(op.rT, op.rL, op.r0)
        params = {p for v in op.order1 for p in op.order1[v]}
        return "diag", arr, arr0, {p: evolution_d(p, op.rT, op.rL, op.r0) for p in params}
    if k == "MATRIX":
        return "mat", op.mat, op.mat0, dict(op.dmats)
    raise ValueError(k)


def _second_coeffs(op, p1, p2):
    """(d2, d2_0) of a differentiable operator w.r.t. a pair of its parameters, None if it vanishes"""
    if op.kind == "T":
        return rf_matrix_d2(op.alpha, op.phi, p1, p2), None
    if op.kind == "E":
        return relax_d2(p1, p2, op.tau, op.T1, op.T2, op.g)
    raise NotImplementedError(f"second derivatives of {op.kind}")


def combine(op1, op2):
    """`op1 @ op2`: one 3x3 operator equal to op1 followed by op2, affine terms included
    (epgpy/operator.py:219-241, opmatrix.py:89-135,173-187; forward part only)"""

    def as_mats(op):
        form, c, c0, _ = _linear_coeffs(op)
        if form == "diag":
            c = c[..., None] * np.eye(3)
            c0 = None if c0 is None else c0[..., None] * np.eye(3)
        return c, c0

    (m1, m01), (m2, m02) = as_mats(op1), as_mats(op2)
    nd = max(m1.ndim, m2.ndim) - 2
    m1, m2 = _left(m1, nd, tail=2), _left(m2, nd, tail=2)
    mat = m2 @ m1
    mat0 = None
    if m01 is not None:
        mat0 = m2 @ _left(m01, nd, tail=2)
    if m02 is not None:
        m02 = np.broadcast_to(_left(m02, nd, tail=2), mat.shape)
        mat0 = m02.copy() if mat0 is None else mat0 + m02
    return MATRIX(mat, mat0, duration=op1.duration + op2.duration)


def _apply_linear(sim, op):
    """differentiable operators: epgpy/diff.py:119-139,264-288 (order 1)"""
    form, c, c0, dcoef = _linear_coeffs(op)
    app = apply_matrix if form == "mat" else apply_diag
    nd = len(sim.grid)
    base = sim.states

    def scaled(part, coeff):
        coeff = np.asarray(coeff)
        if coeff.ndim:
            coeff = _left(coeff, nd)[..., None, None]
        return part * coeff

    # 0. second-order partials (epgpy/diff.py:290-378 with every cross term, i.e. `auto_cross_derivatives`), from the
    #    PRE-operator base and first-order states:
    #    x_ab <- Op x_ab + sum_p c_ap dOp_p x_b + sum_p c_bp dOp_p x_a + sum_pq c_ap c_bq d2Op_pq x_0 + sum_p c2_ab,p dOp_p x_0
    for a, b in sim.pairs:
        acc = None
        for src, var in ((b, a), (a, b)):  # a == b: both terms, hence the factor 2 of diff.py:349-362
            if var in op.order1 and src in sim.partials:
                for p_, coeff in op.order1[var].items():
                    part = scaled(app(sim.partials[src], dcoef[p_][0]), coeff)
                    acc = part if acc is None else acc + part
        if a in op.order1 and b in op.order1:
            for p_, ca in op.order1[a].items():
                for q_, cb in op.order1[b].items():
                    d2 = _second_coeffs(op, p_, q_)
                    if d2 is None:
                        continue
                    part = scaled(scaled(app(base, d2[0], sim.eq, d2[1]), ca), cb)
                    acc = part if acc is None else acc + part
        for p_, c2 in (op.order2 or {}).get((a, b), {}).items():
            d, d0 = dcoef[p_] if p_ in dcoef else _linear_coeffs(Op(op, order1={"_": {p_: 1}}))[3][p_]
            part = scaled(app(base, d, sim.eq, d0), c2)
            acc = part if acc is None else acc + part
        if (a, b) in sim.partials2:
            prev = app(sim.partials2[(a, b)], c)
            sim.partials2[(a, b)] = prev if acc is None else prev + np.broadcast_to(acc, prev.shape)
        elif acc is not None:
            sim.partials2[(a, b)] = np.broadcast_to(acc, base.shape).copy()
    # 1. propagate the existing partials (their equilibrium is zero: no affine term)
    for v in list(sim.partials):
        sim.partials[v] = app(sim.partials[v], c, inplace=True)
    # 2. new contributions from the pre-operator base state (diff.py:279-286,556-579)
    for v, pc in op.order1.items():
        acc = None
        for p, coeff in pc.items():
            d, d0 = dcoef[p]
            part = app(base, d, sim.eq, d0)
            coeff = np.asarray(coeff)
            if coeff.ndim:
                coeff = _left(coeff, nd)[..., None, None]
            part = part * coeff
            acc = part if acc is None else acc + part
        if acc is None:
            continue
        acc = np.broadcast_to(acc, base.shape)
        sim.partials[v] = sim.partials[v] + acc if v in sim.partials else acc.copy()
    # 3. the operator itself (in place, like the reference: opscalar.py:222-232, opmatrix.py:208-221)
    sim.states = app(base, c, sim.eq, c0, inplace=True)


def _apply_shift(sim, op):
    """epgpy/shift.py:82-101"""
    k = op.k
    if sim.coords is not None or sim.kgrid is not None or (not isinstance(k, (int, np.integer)) and sim.kvec is None):
        # general integer lattice (the reference's `shift-nd` method, epgpy/shift.py:103-117): every state set moves on
        # the same lattice; rows are pruned only where ALL sets are empty so that they keep one common row order
        kv = np.atleast_1d(np.asarray(k)).reshape(-1)
        if sim.kgrid is not None:
            # the quantisation of shift.py:401-404 acts on wavenumbers = shifts x ktvalue (shift.py:136-141,
            # statematrix.py:203-211): (kvalue, kvalue, kvalue, tvalue); exact for multiples of the grid
            ktv = np.asarray(([float(sim.kvalue)] * 3 + [float(sim.tvalue)])[:len(kv)] if len(kv) == 4 else [float(sim.kvalue)] * len(kv))
            q = kv * ktv / sim.kgrid
            if not np.allclose(q, np.round(q), atol=1e-6):
                raise NotImplementedError("float shifts off the grid merge states approximately: outside the hot path")
            kv = np.round(q).astype(int)
        if sim.coords is None:
            n = nstate(sim.states)
            sim.coords = np.zeros((2 * n + 1, len(kv)), dtype=int)
            sim.coords[:, 0] = np.arange(-n, n + 1)
        if len(kv) < sim.coords.shape[1]:
            kv = np.pad(kv, (0, sim.coords.shape[1] - len(kv)))
        elif len(kv) > sim.coords.shape[1]:
            sim.coords = np.pad(sim.coords, [(0, 0), (0, len(kv) - sim.coords.shape[1])])
        nmax = (sim.max_nstate or None) if sim.kgrid is None else None  # (shift-merge does not crop)
        keys = [key for key, _ in sim.all_states()]
        stacked = np.stack([st for _, st in sim.all_states()], axis=0)  # [sets, *grid, N, 3]
        stacked, coords = shift_nd(stacked, sim.coords, kv, nmax=nmax, tol=1e-30)
        for key, st in zip(keys, stacked):
            sim.set(key, np.array(st))
        eq = np.zeros(sim.states.shape, dtype=complex)
        eq[..., (coords.shape[0] - 1) // 2, 2] = sim.eq[..., nstate(sim.eq), 2]
        sim.eq, sim.coords = eq, coords
        return
    if not isinstance(k, (int, np.integer)):
        kv = np.asarray(k).reshape(-1)
        if sim.kvec is None:
            raise ValueError("vector shift without base vector")
        base = np.asarray(sim.kvec)
        i = int(np.argmax(np.abs(base)))
        m = kv[i] / base[i]
        if abs(m - round(m)) > 1e-12 or not np.allclose(kv, round(m) * base):
            raise NotImplementedError("non-collinear n-d shifts are outside the hot path")
        k = int(round(m))
    k = int(k)
    n = nstate(sim.states)
    nmax = sim.max_nstate or None
    n_new = n + abs(k) if nmax is None else min(n + abs(k), nmax)
    sim.eq = resize(sim.eq, n_new)
    for v, s in list(sim.all_states()):
        s = resize(s, n_new)
        if abs(k) > n_new * 2:  # everything falls off
            s = s.copy()
            s[..., :2] = 0
        else:
            s = shift_int(s, k, inplace=s.flags.writeable)
        sim.set(v, s)


def _apply_diffusion(sim, op, propagate):
    """epgpy/diffusion.py:60-79"""
    kk = sim.wavenumbers()
    tau = _left(np.asarray(op.tau, dtype=float), len(sim.grid))[..., None, None, None]
    bL = bmatrix(tau, kk)
    if op.k is None:
        bT = bL
    else:
        if isinstance(op.k, (int, np.integer)):
            sh = np.array([op.k], dtype=float) if sim.kvec is None else None
        else:
            sh = np.asarray(op.k, dtype=float).reshape(-1)
        if sh is None:
            raise ValueError("scalar D.k with vector shifts")
        if sim.coords is not None and len(sh) < min(sim.coords.shape[1], 3):
            sh = np.pad(sh, (0, min(sim.coords.shape[1], 3) - len(sh)))
        sh = sh * sim.kvalue
        bT = bmatrix(tau, kk - sh, kk)
    DL, DT = diffusion_factors(bL, bT, op.D if np.ndim(op.D) else float(op.D))
    for v, s in list(sim.all_states()):
        if v is not None and not propagate:
            continue
        s = s.copy()
        s[..., 0] = DT * s[..., 0]
        s[..., 2] = DL * s[..., 2]
        s[..., 1] = s[..., ::-1, 0].conj()
        sim.set(v, s)


def _apply_exchange(sim, op, propagate):
    """epgpy/exchange.py:89-120"""
    mats, ax = exchange_matrices(op)
    nd = len(sim.grid)
    n = mats.shape[ax]
    khi = op.khi if np.ndim(op.khi) else kinetic_matrix(op.khi, axis=op.axis, ncomp=2)
    dens = sim.eq[..., nstate(sim.eq), 2].real
    chk = np.einsum("...i,...i->...", np.moveaxis(np.asarray(khi, dtype=float), -1, -1),
                    np.moveaxis(np.broadcast_to(dens, sim.grid)[..., None], ax, -1))
    if not np.allclose(chk, 0):
        raise RuntimeError("Exchange matrix `khi` does not conserve total magnetization")
    # mats: lead axes (op's own, with the 2 pool axes at ax, ax+1), last axis = component
    lead = mats.ndim - 1
    m = mats.reshape(mats.shape[:lead] + (1,) * (nd + 1 - lead) + (1, 3)) if lead < nd + 1 else mats[..., None, :]
    for v, s in list(sim.all_states()):
        if v is not None and not propagate:
            continue
        eq = sim.eq if v is None else 0 * sim.eq
        d = np.expand_dims(s - eq, ax)  # insert destination-pool axis: source pools now at ax+1
        out = (m * d).sum(axis=ax + 1)
        sim.set(v, out + eq)


def simulate(seq, *, init=None, density=1.0, max_nstate=None, kvalue=1.0, kvec=None,
             jacobian=None, jacobian_probe=None, hessian=None, propagate_nondiff=False, adc_time=False, grid=None, kgrid=None,
             tvalue=1.0):
    """forward simulation: values (nADC, *grid) complex128   (epgpy/functions.py:50-192)

    jacobian: list of variable names -> also returns (nADC, *grid, nvars)  (epgpy/diff.py:384-416)
    hessian: (variables1, variables2) -> also returns (nADC, *grid, nvars1, nvars2)  (epgpy/diff.py:419-476)
    propagate_nondiff: False reproduces the reference, whose D / X / SPOILER / RESET / PD never
        touch the partial states (operator.py:96-104); True applies them to the partials too
        (the mathematically correct chain rule, which the CUDA path implements).
    """
    seq = flatten(seq)
    if not any(op.kind == "ADC" for op in seq):
        raise ValueError("Cannot simulate sequence without at least one Probe/ADC operator")
    shape = get_shape(seq)
    if grid is not None:
        shape = broadcast_left(shape, tuple(grid))
    if init is not None and np.ndim(init) > 2:
        shape = broadcast_left(shape, np.shape(init)[:-2])
    sim = _Sim(shape, init, density, max_nstate, kvalue, kvec)
    sim.kgrid, sim.tvalue = kgrid, tvalue
    nd = len(shape)
    if hessian is not None:
        v1s, v2s = hessian
        sim.pairs = sorted({tuple(sorted((a, b))) for a in v1s for b in v2s if "magnitude" not in (a, b)})
    values, jacs, hesss, times, tic = [], [], [], [], 0
    for op in seq:
        kind = op.kind
        if kind in ("T", "Phi", "E", "P", "R", "MATRIX"):
            _apply_linear(sim, op)
        elif kind == "S":
            _apply_shift(sim, op)
        elif kind == "D":
            _apply_diffusion(sim, op, propagate_nondiff)
        elif kind == "X":
            _apply_exchange(sim, op, propagate_nondiff)
        elif kind == "SPOILER":
            for v, s in list(sim.all_states()):
                if v is None or propagate_nondiff:
                    s = s.copy()
                    s[..., :2] = 0
                    sim.set(v, s)
        elif kind == "RESET":
            if sim.coords is not None:  # back to the single lattice point k = 0
                c = (sim.coords.shape[0] - 1) // 2
                sim.eq = sim.eq[..., c:c + 1, :].copy()
                sim.coords = np.zeros((1, sim.coords.shape[1]), dtype=int)
            sim.eq = resize(sim.eq, 0)
            sim.states = np.broadcast_to(sim.eq, sim.grid + (1, 3)).copy()
            for v in list(sim.partials):
                sim.partials[v] = resize(sim.partials[v], 0) if not propagate_nondiff else 0 * sim.states
            for v in list(sim.partials2):
                sim.partials2[v] = resize(sim.partials2[v], 0) if not propagate_nondiff else 0 * sim.states
        elif kind == "PD":
            n = nstate(sim.states)
            eq = np.zeros(sim.grid + (2 * n + 1, 3), dtype=complex)
            eq[..., n, 2] = np.broadcast_to(_left(np.asarray(op.pd, dtype=float), nd), sim.grid)
            sim.eq = eq
            if op.reset:
                sim.states = eq.copy()
                if propagate_nondiff:
                    for v in list(sim.partials):
                        sim.partials[v] = 0 * eq
        elif kind in ("WAIT", "ADC"):
            pass
        else:
            raise ValueError(f"unknown op kind {kind}")
        tic = tic + op.duration
        if kind == "ADC":
            values.append(_acquire(sim.states, op, op.attr, sim))
            times.append(tic)
            if jacobian is not None:
                attr = jacobian_probe or "F0"
                cols = []
                for var in jacobian:
                    if var == "magnitude":
                        cols.append(_acquire(sim.states, op, attr))
                    elif var in sim.partials:
                        cols.append(_acquire(sim.partials[var], op, attr))
                    else:
                        cols.append(np.zeros(sim.grid, dtype=complex) * _acquire(sim.states, op, attr))
                jacs.append(np.stack(np.broadcast_arrays(*cols), axis=-1))
            if hessian is not None:
                attr = jacobian_probe or "F0"
                zero = np.zeros(sim.grid, dtype=complex)
                rows = []
                for a in hessian[0]:
                    cols = []
                    for b in hessian[1]:
                        if a == "magnitude" or b == "magnitude":  # first derivatives w.r.t. the other one (diff.py:451-464)
                            v = b if a == "magnitude" else a
                            st = sim.partials.get(v)
                        else:
                            st = sim.partials2.get(tuple(sorted((a, b))))
                        cols.append(zero if st is None else np.broadcast_to(_acquire(st, op, attr), sim.grid))
                    rows.append(np.stack(cols, axis=-1))
                hesss.append(np.stack(rows, axis=-2))
    out = np.asarray(values)
    res = (out,) if jacobian is None else (out, np.asarray(jacs))
    if hessian is not None:
        res = res + (np.asarray(hesss),)
    if adc_time:
        res = (np.asarray(times),) + res
    return res[0] if len(res) == 1 else res


def _acquire(states, op, attr, sim=None):
    """epgpy/probe.py:138-165, statematrix.py:148-175"""
    n = nstate(states)
    col = {"F0": 0, "Z0": 2}[attr]
    if sim is not None and sim.coords is not None and sim.coords.shape[1] == 4:
        # accumulated-time coordinate (statematrix.py:136-156): F0 sums exp(-|t|) F over the rows with zero wavenumber
        if attr != "F0":
            raise NotImplementedError("Z0 with accumulated-time coordinates")
        i0 = np.all(sim.coords[:, :3] == 0, axis=-1)
        t = sim.coords[:, 3].astype(float) * (sim.tvalue if sim.kgrid is None else np.broadcast_to(np.asarray(sim.kgrid, dtype=float), (4,))[3])
        arr = (states[..., 0] * (i0 * np.exp(-np.abs(t)))).sum(axis=-1)
    else:
        arr = states[..., n, col]
    if op.weights is not None:
        w = op.weights
        if w.size > 1 and w.ndim < arr.ndim:
            w = np.expand_dims(w, tuple(range(w.ndim, arr.ndim)))
        arr = arr * w
    red = op.reduce
    if op.weights is not None and red is None:
        red = tuple(range(max(op.weights.ndim, 1)))
    if red is True:
        arr = arr.sum()
    elif red is not None and red is not False:
        arr = arr.sum(axis=(red,) if isinstance(red, int) else tuple(red))
    arr = np.array(arr, copy=True)
    if op.phase is not None:
        ph = np.exp(1j * op.phase * DEG)
        if ph.size > 1 and ph.ndim < arr.ndim:
            ph = np.expand_dims(ph, tuple(range(ph.ndim, arr.ndim)))
        arr = arr * ph
    return arr
