"""Host-side mirror of the reference's operator API (epgpy/operators.py:1-25, epgpy/core.py:80-83).

Same constructors, attributes, shapes and error behaviour as the reference's operators, but an
operator here is only a DESCRIPTION: it stores its parameters and knows how to emit compact
coefficient blocks for the op-tape (`_form`, `_dform`).  Nothing is applied on the host -- the
arithmetic of the reference's `_apply` bodies (opmatrix.py:199-221, opscalar.py:213-232,
shift.py:271-294, diffusion.py:60-79, exchange.py:89-120) runs in the sm_100a kernels of
csrc/ through lowering.py.

Coefficient "forms" (see include/epgx.h for the device-side layout); every block is a float64 array
`lead_shape + (entry,)` where lead_shape is the operator's own, un-broadcast, left-aligned shape:
  ('tgen',  [a, w, B.re, B.im, U.re, U.im])                        RF pulse in rotated-Rx form
  ('e',     [e1, r0], [e2], [cos, sin] | None)                     relaxation / precession / recovery
  ('diag',  [aP.re, aP.im, aM.re, aM.im, aZ.re, aZ.im, a0.re, a0.im])   generic diagonal + affine
  ('matrix',[18 reals], [6 reals] | None)                          generic 3x3 + affine
"""

import numpy as np

from .utils import get_wavenumber

from . import common
from .common import DEG, asparam, expand_left, get_shape, isscalar, op_shape


# --------------------------------------------------------------------------------------------- #
# base classes (epgpy/operator.py)
# --------------------------------------------------------------------------------------------- #


class Operator:
    """base operator (epgpy/operator.py:13-115)"""

    def __init__(self, *, name=None, duration=None):
        if duration is None:
            duration = 0
        elif np.any(np.asarray(duration) < 0):
            raise ValueError("Cannot have duration < 0")
        self.duration = duration
        self.name = name if name else type(self).__name__

    @property
    def shape(self):
        return (1,)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def nshift(self):
        return 0

    def __repr__(self):
        return self.name

    def __mul__(self, other):
        return MultiOperator([self, other])

    def __call__(self, sm, *, inplace=False):
        """apply to a StateMatrix (epgpy/operator.py:96-104): a one-operator tape on the engine whose final state is
        read back (functions.apply_operators); inplace=True updates `sm` itself like the reference"""
        from .functions import apply_operators

        new = apply_operators([self], sm)
        if inplace:
            sm._states, sm._density = new._states, new._density
            return sm
        return new

    def copy(self, name=None, duration=None):
        import copy as _copy

        new = _copy.copy(self)
        new.name = name or self.name
        new.duration = duration or self.duration
        return new


class MultiOperator(Operator):
    """an operator made of a sequence of operators (epgpy/operator.py:118-203)"""

    def __init__(self, operators=None, *, name=None, duration=None):
        self._nshift = 0
        self._shape = (1,)
        self.operators = []
        self.duration = 0
        operators = [] if not operators else list(operators)
        for op in operators:
            self.append(op)
        if not name:
            name = " | ".join(op.name for op in operators)
        if duration is None:
            duration = self.duration
        super().__init__(name=name, duration=duration)

    @property
    def shape(self):
        return self._shape

    @property
    def nshift(self):
        return self._nshift

    def __iter__(self):
        return iter(self.operators)

    def __len__(self):
        return len(self.operators)

    def __getitem__(self, i):
        return self.operators[i]

    def __mul__(self, other):
        self.append(other)
        return self

    def append(self, op):
        if not isinstance(op, Operator):
            raise TypeError("Invalid operator: %s" % str(op))
        shape = common.broadcast_shapes(self.shape, op.shape, append=True)
        if isinstance(op, MultiOperator):
            self.operators.extend(op.operators)
        else:
            self.operators.append(op)
        self._shape = shape
        self._nshift += op.nshift
        self.duration = self.duration + op.duration


class EmptyOperator(Operator):
    """does nothing (epgpy/operator.py:248-252)"""


NULL = EmptyOperator(name="NULL")


class Wait(EmptyOperator):
    """empty operator with a duration (epgpy/operator.py:259-265)"""

    def __init__(self, duration, name=None):
        super().__init__(duration=duration, name=name if name is not None else f"Wait({duration})")


class Offset(EmptyOperator):
    """empty operator with a possibly negative duration (epgpy/operator.py:268-274)"""

    def __init__(self, duration, name=None):
        super().__init__(duration=abs(duration), name=name if name is not None else f"Offset({duration})")
        self.duration = duration


class Spoiler(Operator):
    """perfect spoiler: F+ = F- = 0 (epgpy/operator.py:281-286)"""


SPOILER = Spoiler(name="Spoiler")


class Reset(Operator):
    """return to equilibrium, order 0 (epgpy/operator.py:297-304)"""


RESET = Reset(name="Reset")


class PD(Operator):
    """set the proton density / equilibrium (epgpy/operator.py:315-341)"""

    def __init__(self, pd, *, reset=True, name=None, **kwargs):
        self.pd = asparam(pd)
        self.reset = reset
        if name is None:
            name = common.repr_operator("PD", ["pd"], [self.pd], [".1f"])
        super().__init__(name=name, **kwargs)

    @property
    def shape(self):
        return getattr(self.pd, "shape", (1,)) or (1,)


class System(Operator):
    """system arrays of the Imaging probe (epgpy/operator.py:348-361): outside the hot path"""

    def __init__(self, name=None, **properties):
        super().__init__(name=name)
        self.properties = properties


# --------------------------------------------------------------------------------------------- #
# differentiable operators (epgpy/diff.py:20-262, order 1)
# --------------------------------------------------------------------------------------------- #


def parse_order1(order1, parameters):
    """normalise the `order1` keyword to {variable: {parameter: coefficient}} (epgpy/diff.py:153-195)"""
    if isinstance(order1, str):
        order1 = [order1]
    if not order1:
        return {}
    if order1 is True:
        order1 = {p: {p: 1} for p in sorted(parameters)}
    elif isinstance(order1, (list, tuple, set)):
        order1 = {p: {p: 1} for p in order1}
    elif isinstance(order1, dict) and all(isinstance(v, str) for v in order1.values()):
        order1 = {var: {order1[var]: 1} for var in order1}
    elif isinstance(order1, dict) and all(isinstance(v, dict) for v in order1.values()):
        order1 = {var: dict(order1[var]) for var in order1}
    else:
        raise ValueError(f"Invalid parameter 'order1' value: {order1}")
    invalid = {p for var in order1 for p in set(order1[var]) - set(parameters)}
    if invalid:
        raise ValueError(f"Unknown parameter(s): {invalid}")
    return order1


def Pair(p1, p2=None):
    """sorted pair (epgpy/diff.py:534-540)"""
    if p2 is None:
        p1, p2 = p1
    return (p2, p1) if p1 > p2 else (p1, p2)


def parse_order2(order1, order2, parameters1, parameters2):
    """normalise the `order2` keyword to {(var1, var2): {parameter: coefficient}} (epgpy/diff.py:197-262).
    The coefficients are the SECOND derivatives of the operator's parameters with respect to the pair of variables
    (empty for parameters that are linear in the variables).  Which pairs are actually propagated is decided by the
    Hessian probe of the simulation, not by this keyword: every pair of variables it asks for gets its exact chain-rule
    update at every operator, the `auto_cross_derivatives` case of the reference (diff.py:333-362)."""
    if not order2:
        return {}
    if not order1:
        raise ValueError("order1 must be set.")
    if order2 is True:
        order2 = {Pair(pair): {} for pair in parameters2}
    elif isinstance(order2, str):
        order2 = {(order2, order2): {}}
    elif not isinstance(order2, dict) and all(isinstance(item, str) for item in order2):
        items = list(order2)
        order2 = {Pair(a, b): {} for a in items for b in items}
    elif not isinstance(order2, dict) and all(isinstance(pair, tuple) for pair in order2):
        order2 = {Pair(pair): {} for pair in order2}
    elif isinstance(order2, dict) and all(isinstance(pair, tuple) and isinstance(order2[pair], dict) for pair in order2):
        order2 = {Pair(pair): dict(order2[pair]) for pair in order2}
    else:
        raise ValueError(f"Invalid parameter 'order2' value: {order2}")
    invalid = {pair for pair in order2 if not (set(pair) & set(order1))}
    if invalid:
        raise ValueError(f"Invalid variable pair(s), no match in order1 variables: {invalid}")
    invalid = {pair for pair in order2 if (set(pair) - set(order1)) and order2[pair]}
    if invalid:
        raise ValueError(f"Invalid variable pair(s), expecting no coefficient: {invalid}")
    invalid = {param for pair in order2 for param in (set(order2[pair]) - set(parameters1))}
    if invalid:
        raise ValueError(f"Unknown parameter(s) in order2: {invalid}")
    return order2


class DiffOperator(Operator):
    """operator with order-1 and order-2 partial derivatives (epgpy/diff.py:20-378)

    order1: False | True | parameter name(s) | {alias: parameter} | {variable: {parameter: coeff}}
    order2: False | True | parameter name(s) | [(var1, var2), ...] | {(var1, var2): {parameter: coeff2}}
    The partial state matrices are propagated on the device in the same pass as the base state.
    """

    PARAMETERS_ORDER1 = set()
    PARAMETERS_ORDER2 = set()

    def __init__(self, *, order1=False, order2=False, name=None, duration=None):
        super().__init__(name=name, duration=duration)
        if (not order1) and isinstance(order2, (bool, str)):  # order2 alone names the variables (diff.py:158-159)
            order1 = order2
        self.order1 = parse_order1(order1, self.PARAMETERS_ORDER1)
        self.order2 = parse_order2(self.order1, order2, self.PARAMETERS_ORDER1, self.PARAMETERS_ORDER2)

    @property
    def parameters_order1(self):
        return {p for var in self.order1 for p in self.order1[var]}

    # lowering interface
    def _form(self):
        raise NotImplementedError

    def _lowered_form(self):
        """the coefficient form of the operator itself as the tape takes it (pulses reduced to their real-coupling
        kinds), computed once per operator object: T and E build it in their constructors, like the reference computes
        its coefficient arrays eagerly (epgpy/transition.py:32-36, evolution.py:107-118)"""
        f = self.__dict__.get("_lform")
        if f is None:
            f = self._form()
            if f[0] == "tgen":
                f = reduce_pulse(f)
            self._lform = f
        return f

    def _dform(self, param):
        raise NotImplementedError

    def _gform(self, param):
        """Op^-1 dOp/dparam as a coefficient form when it has a closed form (pre-injection), else None"""
        return None

    def _d2form(self, p1, p2):
        """d2 Op / dp1 dp2 as a coefficient form; None when it vanishes identically"""
        raise NotImplementedError(f"second derivatives of {type(self).__name__} w.r.t. ({p1}, {p2})")

    # generic dense coefficients (for `@` and for the public .mat / .arr attributes)
    def _dense(self):
        """('mat', mat[...,3,3], mat0|None) or ('diag', arr[...,3], arr0|None)"""
        raise NotImplementedError

    def _ddense(self, param):
        raise NotImplementedError


def _combine(op1, op2, name=None, duration=None):
    """`op1 @ op2`: one operator equal to op1 followed by op2 (epgpy/operator.py:206-241,
    opmatrix.py:89-135, opscalar.py:101-147).  Derivatives are carried per VARIABLE with the chain-rule
    coefficients folded in: d(M2 M1) = M2 dM1 + dM2 M1, affine terms included."""
    if not isinstance(op1, DiffOperator) or not isinstance(op2, DiffOperator) or isinstance(op1, S) or isinstance(op2, S):
        raise TypeError(f"Non-combinable operator: {op2 if isinstance(op1, DiffOperator) else op1}")
    k1, c1, c01 = op1._dense()
    k2, c2, c02 = op2._dense()
    diag = k1 == "diag" and k2 == "diag"

    def as_mat(kind, c):
        if c is None or kind == "mat" or diag:
            return c
        return c[..., None] * np.eye(3)

    tail = 1 if diag else 2

    def mul(b, a):  # b after a
        if a is None or b is None:
            return None
        a, b = expand_left(a, b, tail=tail)
        return b * a if diag else b @ a

    def add(a, b):
        if a is None:
            return b
        if b is None:
            return a
        a, b = expand_left(a, b, tail=tail)
        return a + b

    m1, m01, m2, m02 = as_mat(k1, c1), as_mat(k1, c01), as_mat(k2, c2), as_mat(k2, c02)
    mat = mul(m2, m1)
    mat0 = add(mul(m2, m01), m02)

    def scaled(c, coeff):
        if c is None:
            return None
        coeff = np.asarray(coeff)
        if coeff.ndim == 0:
            return c * coeff
        c, coeff = expand_left(c, coeff[(...,) + (None,) * tail], tail=tail)
        return c * coeff

    dvars = {}
    for op, first in ((op1, True), (op2, False)):
        for var, pc in op.order1.items():
            for param, coeff in pc.items():
                kd, d, d0 = op._ddense(param)
                d, d0 = as_mat(kd, d), as_mat(kd, d0)
                if first:  # M2 dM1, M2 dM01
                    d, d0 = mul(m2, d), mul(m2, d0)
                else:  # dM2 M1, dM2 M01 + dM02
                    d, d0 = mul(d, m1), add(mul(d, m01), d0)
                d, d0 = scaled(d, coeff), scaled(d0, coeff)
                if var in dvars:
                    dvars[var] = (add(dvars[var][0], d), add(dvars[var][1], d0))
                else:
                    dvars[var] = (d, d0)
    if name is None:
        name = f"{op1.name}|{op2.name}"
    if duration is None:
        duration = op1.duration + op2.duration
    cls = ScalarOp if diag else MatrixOp
    key = "darrs" if diag else "dmats"
    return cls(mat, mat0, **{key: dvars}, order1={v: {v: 1} for v in dvars}, name=name, duration=duration, check=False)


class CombinableOperator(DiffOperator):
    def __matmul__(self, other):
        return _combine(self, other)

    def __rmatmul__(self, other):
        return _combine(other, self)

    def combine(self, other, *, right=False, name=None, duration=None):
        return _combine(other, self, name, duration) if right else _combine(self, other, name, duration)


def reduce_pulse(form):
    """('tgen', blk6) -> ('tre' | 'tim', blk4) when B and U are real / B real and U imaginary (phi = +-90 / 0, 180 deg)"""
    blk = form[1]
    if not np.any(blk[..., 3]):
        if not np.any(blk[..., 5]):
            return ("tre", np.ascontiguousarray(blk[..., [0, 1, 2, 4]]))
        if not np.any(blk[..., 4]):
            b4 = blk[..., [0, 1, 2, 5]].copy()
            b4[..., 3] *= -1  # U = -i u
            return ("tim", b4)
    return form


def _cplx_block(*cols):
    """stack complex / real columns into a float64 block lead + (entry,)"""
    cols = np.broadcast_arrays(*cols)
    parts = []
    for c in cols:
        if np.iscomplexobj(c):
            parts += [c.real, c.imag]
        else:
            parts.append(np.asarray(c, dtype=float))
    return np.stack(parts, axis=-1).astype(np.float64)


class MatrixOp(CombinableOperator):
    """generic state-wise 3x3 operator with affine term (epgpy/opmatrix.py:10-135)"""

    def __init__(self, mat, mat0=None, *, dmats=None, d2mats=None, axes=None, check=True, **kwargs):
        dmats = dmats or {}
        d2mats = d2mats or {}
        self.PARAMETERS_ORDER1 = set(kwargs.pop("parameters_order1", None) or dmats)
        self.PARAMETERS_ORDER2 = {Pair(pair) for pair in (kwargs.pop("parameters_order2", None) or d2mats)}
        super().__init__(**kwargs)
        self.mat, self.mat0 = _matrix_setup(mat, mat0, check=check)
        self.d2mats = {Pair(p): _matrix_setup(*(d if isinstance(d, tuple) else (d, None)), check=check) for p, d in d2mats.items()}
        self.dmats = {p: _matrix_setup(*(d if isinstance(d, tuple) else (d, None)), check=check) for p, d in dmats.items()}
        if axes is not None:
            raise NotImplementedError("the `axes` keyword is not supported: give parameters their grid axes directly")

    @property
    def shape(self):
        return self.mat.shape[:-2]

    def _dense(self):
        return "mat", self.mat, self.mat0

    def _ddense(self, param):
        return ("mat",) + tuple(self.dmats[param])

    @staticmethod
    def _matrix_form(mat, mat0):
        blk = np.stack([mat.real, mat.imag], axis=-1).reshape(mat.shape[:-2] + (18,)).astype(np.float64)
        blk0 = None
        if mat0 is not None:
            col = mat0[..., :, 2]
            blk0 = np.stack([col.real, col.imag], axis=-1).reshape(col.shape[:-1] + (6,)).astype(np.float64)
        return ("matrix", blk, blk0)

    def _form(self):
        return self._matrix_form(self.mat, self.mat0)

    def _dform(self, param):
        return self._matrix_form(*self.dmats[param])

    def _d2form(self, p1, p2):
        d2 = self.d2mats.get(Pair(p1, p2))
        return None if d2 is None else self._matrix_form(*d2)


def _matrix_setup(mat, mat0=None, check=True):
    """epgpy/opmatrix.py:140-170"""
    mat = np.asarray(mat, dtype=complex)
    if mat.ndim == 2:
        mat = mat[None]
    if mat.ndim < 3 or mat.shape[-2:] != (3, 3):
        raise ValueError(f"Expected ...x3x3 array shape, found: {mat.shape}")
    if check and not np.allclose(mat, mat[..., (1, 0, 2), :][..., (1, 0, 2)].conj()):
        raise ValueError(f"Invalid matrix coefficients: {mat}")
    if mat0 is not None:
        mat0 = np.asarray(mat0, dtype=complex)
        if mat0.ndim == 2:
            mat0 = mat0[None]
        if mat0.ndim < 3 or mat0.shape[-2:] != (3, 3):
            raise ValueError(f"Expected ...x3x3 array shape, found: {mat0.shape}")
        mat, mat0 = expand_left(mat, mat0, tail=2)
        mat, mat0 = np.broadcast_arrays(mat, mat0)
    return mat, mat0


class ScalarOp(CombinableOperator):
    """generic diagonal operator with affine term (epgpy/opscalar.py:11-147)"""

    def __init__(self, arr, arr0=None, *, darrs=None, d2arrs=None, axes=None, check=True, **kwargs):
        darrs = darrs or {}
        d2arrs = d2arrs or {}
        self.PARAMETERS_ORDER1 = set(kwargs.pop("parameters_order1", None) or darrs)
        self.PARAMETERS_ORDER2 = {Pair(pair) for pair in (kwargs.pop("parameters_order2", None) or d2arrs)}
        super().__init__(**kwargs)
        self.arr, self.arr0 = _scalar_setup(arr, arr0, check=check)
        self.d2arrs = {Pair(p): _scalar_setup(*(d if isinstance(d, tuple) else (d, None)), check=check) for p, d in d2arrs.items()}
        self.darrs = {p: _scalar_setup(*(d if isinstance(d, tuple) else (d, None)), check=check) for p, d in darrs.items()}
        if axes is not None:
            raise NotImplementedError("the `axes` keyword is not supported: give parameters their grid axes directly")

    @property
    def shape(self):
        return self.arr.shape[:-1]

    def _dense(self):
        return "diag", self.arr, self.arr0

    def _ddense(self, param):
        return ("diag",) + tuple(self.darrs[param])

    @staticmethod
    def _diag_form(arr, arr0):
        a0 = np.zeros(arr.shape[:-1], dtype=complex) if arr0 is None else arr0[..., 2]
        return ("diag", _cplx_block(arr[..., 0].astype(complex), arr[..., 1].astype(complex),
                                    arr[..., 2].astype(complex), a0.astype(complex)), arr0 is not None)

    def _form(self):
        return self._diag_form(self.arr, self.arr0)

    def _dform(self, param):
        return self._diag_form(*self.darrs[param])

    def _d2form(self, p1, p2):
        d2 = self.d2arrs.get(Pair(p1, p2))
        return None if d2 is None else self._diag_form(*d2)


def _scalar_setup(arr, arr0=None, check=True):
    """epgpy/opscalar.py:161-192"""
    arr = np.asarray(arr, dtype=complex)
    if arr.ndim == 1:
        arr = arr[None]
    if arr.ndim < 2 or arr.shape[-1] != 3:
        raise ValueError(f"Expected ...x3 array shape, found: {arr.shape}")
    if check and not np.allclose(arr, arr[..., (1, 0, 2)].conj()):
        raise ValueError(f"Invalid coefficients: {arr}")
    if arr0 is not None:
        arr0 = np.asarray(arr0, dtype=complex)
        if arr0.ndim == 1:
            arr0 = arr0[None]
        arr, arr0 = expand_left(arr, arr0, tail=1)
        arr, arr0 = np.broadcast_arrays(arr, arr0)
    return arr, arr0


# --------------------------------------------------------------------------------------------- #
# T, Phi (epgpy/transition.py)
# --------------------------------------------------------------------------------------------- #


def _snap(x):
    """remove the 1e-17 residue of cos/sin at multiples of 90 degrees"""
    x = np.asarray(x, dtype=float).copy()
    x[np.abs(x) < 4e-16] = 0.0
    return x


def _cis_deg(phi, mult=1):
    p = DEG * np.asarray(phi, dtype=float) * mult
    return _snap(np.cos(p)) + 1j * _snap(np.sin(p))


class T(CombinableOperator):
    """instantaneous RF pulse: T = Rz(phi) Rx(alpha) Rz(-phi), angles in degrees
    (epgpy/transition.py:7-65, 114-151)"""

    PARAMETERS_ORDER1 = {"alpha", "phi"}

    def __init__(self, alpha, phi, *, axes=None, name=None, duration=None, **kwargs):
        self.alpha, self.phi = asparam(alpha), asparam(phi)
        if axes is not None:
            self.alpha, self.phi = common.set_axes([self.alpha, self.phi], axes)
        if not name:
            name = common.repr_operator("T", ["alpha", "phi"], [alpha, phi], [".1f", ".1f"])
        super().__init__(name=name, duration=duration, **kwargs)
        self._shape = op_shape(self.alpha, self.phi)
        self._lowered_form()

    @property
    def shape(self):
        return self._shape

    def _ap(self):
        a, p = expand_left(np.asarray(self.alpha, dtype=float), np.asarray(self.phi, dtype=float))
        return DEG * a, p

    def _form(self):
        # a = cos^2(alpha/2), w = cos(alpha), B = sin^2(alpha/2) e^{2 i phi}, U = -i sin(alpha) e^{i phi}
        a, phi = self._ap()
        B = np.sin(a / 2) ** 2 * _cis_deg(phi, 2)
        U = -1j * _snap(np.sin(a)) * _cis_deg(phi)  # sin(180 deg) = 1.2e-16 -> 0
        blk = np.empty(np.broadcast_shapes(a.shape, phi.shape) + (6,))
        blk[..., 0] = np.cos(a / 2) ** 2
        blk[..., 1] = _snap(np.cos(a))
        blk[..., 2], blk[..., 3], blk[..., 4], blk[..., 5] = B.real, B.imag, U.real, U.imag
        return ("tgen", blk)

    def _dform(self, param):
        a, phi = self._ap()
        if param == "alpha":  # transition.py:172-186 (per degree)
            B = 0.5 * np.sin(a) * _cis_deg(phi, 2) * DEG
            U = -1j * np.cos(a) * _cis_deg(phi) * DEG
            return ("tgen", _cplx_block(-0.5 * np.sin(a) * DEG + 0 * B.real, -np.sin(a) * DEG + 0 * B.real, B, U))
        if param == "phi":  # transition.py:165-169, 189-196: dB = 2i B, dU = i U (per degree)
            B = 2j * np.sin(a / 2) ** 2 * _cis_deg(phi, 2) * DEG
            U = np.sin(a) * _cis_deg(phi) * DEG
            return ("tgen", _cplx_block(0 * B.real, 0 * B.real, B, U))
        raise ValueError(param)

    PARAMETERS_ORDER2 = {("alpha", "alpha"), ("alpha", "phi"), ("phi", "phi")}

    def _d2form(self, p1, p2):
        """second derivatives per degree^2 (epgpy/transition.py:203-247): with a = cos^2(alpha/2), w = cos(alpha),
        B = sin^2(alpha/2) e^{2 i phi}, U = -i sin(alpha) e^{i phi} the pulse is linear in (a, w, B, U), hence so are
        its derivatives"""
        a, phi = self._ap()
        z1, z2 = _cis_deg(phi), _cis_deg(phi, 2)
        pair = Pair(p1, p2)
        if pair == ("alpha", "alpha"):
            B, U = 0.5 * np.cos(a) * z2, 1j * np.sin(a) * z1
            return ("tgen", _cplx_block(-0.5 * np.cos(a) * DEG**2 + 0 * B.real, -np.cos(a) * DEG**2 + 0 * B.real, B * DEG**2, U * DEG**2))
        if pair == ("alpha", "phi"):
            B, U = 1j * np.sin(a) * z2, np.cos(a) * z1
            return ("tgen", _cplx_block(0 * B.real, 0 * B.real, B * DEG**2, U * DEG**2))
        if pair == ("phi", "phi"):
            B, U = -4 * np.sin(a / 2) ** 2 * z2, 1j * np.sin(a) * z1
            return ("tgen", _cplx_block(0 * B.real, 0 * B.real, B * DEG**2, U * DEG**2))
        raise ValueError(pair)

    def _gform(self, param):
        """generator N = T^-1 dT/dparam, when it has a closed form: dT/dalpha = T . Rz(phi) Rx'(0) Rz(-phi), so a
        derivative injection can be done BEFORE the pulse (x_v += c N x_0) and the pulse then applied to all
        state sets at once: T(x_v + c N x_0) = T x_v + c dT x_0 (epgpy/diff.py:264-288 restated)."""
        if param != "alpha":
            return None
        _, phi = self._ap()
        U = -1j * _cis_deg(phi) * DEG
        zero = 0 * U.real
        return ("tgen", _cplx_block(zero, zero, zero + 0j, U))

    @staticmethod
    def _tgen_dense(blk):
        a, w = blk[..., 0], blk[..., 1]
        B, U = blk[..., 2] + 1j * blk[..., 3], blk[..., 4] + 1j * blk[..., 5]
        m = np.zeros(blk.shape[:-1] + (3, 3), dtype=complex)
        m[..., 0, 0], m[..., 0, 1], m[..., 0, 2] = a, B, U
        m[..., 1, 0], m[..., 1, 1], m[..., 1, 2] = B.conj(), a, U.conj()
        m[..., 2, 0], m[..., 2, 1], m[..., 2, 2] = -0.5 * U.conj(), -0.5 * U, w
        return m

    def _dense(self):
        return "mat", self._tgen_dense(self._form()[1]), None

    def _ddense(self, param):
        return "mat", self._tgen_dense(self._dform(param)[1]), None

    @property
    def mat(self):
        return self._dense()[1]

    mat0 = None


class Tx(T):
    def __init__(self, alpha, **kwargs):
        T.__init__(self, alpha, 0, **kwargs)


class Ty(T):
    def __init__(self, alpha, **kwargs):
        T.__init__(self, alpha, 90, **kwargs)


class Phi(CombinableOperator):
    """phase offset Rz(phi) (epgpy/transition.py:79-108, 140-151)"""

    PARAMETERS_ORDER1 = {"phi"}

    def __init__(self, phi, *, axes=None, name=None, duration=0, **kwargs):
        self.phi = asparam(phi)
        if axes is not None:
            (self.phi,) = common.set_axes([self.phi], axes)
        if not name:
            name = common.repr_operator("Phi", ["phi"], [phi], [".1f"])
        super().__init__(name=name, duration=duration, **kwargs)

    @property
    def shape(self):
        return op_shape(self.phi)

    def _arrs(self, deriv=False):
        z = np.atleast_1d(_cis_deg(self.phi))
        if deriv:
            return np.stack([1j * z * DEG, -1j * z.conj() * DEG, 0 * z], axis=-1)
        return np.stack([z, z.conj(), 1 + 0 * z], axis=-1)

    PARAMETERS_ORDER2 = {("phi", "phi")}

    def _form(self):
        return ScalarOp._diag_form(self._arrs(), None)

    def _dform(self, param):
        return ScalarOp._diag_form(self._arrs(True), None)

    def _d2form(self, p1, p2):
        z = np.atleast_1d(_cis_deg(self.phi))
        return ScalarOp._diag_form(np.stack([-z * DEG**2, -z.conj() * DEG**2, 0 * z], axis=-1), None)

    def _dense(self):
        return "diag", self._arrs(), None

    def _ddense(self, param):
        return "diag", self._arrs(True), None

    @property
    def mat(self):
        return self._arrs()[..., None] * np.eye(3)


# --------------------------------------------------------------------------------------------- #
# E, P, R (epgpy/evolution.py)
# --------------------------------------------------------------------------------------------- #


class _Evolution(CombinableOperator):
    def _dense(self):
        return ("diag",) + tuple(self._arrs())

    def _ddense(self, param):
        return ("diag",) + tuple(self._darrs(param))

    def _dform(self, param):
        return ScalarOp._diag_form(*self._darrs(param))

    @property
    def arr(self):
        return self._arrs()[0]

    @property
    def arr0(self):
        return self._arrs()[1]


def _evolution_arrays(rT, rL, r0=None):
    """arr = [conj e^{-rT}, e^{-rT}, e^{-rL}], arr0 = [0, 0, 1 - e^{-r0}] (epgpy/evolution.py:220-242)"""
    args = [np.asarray(rT), np.asarray(rL)] + ([] if r0 is None else [np.asarray(r0)])
    args = np.broadcast_arrays(*expand_left(*args))
    shape = args[0].shape
    arr = np.zeros(shape + (3,), dtype=complex)
    arr[..., 1] = np.exp(-args[0])
    arr[..., 0] = arr[..., 1].conj()
    arr[..., 2] = np.exp(-args[1])
    arr0 = None
    if r0 is not None:
        arr0 = np.zeros(shape + (3,), dtype=complex)
        arr0[..., 2] = 1 - np.exp(-args[2])
    return arr, arr0


class E(_Evolution):
    """relaxation, precession and recovery (epgpy/evolution.py:69-153, 251-256)
    tau, T1, T2 in ms, g in kHz"""

    PARAMETERS_ORDER1 = {"tau", "T1", "T2", "g"}

    def __init__(self, tau, T1, T2, g=0, *, axes=None, name=None, duration=None, **kwargs):
        self.tau, self.T1, self.T2, self.g = asparam(tau), asparam(T1), asparam(T2), asparam(g)
        if axes is not None:
            self.tau, self.T1, self.T2, self.g = common.set_axes([self.tau, self.T1, self.T2, self.g], axes)
        if not name:
            name = common.repr_operator("E", ["tau", "T1", "T2", "g"], [tau, T1, T2, g], [".1f", ".1f", ".1f", ".3f"])
        self._duration = duration
        duration = self.tau if duration is True else duration
        super().__init__(name=name, duration=duration, **kwargs)
        self._shape = op_shape(self.tau, self.T1, self.T2, self.g)
        self._lowered_form()

    @property
    def shape(self):
        return self._shape

    def _params(self):
        return expand_left(*[np.asarray(x, dtype=float) for x in (self.tau, self.T1, self.T2, self.g)])

    def _form(self):
        tau, T1, T2, g = self._params()
        e1 = np.exp(-tau / T1)
        blk0 = np.empty(e1.shape + (2,))
        blk0[..., 0], blk0[..., 1] = e1, 1 - e1
        blk1 = np.exp(-tau / T2)[..., None]
        blk2 = None
        if np.any(g != 0):
            ph = 2 * np.pi * g * tau
            blk2 = np.empty(ph.shape + (2,))
            blk2[..., 0], blk2[..., 1] = np.cos(ph), np.sin(ph)
        return ("e", blk0, blk1, blk2, True)

    def _arrs(self):
        tau, T1, T2, g = self._params()
        return _evolution_arrays(tau * (1 / T2 + 2j * np.pi * g), tau / T1, tau / T1)

    def _gform(self, param):
        """E^-1 dE/dparam (diagonal; evolution.py:360-399 divided by the operator itself): the recovery term
        makes Z'_v = e1 (Z_v + r (Z_0 - M0)), hence the affine part -r M0"""
        tau, T1, T2, g = self._params()
        # every block keeps the smallest broadcast shape of the parameters it really depends on
        if param == "tau":
            rM = -(1 / T2 + 2j * np.pi * g) + 0 * tau
            return _diag_gen(rM.conj(), rM, -1 / T1 + 0 * tau, True)
        if param == "T1":
            r = tau / T1**2
            return _diag_gen(0 * r, 0 * r, r, True)
        if param == "T2":
            r = tau / T2**2
            return _diag_gen(r, r, 0 * r, False)
        if param == "g":
            rM = -2j * np.pi * tau + 0 * g
            return _diag_gen(rM.conj(), rM, 0 * rM, False)
        return None

    PARAMETERS_ORDER2 = {("tau", "tau"), ("T1", "T1"), ("T2", "T2"), ("g", "g"), ("T1", "tau"), ("T2", "tau"), ("g", "tau"),
                         ("T2", "g")}

    def _d2form(self, p1, p2):
        """second derivatives (epgpy/evolution.py:405-488); pairs outside PARAMETERS_ORDER2 vanish (T1 acts on Z only,
        T2 and g on F+- only)"""
        tau, T1, T2, g = self._params()
        rT, rL = tau * (1 / T2 + 2j * np.pi * g), tau / T1
        eM, e1 = np.exp(-rT), np.exp(-rL)  # factors of F- and Z
        zero = 0 * (eM * e1)
        pair = Pair(p1, p2)
        fM = fZ = None
        if pair == ("tau", "tau"):
            fM, fZ = (rT / tau) ** 2 * eM, e1 / T1**2
        elif pair == ("T1", "T1"):
            fZ = e1 * (tau**2 / T1**4 - 2 * tau / T1**3)
        elif pair == ("T2", "T2"):
            fM = eM * (tau**2 / T2**4 - 2 * tau / T2**3)
        elif pair == ("g", "g"):
            fM = (-2j * np.pi * tau) ** 2 * eM
        elif pair == ("T1", "tau"):
            fZ = e1 * (1 - rL) / T1**2
        elif pair == ("T2", "tau"):
            fM = eM * (1 - rT) / T2**2
        elif pair == ("g", "tau"):
            fM = -2j * np.pi * (1 - rT) * eM
        elif pair == ("T2", "g"):
            fM = -2j * np.pi * (tau / T2) ** 2 * eM
        else:
            return None
        fM = zero if fM is None else fM + zero
        fZ = zero if fZ is None else fZ + zero
        # the recovery term 1 - e1 differentiates to minus the Z factor
        return ("diag", _cplx_block(np.conj(fM), fM, fZ + 0j, -fZ + 0j), bool(np.any(fZ != 0)))

    def _darrs(self, param):
        """first derivatives (epgpy/evolution.py:360-399)"""
        tau, T1, T2, g = self._params()
        rT, rL = tau * (1 / T2 + 2j * np.pi * g), tau / T1
        if param == "tau":
            arr, arr0 = _evolution_arrays(rT, rL, rL)
            arr[..., 1] *= -rT / tau
            arr[..., 0] = arr[..., 1].conj()
            arr[..., 2] *= -1 / T1
            arr0[..., 2] = -arr[..., 2]
            return arr, arr0
        if param == "T1":
            arr, arr0 = _evolution_arrays(0 * rT, rL, rL)
            arr[..., :2] = 0
            arr[..., 2] *= tau / T1**2
            arr0[..., 2] = -arr[..., 2]
            return arr, arr0
        if param == "T2":
            arr, _ = _evolution_arrays(rT, 0 * rL)
            arr[..., :2] *= (tau / T2**2)[..., None]
            arr[..., 2] = 0
            return arr, None
        if param == "g":
            arr, _ = _evolution_arrays(rT, 0 * rL)
            arr[..., 1] *= -2j * np.pi * tau
            arr[..., 0] = arr[..., 1].conj()
            arr[..., 2] = 0
            return arr, None
        raise ValueError(param)


def _diag_gen(rP, rM, rZ, affine):
    """pre-injection form of a diagonal operator: x_v += diag(rP, rM, rZ) x_0 (- rZ M0 at Z(0) if affine)"""
    rP, rM, rZ = np.broadcast_arrays(np.asarray(rP, dtype=complex), np.asarray(rM, dtype=complex), np.asarray(rZ, dtype=complex))
    a0 = -rZ if affine else 0 * rZ
    return ("diag", _cplx_block(rP, rM, rZ, a0), bool(affine))


class P(_Evolution):
    """precession only (epgpy/evolution.py:156-213, 245-248)"""

    PARAMETERS_ORDER1 = {"tau", "g"}

    def __init__(self, tau, g, *, axes=None, name=None, duration=None, **kwargs):
        self.tau, self.g = asparam(tau), asparam(g)
        if axes is not None:
            self.tau, self.g = common.set_axes([self.tau, self.g], axes)
        if not name:
            name = common.repr_operator("P", ["tau", "g"], [tau, g], [".1f", ".3f"])
        self._duration = duration
        duration = self.tau if duration is True else duration
        super().__init__(name=name, duration=duration, **kwargs)

    @property
    def shape(self):
        return op_shape(self.tau, self.g)

    def _params(self):
        return expand_left(np.asarray(self.tau, dtype=float), np.asarray(self.g, dtype=float))

    def _form(self):
        tau, g = self._params()
        ph = 2 * np.pi * g * tau
        one = np.ones((1,) * ph.ndim)
        return ("e", _cplx_block(one, 0 * one), _cplx_block(one), _cplx_block(np.cos(ph), np.sin(ph)), False)

    def _arrs(self):
        tau, g = self._params()
        return _evolution_arrays(2j * np.pi * g * tau, 0 * tau * g)

    def _gform(self, param):
        tau, g = self._params()
        rM = -2j * np.pi * (g if param == "tau" else tau) + 0 * (tau + g)
        return _diag_gen(rM.conj(), rM, 0 * rM, False)

    PARAMETERS_ORDER2 = {("tau", "tau"), ("g", "g"), ("g", "tau")}

    def _d2form(self, p1, p2):
        """epgpy/evolution.py:331-355"""
        tau, g = self._params()
        eM = np.exp(-2j * np.pi * g * tau)
        pair = Pair(p1, p2)
        if pair == ("tau", "tau"):
            fM = (-2j * np.pi * g) ** 2 * eM
        elif pair == ("g", "g"):
            fM = (-2j * np.pi * tau) ** 2 * eM
        else:
            fM = -2j * np.pi * (1 - 2j * np.pi * g * tau) * eM
        return ("diag", _cplx_block(np.conj(fM), fM, 0 * fM, 0 * fM), False)

    def _darrs(self, param):
        """epgpy/evolution.py:313-328"""
        tau, g = self._params()
        arr, _ = _evolution_arrays(2j * np.pi * g * tau, 0 * tau * g)
        arr[..., 1] *= -2j * np.pi * (g if param == "tau" else tau)
        arr[..., 0] = arr[..., 1].conj()
        arr[..., 2] = 0
        return arr, None


class R(_Evolution):
    """raw-rate evolution: F+- *= e^{-rT} (conj for F+), Z *= e^{-rL}, recovery 1 - e^{-r0}
    (epgpy/evolution.py:9-66, 220-242)"""

    PARAMETERS_ORDER1 = {"rT", "rL", "r0"}

    def __init__(self, rT=0, rL=0, *, r0=None, axes=None, name=None, duration=None, **kwargs):
        self.rT, self.rL, self.r0 = asparam(rT), asparam(rL), asparam(r0)
        if axes is not None:
            self.rT, self.rL, self.r0 = common.set_axes([self.rT, self.rL, self.r0], axes)
        if not name:
            name = common.repr_operator("R", ["rT", "rL", "r0"], [rT, rL, r0], [".1f", ".1f", ".1f"])
        super().__init__(name=name, duration=duration, **kwargs)

    @property
    def shape(self):
        return op_shape(self.rT, self.rL, self.r0)

    def _arrs(self):
        return _evolution_arrays(self.rT, self.rL, self.r0)

    def _form(self):
        return ScalarOp._diag_form(*self._arrs())

    PARAMETERS_ORDER2 = {("rT", "rT"), ("rL", "rL"), ("r0", "r0")}

    def _d2form(self, p1, p2):
        """epgpy/evolution.py:283-310: the second derivative of e^{-r} is e^{-r}; mixed pairs vanish"""
        if p1 != p2:
            return None
        d, d0 = self._darrs(p1)
        return ScalarOp._diag_form(-d, None if d0 is None else -d0)

    def _darrs(self, param):
        """epgpy/evolution.py:263-280"""
        rT, rL = np.asarray(self.rT), np.asarray(self.rL)
        if param == "rT":
            arr, _ = _evolution_arrays(rT, 0 * rL)
            arr[..., 2] = 0
            return -arr, None
        if param == "rL":
            arr, _ = _evolution_arrays(0 * rT, rL)
            arr[..., :2] = 0
            return -arr, None
        if param == "r0":
            arr, arr0 = _evolution_arrays(0 * rT, 0 * rL, 0 if self.r0 is None else self.r0)
            arr[:] = 0
            arr0[..., 2] -= 1
            return arr, -arr0
        raise ValueError(param)


# --------------------------------------------------------------------------------------------- #
# S (epgpy/shift.py), D (epgpy/diffusion.py)
# --------------------------------------------------------------------------------------------- #


class S(DiffOperator):
    """configuration shift (epgpy/shift.py:14-101).

    The device path implements the reference's `shift-1d` method: an integer k.  An integer VECTOR k
    (the reference's `shift-nd`) is accepted when every vector shift of the sequence is an integer
    multiple of one base vector: the occupied configurations are then collinear and the problem is
    exactly the 1-d one.  Float shifts (`shift-merge` / `shift-prune`) are outside the hot path."""

    def __init__(self, k, *, nmax=None, kgrid=None, prune=1e-8, name=None, duration=None):
        if np.allclose(k, 0):
            raise TypeError("Cannot have k == 0")
        if isinstance(k, (int, np.integer)) and not isinstance(k, bool):
            k = int(k)
        else:
            k = np.atleast_2d(k)
            if k.shape[-1] not in (1, 2, 3, 4):
                raise ValueError("k.shape[-1] must belong to [1, 2, 3, 4]")
        self.k, self.nmax, self.prune, self.kgrid = k, nmax, prune, kgrid
        if not name:
            name = common.repr_operator("S", ["k"], [k], ["" if isinstance(k, int) else ".2f"])
        super().__init__(name=name, duration=duration)

    @property
    def nshift(self):
        if isscalar(self.k):
            return abs(self.k)
        return int(np.round(np.max(np.abs(self.k))))

    @property
    def shape(self):
        return (1,) if isscalar(self.k) else self.k.shape[:-1]

    @property
    def kdim(self):
        return 1 if isscalar(self.k) else self.k.shape[-1]


class G(S):
    """gradient lobe (epgpy/shift.py:163-185): a shift by the wavenumber k = 2 pi gamma tau gradient (rad/m; tau in ms,
    gradient in mT/m, scalar or up to three components).  A float shift: it runs on the lattice when `kgrid` divides
    k * kvalue (lowering.py)."""

    def __init__(self, tau, gradient, *, duration=None, **kwargs):
        tau, gradient = np.asarray(tau, dtype=float), np.asarray(gradient, dtype=float)
        if np.any(tau < 0):
            raise ValueError("Cannot have negative time")
        if gradient.ndim and gradient.shape[-1] > 3:
            raise ValueError("Only 3d gradients are allowed")
        self.tau, self.gradient = tau, gradient
        super().__init__(get_wavenumber(tau, gradient), duration=tau if duration is True else duration, **kwargs)


class C(S):
    """time accumulation for temporal dephasing (epgpy/shift.py:188-210): a shift of tau * R2 along the FOURTH
    coordinate of the configurations (the first three are the wavenumber).  Configurations that differ only in accumulated
    time stay apart: F0 is the state refocused in space AND time."""

    def __init__(self, tau, R2=1, *, duration=None, **kwargs):
        tau, R2 = np.asarray(tau, dtype=float), np.asarray(R2, dtype=float)
        if np.any(tau < 0):
            raise ValueError("Cannot have negative time")
        evol = tau * R2
        self.tau, self.R2 = tau, R2
        super().__init__(np.stack([0 * evol] * 3 + [evol], axis=-1), duration=tau if duration is True else duration, **kwargs)


class D(Operator):
    """diffusion attenuation (epgpy/diffusion.py:14-79): tau in ms, D in mm^2/s (scalar or kdim x kdim),
    k (rad/m per unit shift) the shift of the S operator that immediately precedes it, if any"""

    def __init__(self, tau, D, k=None, *, method=None, name=None, duration=None):
        tau, D, k = asparam(tau), asparam(D), asparam(k)
        tau_shape, k_shape, D_shape = get_shape(tau), get_shape(k), get_shape(D)
        if len(k_shape) == 1:
            k_shape = (1,) + k_shape
        if len(D_shape) == 1:
            raise ValueError("D can only be a scalar or a 2d matrix")
        if len(set(D_shape[-2:])) == 2:
            raise ValueError("D must be a square 2d matrix")
        if len(D_shape) and len(k_shape) and D_shape[-1] != k_shape[-1]:
            raise ValueError("Incompatible D and k dimensions")
        self._shape = common.broadcast_shapes(tau_shape, D_shape[:-2], k_shape[:-1], (1,))
        self._kdim = k_shape[-1] if k_shape else 1
        if name is None:
            name = common.repr_operator("D", ["tau", "D", "k"], [tau, D, k], [".1f", "", ""])
        self._duration = duration
        if duration is True:
            duration = tau
        self.tau, self.D, self.k = tau, D, k
        super().__init__(name=name, duration=duration)

    @property
    def shape(self):
        return self._shape

    @property
    def kdim(self):
        return self._kdim


# --------------------------------------------------------------------------------------------- #
# probes (epgpy/probe.py, epgpy/diff.py:384-416)
# --------------------------------------------------------------------------------------------- #


class Probe(EmptyOperator):
    """base probe (epgpy/probe.py:7-79).  `obj` is a state-matrix attribute ('F0', 'Z0') or an eval expression over
    them, e.g. "abs(F0)" or "(real(F0), imag(F0))" (numpy names allowed, as in the reference): F0 and Z0 are read
    on the device and the expression is evaluated on the host copy, like the reference's `acquire`
    (probe.py:57-66).  Callables and expressions over other attributes (states, F, Z, k, ...) would need the
    whole state matrix on the host and are not supported (no CPU path)."""

    DEVICE_ATTRS = ("F0", "Z0")

    def __init__(self, obj, *args, post=None, **kwargs):
        self.attr = self.expr = None
        if isinstance(obj, str) and obj.strip() in self.DEVICE_ATTRS:
            self.attr = obj.strip()
        elif isinstance(obj, str):
            names = set(compile(obj, "<probe>", "eval").co_names)
            other = {n for n in names if n in Adc.SM_LOCALS and n not in self.DEVICE_ATTRS}
            if other:
                raise NotImplementedError(
                    f"Probe({obj!r}): state-matrix attributes {sorted(other)} are not available on the device path "
                    "(only F0 and Z0 are read back)")
            self.expr = obj
        else:
            raise NotImplementedError("callable probes need the whole state matrix on the host; use Adc / Jacobian "
                                      "or an expression over F0 / Z0")
        self._kwargs = kwargs
        self.phase = self.reduce = self.weights = None
        self._post = post
        super().__init__()
        self.name = f"Probe('{obj}')"

    def __call__(self, sm, **kwargs):
        return sm

    def post(self, obj):
        return obj if not getattr(self, "_post", None) else self._post(obj)

    def _eval(self, F0, Z0):
        env = {"F0": F0, "Z0": Z0}
        env.update(self._kwargs)
        return eval(self.expr, vars(np), env)


class Adc(Probe):
    """read-out with phase compensation / weights / reduction (epgpy/probe.py:82-165)"""

    SM_LOCALS = ["nstate", "ndim", "kdim", "states", "coords", "F", "F0", "F0t", "Z", "Z0", "k", "t", "t0"]

    def __init__(self, attr="F0", *, phase=None, reduce=None, weights=None, name="ADC"):
        if attr not in self.SM_LOCALS:
            raise ValueError(f"Invalid StateMatrix attribute: {attr}")
        self.attr = attr
        if phase is not None:
            phase = np.asarray(phase)
            self.phasor = np.exp(1j * phase / 180 * np.pi)
        self.phase = phase
        if reduce is not None and reduce is not True and reduce:
            reduce = (reduce,) if isinstance(reduce, int) else tuple(reduce)
            if not all(isinstance(ax, int) for ax in reduce):
                raise ValueError(f"Expected (tuple of) int, got: {reduce}")
        self.reduce = reduce
        if weights is not None:
            weights = np.asarray(weights)
            ndim = max(weights.ndim, 1)
            if reduce is None:
                self.reduce = tuple(range(ndim))
            elif reduce is not True and reduce:
                if not set(reduce) <= set(range(ndim)):
                    raise ValueError(f"Invalid reduce dimension(s): {reduce}")
        self.weights = weights
        Operator.__init__(self, name=name)

    def _post(self, obj):
        """phase compensation, applied on the host AFTER reduction like the reference (probe.py:155-165)"""
        arr = np.asarray(obj)
        if self.phase is not None:
            phasor = self.phasor
            if phasor.size > 1 and phasor.ndim < arr.ndim:
                phasor = np.expand_dims(phasor, tuple(range(phasor.ndim, arr.ndim)))
            arr = arr * phasor
        return arr


ADC = Adc()


class _FourierProbe(Probe):
    """probes that are linear functionals of ALL transverse configurations (epgpy/probe.py:168-219).  The engine reads
    the configurations (one row per lattice slot, EPGX_FLAG_SLOT) and `combine` applies the probe's weights on the host."""

    def __init__(self, coords=None, *, name=None, **opts):
        self.attr = self.expr = None
        self.coords = None if coords is None else np.asarray(coords)
        self.opts = opts
        self.phase = self.reduce = self.weights = None
        self._post = None
        self._kwargs = {}
        EmptyOperator.__init__(self, name=name or type(self).__name__)

    @staticmethod
    def _like_reference(F, k, t):
        """wavenumbers / accumulated times with the leading grid axes of the reference's sm.k / sm.t (statematrix.py:177-200)"""
        lead = (1,) * (np.ndim(F) - 1)
        return np.reshape(k, lead + np.shape(k)), np.reshape(t, lead + np.shape(t))

    def _coords(self, system):
        coords = self.coords if self.coords is not None else system.get("coords")
        if coords is None:
            raise ValueError(f"{type(self).__name__}: no coordinates (argument `coords` or System(coords=...))")
        return np.asarray(coords)


class DFT(_FourierProbe):
    """discrete Fourier transform of the F states at the positions `coords` (epgpy/probe.py:168-181): (*grid, *positions)"""

    def __init__(self, coords=None, *, name=None):
        super().__init__(coords, name=name)

    def combine(self, F, k, t, system, kdim):
        from .utils import dft

        k, t = self._like_reference(F, k, t)
        return dft(self._coords(system), F, k)


class Imaging(_FourierProbe):
    """imaging read-out (epgpy/probe.py:184-219): voxel shape, T2' / off-resonance modulation over the accumulated time,
    weights, reduction (`utils.imaging`); coordinates, modulation and weights default to the System arrays.
    (The reference POPS `modulation` / `weights` from the probe's options at its first acquisition, probe.py:205-210, so an
    Imaging object acquired twice loses them the second time; here the options are kept.)"""

    def combine(self, F, k, t, system, kdim):
        from .utils import imaging

        k, t = self._like_reference(F, k, t)
        opts = dict(self.opts)
        modulation = opts.pop("modulation", None)
        weights = opts.pop("weights", None)
        return imaging(self._coords(system), F, k, acctime=t if kdim == 4 else None,
                       modulation=system.get("modulation") if modulation is None else modulation,
                       weights=system.get("weights") if weights is None else weights, **opts)


class Jacobian(Probe):
    """probe returning the order-1 derivatives of the signal, stacked on a last axis
    (epgpy/diff.py:384-416); 'magnitude' inserts the signal itself"""

    def __init__(self, variables, *, probe="F0"):
        if probe not in ("F0", "Z0"):
            raise NotImplementedError("Jacobian probes 'F0' or 'Z0' only")
        self.probe = probe
        self.variables = variables if isinstance(variables, list) else [variables]
        self.phase = self.reduce = self.weights = None
        self.duration = 0
        self.name = f"Jacobian({probe})"

    def __repr__(self):
        return self.name


class Hessian(Probe):
    """probe returning the second derivatives of the signal (epgpy/diff.py:419-476): (..., nvar1, nvar2); a
    'magnitude' entry in either list selects the first derivatives with respect to the other variable"""

    def __init__(self, variables1, variables2=None, *, probe="F0"):
        if probe not in ("F0", "Z0"):
            raise NotImplementedError("Hessian probes 'F0' or 'Z0' only")
        self.probe = probe
        variables1 = variables1 if isinstance(variables1, list) else [variables1]
        if not variables2:
            variables2 = variables1
        elif not isinstance(variables2, list):
            variables2 = [variables2]
        self.variables1, self.variables2 = list(variables1), list(variables2)
        self.phase = self.reduce = self.weights = None
        self.duration = 0
        self.name = f"Hessian({probe})"

    def __repr__(self):
        return self.name

    def pairs(self):
        """unordered pairs of variables whose second derivatives this probe reads"""
        return {Pair(a, b) for a in self.variables1 for b in self.variables2 if "magnitude" not in (a, b)}

    def first(self):
        """variables whose first derivatives this probe reads (partners of 'magnitude')"""
        out = set()
        if "magnitude" in self.variables1:
            out |= set(self.variables2)
        if "magnitude" in self.variables2:
            out |= set(self.variables1)
        return out - {"magnitude"}


class PartialsPruner:
    """the reference's callback that drops partial states of negligible energy while a sequence runs
    (epgpy/diff.py:479-527): a speed-for-accuracy trade of its CPU loop, where every live partial state costs a full pass
    per operator.  On the device a partial state costs nothing before its injection (variables not yet injected are
    skipped, csrc/epgx_ring.cuh `alive`; csrc/epgx_pulsejac.cuh) and is never dropped afterwards, so
    `simulate(..., callback=PartialsPruner(...))` is accepted and returns the EXACT derivatives: they differ from the
    reference's pruned ones by at most what the pruning itself discards (`condition`, relative to the state norm)."""

    def __init__(self, *, condition=1e-5, variables=None):
        if not (callable(condition) or np.isscalar(condition)):
            raise TypeError(condition)
        self.condition = condition
        self.variables = set(variables) if variables else None

    def __repr__(self):
        return f"PartialsPruner({len(self.variables)} variables)" if self.variables else "PartialsPruner(all variables)"


from .exchange import X  # noqa: E402  (needs Operator)
