"""Adapter: operator objects of the reference package -> the mirror classes of this package.

`from_reference(sequence)` reads only PUBLIC attributes of the reference's operators (SURVEY.md 8b:
`T.alpha/phi`, `E.tau/T1/T2/g`, `P.tau/g`, `R.rT/rL/r0`, `S.k/nmax`, `D.tau/D/k`, `X.tau/khi/axis/T1/T2/g`,
`Adc.attr/phase/reduce/weights`, `DFT.coords`, `Imaging.coords/opts`, `System.properties`, `PD.pd/reset`, `op.order1`, `op.duration`, generic `MatrixOp.mat/mat0/dmats`,
`ScalarOp.arr/arr0/darrs`) and rebuilds the sequence with `epgpy_b200.operators`, so a script written against
`epgpy` can hand its existing operator list to the B200 engine:

    import epgpy, epgpy_b200
    seq = [epgpy.epg.T(90, 90), ...]                      # built with the reference
    signal = epgpy_b200.epg.simulate(epgpy_b200.compat.from_reference(seq))

Identity is preserved (an operator object used several times maps to ONE mirror object), which keeps the
coefficient de-duplication of the lowering.  Classes are matched by NAME, so the reference need not be importable
here."""

import numpy as np

from . import operators as ops
from .exchange import X


def _order1(op):
    o1 = getattr(op, "order1", None) or {}
    return {var: dict(pc) for var, pc in o1.items()} or False


def _order2(op):
    """{(var1, var2): {param: coeff2}} of the reference operator; a pair without coefficient only SELECTS the pair in the
    reference (here the Hessian probe selects) but must stay listed so that `order2=` alone still names variables"""
    o2 = getattr(op, "order2", None) or {}
    if not isinstance(o2, dict):
        o2 = {pair: {} for pair in o2}
    return {tuple(pair): dict(c) for pair, c in o2.items()} or False


def _diff(op):
    return dict(order1=_order1(op), order2=_order2(op))


def _convert(op):
    name = type(op).__name__
    dur = getattr(op, "duration", 0)
    dur = getattr(op, "_duration", dur) if getattr(op, "_duration", None) is True else dur
    if name in ("T", "Tx", "Ty"):
        return ops.T(op.alpha, op.phi, **_diff(op), duration=dur, name=op.name)
    if name == "Phi":
        return ops.Phi(op.phi, **_diff(op), duration=dur, name=op.name)
    if name == "E":
        return ops.E(op.tau, op.T1, op.T2, op.g, **_diff(op), duration=dur, name=op.name)
    if name == "P":
        return ops.P(op.tau, op.g, **_diff(op), duration=dur, name=op.name)
    if name == "R":
        return ops.R(op.rT, op.rL, r0=op.r0, **_diff(op), duration=dur, name=op.name)
    if name in ("S", "G", "C"):  # (G and C are shifts by their wavenumber / accumulated time: shift.py:163-210)
        k = op.k if isinstance(op.k, (int, np.integer)) else np.asarray(op.k)
        return ops.S(k, nmax=op.nmax, kgrid=getattr(op, "kgrid", None), duration=dur, name=op.name)
    if name == "DFT":
        return ops.DFT(None if op.coords is None else np.asarray(op.coords), name=op.name)
    if name == "Imaging":
        return ops.Imaging(None if op.coords is None else np.asarray(op.coords), name=op.name, **dict(op.opts))
    if name == "System":
        return ops.System(name=op.name, **dict(op.properties))
    if name == "D":
        return ops.D(op.tau, op.D, op.k, duration=dur, name=op.name)
    if name == "X":
        return X(op.tau, op.khi, axis=op.axis, T1=op.T1, T2=op.T2, g=op.g, duration=dur, name=op.name)
    if name == "Adc":
        return ops.Adc(op.attr, phase=op.phase, reduce=op.reduce, weights=op.weights, name=op.name)
    if name == "Jacobian":
        return ops.Jacobian(list(op.variables), probe=op.probe)
    if name == "Hessian":
        return ops.Hessian(list(op.variables1), list(op.variables2), probe=op.probe)
    if name == "Probe":
        expr = getattr(op, "_expr", None)
        if expr is None:
            raise NotImplementedError("callable probes need the whole state matrix on the host")
        return ops.Probe(expr, post=getattr(op, "_post", None), **getattr(op, "_kwargs", {}))
    if name == "Spoiler":
        return ops.SPOILER
    if name == "Reset":
        return ops.RESET
    if name == "PD":
        return ops.PD(op.pd, reset=op.reset, name=op.name)
    if name in ("Wait", "Offset", "EmptyOperator"):
        new = ops.EmptyOperator(name=op.name)
        new.duration = dur
        return new
    if name == "MatrixOp":
        dm = {p: tuple(np.asarray(a) if a is not None else None for a in d) for p, d in getattr(op, "dmats", {}).items()}
        return ops.MatrixOp(np.asarray(op.mat), None if op.mat0 is None else np.asarray(op.mat0), dmats=dm,
                            order1=_order1(op), duration=dur, name=op.name, check=False)
    if name == "ScalarOp":
        da = {p: tuple(np.asarray(a) if a is not None else None for a in d) for p, d in getattr(op, "darrs", {}).items()}
        return ops.ScalarOp(np.asarray(op.arr), None if op.arr0 is None else np.asarray(op.arr0), darrs=da,
                            order1=_order1(op), duration=dur, name=op.name, check=False)
    raise NotImplementedError(f"operator {name} of the reference has no device implementation (outside the hot path)")


def from_reference(sequence, _memo=None):
    """nested list / MultiOperator of reference operators -> nested list of mirror operators"""
    memo = {} if _memo is None else _memo
    if isinstance(sequence, (list, tuple)):
        return [from_reference(item, memo) for item in sequence]
    if isinstance(sequence, ops.Operator):
        return sequence
    if type(sequence).__name__ == "MultiOperator":
        return [from_reference(item, memo) for item in sequence.operators]
    if id(sequence) not in memo:
        memo[id(sequence)] = _convert(sequence)
    return memo[id(sequence)]
