"""Adapter: exposes the CPU oracle (oracle/epg_oracle.py) under the reference's operator
names so that tests/cases.py can build every workload for it.  Test infrastructure only."""

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import epg_oracle as O  # noqa: E402

def _gradient(tau, gradient, **kw):
    """gradient lobe = shift by k = 2 pi gamma tau gradient (epgpy/shift.py:163-185, utils.py:157-169)"""
    return O.S(2 * np.pi * 42576.0 * np.asarray(gradient, dtype=float) * 1e-3 * float(tau), **kw)


def _time_accumulation(tau, R2=1, **kw):
    """shift of tau * R2 along the fourth (time) coordinate (epgpy/shift.py:188-210)"""
    return O.S(np.array([0.0, 0.0, 0.0, float(tau) * float(R2)]), **kw)


epg = types.SimpleNamespace(
    T=O.T, Phi=O.Phi, E=O.E, P=O.P, R=O.R, S=O.S, D=O.D, X=O.X, G=_gradient, C=_time_accumulation,
    Adc=O.ADC, ADC=O.ADC(), SPOILER=O.SPOILER(), RESET=O.RESET(), PD=O.PD, Wait=O.WAIT,
    exchange_matrix=O.kinetic_matrix,
)


def run(case, **extra):
    """run a tests/cases.py case with the oracle -> (signal, jacobian|None)"""
    opts = dict(case.get("options") or {})
    kw = dict(
        max_nstate=opts.get("max_nstate"),
        kvalue=opts.get("kvalue", 1.0),
        kvec=case.get("kvec"),
        density=case.get("density") if case.get("density") is not None else 1.0,
        jacobian=case.get("jac"),
        init=case.get("init"),
        kgrid=opts.get("kgrid"),
        tvalue=opts.get("tvalue", 1.0),
    )
    kw.update(extra)
    res = O.simulate(case["seq"], **kw)
    if case.get("jac"):
        return np.asarray(res[0]), np.asarray(res[1])
    return np.asarray(res), None


# ---- shaped RF pulse for the oracle (epgpy/rfpulse.py:104-197, 224-305): a plain list of oracle operators


def _net_rotation(alphas, phis):
    total = np.eye(3, dtype=complex)
    for a, p in zip(alphas, phis):
        total = O.rf_matrix(a, p)[0] @ total
    return total


def estimate_rf(values, alpha):
    values = np.asarray(values, dtype=complex)
    guess = alpha / 180.0 / np.abs(np.sum(values))
    if np.all(np.isclose(np.diff(np.mod(np.angle(values, deg=True), 180)), 0, atol=1e-5)):
        return guess
    from scipy import optimize

    eq = np.array([0, 0, 1], dtype=complex)
    target = np.abs(O.rf_matrix(alpha, 90)[0] @ eq)
    unit, phis = 180.0 * np.abs(values), np.angle(values, deg=True)
    cost = lambda rf: float(np.sum((np.abs(_net_rotation(np.ravel(rf)[0] * unit, phis) @ eq) - target) ** 2))
    return float(optimize.minimize(cost, guess, bounds=[(0, None)], tol=1e-8).x[0])


def RFPulse(values, duration, *, rf=None, alpha=None, phi=None, T1=None, T2=None, g=None):
    values = np.asarray(values, dtype=complex)
    if rf is None:
        rf = estimate_rf(values, alpha)
    n = len(values)
    seq = []
    for v in values:
        seq.append(O.T(180.0 * np.abs(v) * rf, np.angle(v, deg=True), duration=duration / n))
        if not (T1 is None and T2 is None and g is None):
            seq.append(O.E(duration / n, 1e10 if T1 is None else T1, 1e10 if T2 is None else T2, 0 if g is None else g))
    if phi:
        seq = [O.Phi(-phi)] + seq + [O.Phi(phi)]
    return seq


epg.RFPulse = RFPulse
