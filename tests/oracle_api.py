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

epg = types.SimpleNamespace(
    T=O.T, Phi=O.Phi, E=O.E, P=O.P, R=O.R, S=O.S, D=O.D, X=O.X,
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
    )
    kw.update(extra)
    res = O.simulate(case["seq"], **kw)
    if case.get("jac"):
        return np.asarray(res[0]), np.asarray(res[1])
    return np.asarray(res), None
