"""`epg` namespace: the reference's public surface (epgpy/core.py:80-83, epgpy/operators.py:1-25)
for the EPG operator-chain hot path.

    from epgpy_b200 import epg
    signal = epg.simulate([epg.T(90, 90), epg.S(1), epg.E(5, 1000, 30), epg.ADC])
"""

from .statematrix import StateMatrix
from .utils import Axes, get_wavenumber
from .operators import (
    Operator, MultiOperator, EmptyOperator, Spoiler, Wait, Offset, Reset, PD, System,
    DiffOperator, MatrixOp, ScalarOp,
    Probe, Adc, DFT, Imaging, Jacobian, Hessian, PartialsPruner,
    E, P, R, T, Tx, Ty, Phi, S, G, C, D,
    ADC, NULL, SPOILER, RESET,
)
from .exchange import X, exchange_matrix
from .functions import simulate, modify, get_adc_times, getshape, getnshift, getkdim

__all__ = [name for name in dir() if not name.startswith("_")]
