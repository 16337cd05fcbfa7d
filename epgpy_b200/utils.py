"""Small host helpers of the reference's `utils` that sequences on the hot path use
(epgpy/utils.py:134-169); imaging / DFT helpers are outside the hot path."""

import enum

import numpy as np

gamma_1H = 42576.0  # kHz/T (epgpy/utils.py:8)


def Axes(*names):
    """Enum of grid axes: `ax = Axes("T2", "B1"); E(..., axes=ax.T2)` (epgpy/utils.py:134-145)"""
    return enum.IntEnum("Axes", names, start=0)


def get_wavenumber(grad, duration, gamma=gamma_1H):
    """wavenumber (rad/m) of a gradient lobe: grad in mT/m, duration in ms (epgpy/utils.py:157-169)"""
    return 2 * np.pi * gamma * np.asarray(grad) * 1e-3 * np.asarray(duration)
