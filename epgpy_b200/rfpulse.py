"""Shaped RF pulses as operator chains (epgpy/rfpulse.py:37-197) -- host-side construction only.

A shaped pulse is a `MultiOperator` of N instantaneous rotations `T(alpha_i, phi_i)` whose angles follow the magnitude
and phase of the N complex samples, optionally interleaved with relaxation / precession over each sample's duration
(`functions.modify`), so that a frequency axis `g` gives the slice profile.  On the device the chain runs as one tape
with n = 0 orders (the register kernels; consecutive E . T . E runs are fused by the lowering).
`estimate_alpha` / `estimate_rf` (rfpulse.py:200-305) are small host computations on 3 x 3 matrices.
"""

import numpy as np

from . import functions
from . import operators as ops


def _net_rotation(alphas, phis):
    """product of the instantaneous rotations, first sample applied first (3 x 3 complex)"""
    total = np.eye(3, dtype=complex)
    for a, p in zip(np.atleast_1d(alphas), np.atleast_1d(phis)):
        total = ops.T(float(a), float(p)).mat.reshape(3, 3) @ total
    return total


def _angles(values, rf):
    """flip angle and phase (degrees) of every sample: alpha_i = 180 rf |v_i| (rfpulse.py:186-187)"""
    values = np.asarray(values)
    return 180.0 * np.abs(values) * rf, np.angle(values, deg=True)


def estimate_alpha(values, rf):
    """flip angle (degrees) reached by the whole pulse at RF amplitude `rf` (rfpulse.py:200-221): from the longitudinal
    magnetisation it leaves of an equilibrium state"""
    alphas, phis = _angles(np.asarray(values, dtype=complex), rf)
    z = float(np.real(_net_rotation(alphas, phis)[2, 2]))
    z = np.mod(z + 1, 2) - 1
    return float(np.mod(np.degrees(np.arccos(z)) + 180, 360) - 180)


def estimate_rf(values, alpha):
    """RF amplitude that makes the pulse reach flip angle `alpha` (rfpulse.py:224-305): closed form for pulses of
    constant phase, otherwise a one-parameter least-squares fit of the end state to the one of an ideal T(alpha, 90)"""
    values = np.asarray(values, dtype=complex)
    if np.max(np.abs(values)) > 1:
        raise ValueError("pulse values must have a magnitude <= 1")
    phases = np.mod(np.angle(values, deg=True), 180)
    guess = alpha / 180.0 / np.abs(np.sum(values))
    if np.all(np.isclose(np.diff(phases), 0, atol=1e-5)):
        return guess
    try:
        from scipy import optimize
    except ImportError as ex:  # pragma: no cover
        raise RuntimeError("Scipy is required for estimating rf") from ex
    eq = np.array([0, 0, 1], dtype=complex)
    target = np.abs(ops.T(alpha, 90).mat.reshape(3, 3) @ eq)
    unit, phis = _angles(values, 1.0)

    def cost(rf):
        return float(np.sum((np.abs(_net_rotation(np.asarray(rf).ravel()[0] * unit, phis) @ eq) - target) ** 2))

    return float(optimize.minimize(cost, guess, bounds=[(0, None)], tol=1e-8).x[0])


def rfpulse(values, duration, rf=None, alpha=None, phi=None, **kwargs):
    """(list of operators, info dict) of a shaped pulse (rfpulse.py:104-138)"""
    values = np.asarray(values, dtype=np.complex128)
    if values.ndim > 1:
        raise ValueError("`values` array must be 1-dimensional")
    if rf is None and alpha is None:
        raise ValueError('Either "rf" or "alpha" must be provided')
    if np.max(np.abs(values)) > 1:
        raise ValueError("pulse values must have a magnitude <= 1")
    if rf is None:
        rf = estimate_rf(values, alpha)
    elif alpha is None:
        alpha = estimate_alpha(values, rf)
    transform = kwargs.pop("transform", ops.T)
    n = len(values)
    if np.isscalar(duration):
        durations = np.full(n, duration / n)
    elif len(duration) == n:
        durations = np.asarray(duration, dtype=float)
    else:
        raise ValueError("duration and values must have the same length")
    shaped = values.reshape((n,) + (1,) * np.ndim(rf)) if np.ndim(rf) > 1 else values
    alphas, phis = 180.0 * np.abs(shaped) * rf, np.angle(shaped, deg=True)
    seq = [transform(a, p, duration=d) for a, p, d in zip(alphas, phis, durations)]
    if phi:  # phase offset of the whole pulse
        seq = [ops.Phi(-phi)] + seq + [ops.Phi(phi)]
    info = {"rf": rf, "alpha": alpha, "phi": phi}
    T1, T2, g = kwargs.get("T1"), kwargs.get("T2"), kwargs.get("g")
    if not (T1 is None and T2 is None and g is None):
        T1 = 1e10 if T1 is None else T1
        T2 = 1e10 if T2 is None else T2
        g = 0 if g is None else g
        seq = functions.modify(seq, T1=T1, T2=T2, g=g, expand=False)
        info.update(T1=T1, T2=T2, g=g)
    return seq, info


class RFPulse(ops.MultiOperator):
    """shaped RF pulse (epgpy/rfpulse.py:37-101): `values` complex samples of magnitude <= 1, `duration` in ms, `rf` the
    amplitude in kHz or `alpha` the target flip angle in degrees; T1 / T2 / g interleave relaxation and precession"""

    def __init__(self, values, duration, *, rf=None, alpha=None, phi=None, **kwargs):
        name = kwargs.pop("name", f"RFPulse({len(values)}, {duration}ms)")
        seq, info = rfpulse(values, duration, rf=rf, alpha=alpha, phi=phi, **kwargs)
        self.values = values
        for key, val in info.items():
            setattr(self, key, val)
        super().__init__(seq, name=name, duration=duration)


def encode_phase(pulse, gradient, fov, *, expand=True, rewind=None, npoint=101, gamma=None):
    """slice-selective version of a shaped pulse (epgpy/rfpulse.py:321-345): the positions `fov` (mm; a scalar field of
    view becomes `npoint` positions across it) see the off-resonance of `gradient` (mT/m) during every sample of the
    pulse -- a new grid axis behind the pulse's own unless expand=False -- and `rewind` (True: one half) appends the
    rephasing precession.  Returns the list of operators (an n = 0 tape on the engine: DESIGN.md section 2.2)."""
    from . import utils

    if not isinstance(pulse, RFPulse):
        raise TypeError("Can only use RFPulse operators")
    if np.isscalar(fov):
        fov = utils.spatial_range(fov, npoint)
    freqs = utils.space_to_freq(gradient, fov, gamma=utils.gamma_1H if gamma is None else gamma)
    if expand:
        freqs = np.reshape(freqs, (1,) * len(pulse.shape) + np.shape(freqs))
    out = functions.modify(pulse, g=freqs, expand=False)
    if rewind is not None:
        out.append(ops.P(pulse.duration * (0.5 if rewind is True else float(rewind)), g=-freqs, duration=0))
    return out
