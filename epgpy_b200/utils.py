"""Small host helpers of the reference's `utils` that sequences on the hot path use (epgpy/utils.py:134-169), and the
weights of its Fourier probes (`imaging`, `dft`: epgpy/utils.py:12-115) -- a probe is a linear functional of the
configurations: the device reads the configurations, the host applies the weights (lowering.py, Row "fourier")."""

import enum

import numpy as np

gamma_1H = 42576.0  # kHz/T (epgpy/utils.py:8)


def Axes(*names):
    """Enum of grid axes: `ax = Axes("T2", "B1"); E(..., axes=ax.T2)` (epgpy/utils.py:134-145)"""
    return enum.IntEnum("Axes", names, start=0)


def get_wavenumber(grad, duration, gamma=gamma_1H):
    """wavenumber (rad/m) of a gradient lobe: grad in mT/m, duration in ms (epgpy/utils.py:157-169)"""
    return 2 * np.pi * gamma * np.asarray(grad) * 1e-3 * np.asarray(duration)


def spatial_range(fov, nvalue=100):
    """`nvalue` positions across a field of view `fov` (mm), centred (epgpy/utils.py:175-183)"""
    return fov * np.linspace(-0.5, 0.5, nvalue)


def space_to_freq(grad, positions, *, gamma=gamma_1H):
    """off-resonance (kHz) of positions (mm) under a gradient (mT/m) (epgpy/utils.py:186-208)"""
    return grad * 1e-6 * gamma * (positions if np.isscalar(positions) else np.asarray(positions))


def freq_to_space(grad, frequencies, *, gamma=gamma_1H):
    """positions (mm) of off-resonances (kHz) under a gradient (mT/m) (epgpy/utils.py:211-213)"""
    return frequencies / grad / gamma * 1e6


def check_states(states):
    """F-(k) = conj F+(-k), Z(k) = conj Z(-k) (epgpy/utils.py:118-121)"""
    states = np.asarray(states)
    return np.allclose(states, states[..., ::-1, [1, 0, 2]].conj())


def get_norm(states):
    """norm of the transverse + longitudinal states without the F+ column (epgpy/utils.py:152-154)"""
    return np.sqrt(np.sum(np.abs(np.asarray(states)[..., 1:]) ** 2, axis=(-2, -1)))


def cexp(arr):
    """exp(1j * arr) (epgpy/utils.py:124-131)"""
    arr = np.asarray(arr, dtype=float)
    return np.cos(arr) + 1j * np.sin(arr)


def _negligible(w, tol):
    """configurations whose weight stays below `tol` for every atom and position (the reference drops them before the
    transform, utils.py:50-53, 63-66: part of its result at the 1e-8 level, so it is reproduced)"""
    return ~np.any(np.abs(w) > tol, axis=tuple(range(w.ndim - 1)))


def imaging(positions, states, wavenumbers, acctime=None, *, phase=None, weights=None, modulation=None, voxel_shape="box",
            voxel_size=1, expand=True, reduce=True, tol=1e-8):
    """image-space signal of the transverse configurations (epgpy/utils.py:12-95): sum over the configurations of
        F * voxel(k) * modulation(t) * exp(i k.x)
    with voxel(k) = prod sinc(k size / 2 pi) for box voxels (1 for points), modulation(t) = exp(-|t| Re m + 2 pi i t Im m)
    over the accumulated time t (when given), an optional phase (degrees) and weights, then the reduction.
    Shapes: states [..., nstate], wavenumbers [..., nstate, ndim] (rad/m), positions [..., ndim]; with `expand` the
    position axes are inserted in front of the configuration axis."""
    x = np.asarray(positions, dtype=float)
    if x.ndim == 1:
        x = x[:, None]
    F, k = np.asarray(states), np.asarray(wavenumbers, dtype=float)
    t = None if acctime is None else np.asarray(acctime, dtype=float)
    if expand:
        npos_axes = x.ndim - 1
        F = F.reshape(F.shape[:-1] + (1,) * npos_axes + F.shape[-1:])
        k = k.reshape(k.shape[:-2] + (1,) * npos_axes + k.shape[-2:])
        if t is not None:
            t = t.reshape(t.shape[:-1] + (1,) * npos_axes + t.shape[-1:])
    if voxel_shape not in ("point", "box"):
        raise ValueError(f"Unknown voxel shape: {voxel_shape}")
    amp = F
    if voxel_shape == "box":
        vox = np.prod(np.sinc(k * voxel_size / (2 * np.pi)), axis=-1)
        drop = _negligible(vox, tol)
        keep = ~drop if drop.ndim else slice(None)
        amp, k, vox = amp[..., keep], k[..., keep, :], vox[..., keep]
        t = None if t is None else t[..., keep]
        amp = amp * vox
    if t is not None:
        m = np.asarray(1.0 if modulation is None else modulation)
        decay = np.exp(-np.abs(t) * m.real[..., None])
        keep = ~_negligible(decay, tol)
        amp, k, t, decay = amp[..., keep], k[..., keep, :], t[..., keep], decay[..., keep]
        amp = amp * decay
        if np.iscomplexobj(m):
            amp = amp * cexp(2 * np.pi * t * m.imag[..., None])
    if phase is not None:
        amp = amp * np.exp(1j * np.pi / 180 * phase)
    ndim = x.shape[-1]
    im = np.sum(amp * cexp(np.sum(k[..., :ndim] * x[..., None, :], axis=-1)), axis=-1)
    if weights is not None:
        im = im * np.asarray(weights)
    if reduce is True:
        return im.sum()
    return im if reduce is False else im.sum(axis=reduce)


def dft(coords, states, wavenumbers, *, reduce=False):
    """simplified imaging function: point voxels, no modulation (epgpy/utils.py:113-115)"""
    return imaging(coords, states, wavenumbers, reduce=reduce, voxel_shape="point")
