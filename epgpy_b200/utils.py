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


def cexp(arr):
    """exp(1j * arr) (epgpy/utils.py:124-131)"""
    arr = np.asarray(arr, dtype=float)
    return np.cos(arr) + 1j * np.sin(arr)


def imaging(positions, states, wavenumbers, acctime=None, *, phase=None, weights=None, modulation=None, voxel_shape="box",
            voxel_size=1, expand=True, reduce=True, tol=1e-8):
    """inverse discrete Fourier transform of the transverse configurations (epgpy/utils.py:12-95):
    states [..., nstate], wavenumbers [..., nstate, ndim] (rad/m), positions [..., ndim]; voxel shape (sinc of a box),
    modulation exp(-|t| Re m + 2 pi i t Im m) over the accumulated time `acctime`, phase (degrees), weights, reduction"""
    F = np.asarray(states)
    k = np.asarray(wavenumbers, dtype=float)
    t = np.asarray(acctime, dtype=float) if acctime is not None else None
    pos = np.asarray(positions, dtype=float)
    pos = pos if pos.ndim > 1 else pos[..., None]
    if expand:  # insert the position axes into F and k
        dims = np.arange(pos.ndim - 1)
        F = np.expand_dims(F, tuple(-2 - dims))
        k = np.expand_dims(k, tuple(-3 - dims))
        if t is not None:
            t = np.expand_dims(t, tuple(-2 - dims))
    if voxel_shape == "point":
        voxel = 1.0
    elif voxel_shape == "box":
        voxel = np.sinc(k * voxel_size / 2 / np.pi).prod(-1)
        kmask = np.any(np.abs(voxel) > tol, axis=tuple(range(F.ndim - 1)))
        F, k, voxel = F[..., kmask], k[..., kmask, :], voxel[..., kmask]
        if t is not None:
            t = t[..., kmask]
    else:
        raise ValueError(f"Unknown voxel shape: {voxel_shape}")
    if t is not None:
        modulation = np.asarray(modulation if modulation is not None else 1.0)
        mod = np.exp(-np.abs(t) * modulation.real[..., None])
        mmask = np.any(mod > tol, axis=tuple(range(F.ndim - 1)))
        F, k, mod = F[..., mmask], k[..., mmask, :], mod[..., mmask]
        if getattr(voxel, "shape", None):
            voxel = voxel[..., mmask]
        if np.iscomplexobj(modulation):
            mod = mod * cexp(t[..., mmask] * 2 * np.pi * modulation.imag[..., None])
    else:
        mod = 1.0
    if phase is not None:
        mod = mod * np.exp(1j * phase * np.pi / 180)
    kdim = pos.shape[-1]
    f = voxel * mod * F
    kp = np.matmul(k[..., :kdim], pos[..., None])[..., 0]
    im = np.matmul(f[..., None, :], cexp(kp)[..., None])[..., 0, 0]
    if weights is not None:
        im = im * np.asarray(weights)
    if reduce is True:
        return im.sum()
    if reduce is not False:
        return im.sum(axis=reduce)
    return im


def dft(coords, states, wavenumbers, *, reduce=False):
    """simplified imaging function: point voxels, no modulation (epgpy/utils.py:113-115)"""
    return imaging(coords, states, wavenumbers, reduce=reduce, voxel_shape="point")
