"""Multi-compartment exchange operator X (epgpy/exchange.py).

Host side: the tiny N x N transition matrices exp((-khi + diag(-1/T2 + 2 pi i g)) tau) and
exp((-khi - diag(1/T1)) tau) are precomputed once per operator (an eigen-decomposition of at most a
few thousand 2x2 / 3x3 matrices); the per-order application s <- m (s - eq) + eq along the pool axis
(exchange.py:89-120) runs on the device (EPGX_OP_X in csrc/).
"""

import numpy as np

from . import common
from .common import asparam, isscalar


def exchange_matrix(k, *, axis=-1, ncomp=2, densities=None):
    """scalar exchange rate(s) -> kinetic matrix with two new axes of size ncomp
    (epgpy/exchange.py:127-151): khi[..., i(axis), ..., j] = k * (delta_ij - (1-delta_ij)/(ncomp-1)) / densities[j]"""
    k = np.asarray(k, dtype=float)
    if np.any(k < 0):
        raise ValueError("Cannot have negative echange rate")
    if axis > k.ndim:
        k = np.expand_dims(k, tuple(range(k.ndim, axis)))
    axis = (k.ndim + axis + 1) if axis < 0 else axis
    eye = np.eye(ncomp)
    kron = eye + (eye - 1) / (ncomp - 1)
    if densities is not None:
        kron = kron / np.asarray(densities, dtype=float)
    return np.moveaxis(k[..., None, None] * kron, -2, axis)


def expm_batched(mat):
    """exp of a stack of small square matrices through an eigen-decomposition of the globally
    normalised stack (the reference's recipe, epgpy/exchange.py:262-282, so that both agree to
    round-off)"""
    nrm = np.linalg.norm(mat)
    n = mat.shape[-1]
    if np.isclose(nrm, 0):
        return np.broadcast_to(np.eye(n), mat.shape).astype(mat.dtype)
    scaled = mat / nrm
    tr = lambda m: np.swapaxes(m, -1, -2)  # noqa: E731
    if np.allclose(mat, tr(mat).conj()):
        w, v = np.linalg.eigh(scaled)
    else:
        w, v = np.linalg.eig(scaled)
    ew = np.expm1(w * nrm) + 1
    return tr(np.linalg.solve(tr(v), ew[..., None] * tr(v)))


def exchange_operator(tau, khi, *, axis=0, T1=None, T2=None, g=None):
    """mat[..., dst(axis), src(axis+1), ..., 3] = (mT, conj mT, mL)  (epgpy/exchange.py:154-203)"""
    khi = np.asarray(khi, dtype=float)
    tau = np.asarray(tau, dtype=float)
    T1 = np.asarray(np.inf if T1 is None else T1, dtype=float)
    T2 = np.asarray(np.inf if T2 is None else T2, dtype=float)
    g = np.asarray(0.0 if g is None else g, dtype=float)
    n = khi.shape[-1]
    eye = np.eye(n)
    minshape = khi.shape[:-1]
    # left-aligned common shape of all parameters
    shape = common.broadcast_shapes(tau.shape, T1.shape, T2.shape, g.shape, minshape)
    nd = len(shape)
    tau, T1, T2, g = [common.left(a, nd) for a in (tau, T1, T2, g)]
    T1, T2, g = [np.broadcast_to(a, shape) for a in (T1, T2, g)]
    khi = np.expand_dims(khi, tuple(range(nd - len(minshape))))
    tau, T1, T2, g = [np.moveaxis(a, axis, -1) for a in (tau, T1, T2, g)]
    xT = -khi + (-1 / T2 + 2j * np.pi * g)[..., None] * eye
    xL = -khi + (-1 / T1)[..., None] * eye
    mT = expm_batched(xT * tau[..., None])
    mL = expm_batched(xL * tau[..., None])
    mT = np.moveaxis(mT, (-2, -1), (axis, axis + 1))
    mL = np.moveaxis(mL, (-2, -1), (axis, axis + 1))
    return np.stack([mT, mT.conj(), mL.astype(complex)], axis=-1)


from .operators import Operator  # noqa: E402


class X(Operator):
    """exchange + relaxation/precession of N coupled compartments along grid axis `axis`
    (epgpy/exchange.py:11-120)"""

    def __init__(self, tau, khi, *, axis=-1, T1=None, T2=None, g=None, name=None, duration=None):
        if isscalar(khi):
            khi = exchange_matrix(khi, axis=axis, ncomp=2)
        else:
            khi = np.asarray(khi, dtype=float)
            if khi.ndim < 2:
                raise ValueError("Exchange matrix matrix must be at least 2D")
            if khi.shape[:-1][axis] != khi.shape[-1]:
                raise ValueError("Exchange matrix must be square")
            if not all(np.allclose(khi[..., i].sum(axis=axis), 0) for i in range(khi.shape[-1])):
                raise ValueError(f"Exchange matrix must sum to 0 along axis {axis}")
        axis = int(khi.ndim + axis - 1) if axis < 0 else int(axis)
        self.mat = exchange_operator(tau, khi, axis=axis, T1=T1, T2=T2, g=g)
        self.axis = axis
        self.khi = khi
        self.T1, self.T2, self.g, self.tau = asparam(T1), asparam(T2), asparam(g), asparam(tau)
        self._duration = duration
        if duration is True:
            duration = self.tau
        if name is None:
            name = common.repr_operator("X", ["tau", "khi"], [tau, khi])
        super().__init__(name=name, duration=duration)

    @property
    def shape(self):
        return tuple(d for i, d in enumerate(self.mat.shape[:-1]) if i != self.axis + 1)

    @property
    def ncomp(self):
        return self.mat.shape[self.axis]
