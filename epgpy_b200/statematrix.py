"""StateMatrix: host-side description of an initial / final EPG state (epgpy/statematrix.py:9-374).

In the reference this class owns the live `[..., 2n+1, 3]` complex128 array every operator mutates.
Here the live state never exists on the host: during `simulate` it lives in the registers / shared
memory of the sm_100a kernel in half storage (orders k >= 0).  The class keeps the reference's
constructor and read-only attributes so that `simulate(init=StateMatrix(...))`, `density=` and the
state-matrix options (`max_nstate`, `kvalue`) work unchanged.
"""

import numpy as np

from . import common


def _format_states(states, check=True):
    """(epgpy/statematrix.py:388-422) -> complex128 [..., 2n+1, 3] with symmetry checks"""
    states = np.asarray(states).astype(np.complex128)
    if states.ndim == 1:
        if check and states.size != 3:
            raise ValueError("The number of state dimensions must be 3")
        states = states.reshape((1, 1, 3))
    elif states.ndim == 2:
        if check and states.shape[1] != 3:
            raise ValueError("The number of state dimensions must be 3")
        if check and states.shape[0] % 2 != 1:
            raise ValueError("The number of states must be odd")
        states = states.reshape((1,) + states.shape)
    else:
        if check and states.shape[-1] != 3:
            raise ValueError("The number of state dimensions must be 3")
        if check and states.shape[-2] % 2 != 1:
            raise ValueError("The number of states must be odd")
    if check:
        if not np.allclose(states[..., 1], states[..., ::-1, 0].conj()):
            raise ValueError("The F-state columns do no match.")
        if not np.allclose(states[..., 2], states[..., ::-1, 2].conj()):
            raise ValueError("The Z-state columns is not symmetrical.")
    return states


def _resize(states, n):
    """symmetric zero-pad / crop to n orders (epgpy/statematrix.py:793-804)"""
    cur = (states.shape[-2] - 1) // 2
    if n > cur:
        pad = [(0, 0)] * (states.ndim - 2) + [(n - cur, n - cur), (0, 0)]
        return np.pad(states, pad)
    if n < cur:
        d = cur - n
        return states[..., d:-d, :]
    return states


class StateMatrix:
    """phase states of an n-dimensional system (constructor: epgpy/statematrix.py:12-80)"""

    def __init__(self, init=None, *, density=1, equilibrium=None, coords=None, kvalue=1.0, tvalue=1.0,
                 nstate=None, shape=None, check=True, **options):
        if coords is not None:
            raise NotImplementedError("state matrices with explicit coordinates (shift-nd / shift-merge) are outside the hot path")
        if equilibrium is None:
            density = np.atleast_1d(np.asarray(density, dtype=float))
            equilibrium = density[..., None, None] * np.array([[[0, 0, 1]]])
        equilibrium = _format_states(equilibrium, check=check)
        n_eq = (equilibrium.shape[-2] - 1) // 2
        if np.any(equilibrium[..., :2] != 0) or np.any(np.delete(equilibrium[..., 2], n_eq, axis=-1) != 0) \
                or np.any(equilibrium[..., n_eq, 2].imag != 0):
            raise NotImplementedError("only equilibria of the form [0, 0, M0] at order 0 are supported")
        self._density = equilibrium[..., n_eq, 2].real
        states = equilibrium if init is None else _format_states(init, check=check)
        n = (states.shape[-2] - 1) // 2
        if nstate and nstate > n:
            states = _resize(states, nstate)
        lead = common.broadcast_shapes(states.shape[:-2], self._density.shape, tuple(shape or (1,)))
        self._states = np.broadcast_to(common.left(states, len(lead), tail=2), lead + states.shape[-2:]).copy()
        self._density = np.broadcast_to(common.left(self._density, len(lead)), lead).copy()
        self.kvalue = kvalue
        self.tvalue = tvalue
        self.options = options
        self.order1 = {}

    # ---- read-only views mirroring the reference's attributes
    @property
    def states(self):
        return self._states

    @property
    def shape(self):
        return self._states.shape[:-2]

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def nstate(self):
        return (self._states.shape[-2] - 1) // 2

    @property
    def density(self):
        return self._density

    @property
    def equilibrium(self):
        eq = np.zeros(self._states.shape, dtype=np.complex128)
        eq[..., self.nstate, 2] = self._density
        return eq

    @property
    def F(self):
        return self._states[..., :2]

    @property
    def Z(self):
        return self._states[..., 2]

    @property
    def F0(self):
        return self._states[..., self.nstate, 0]

    @property
    def Z0(self):
        return self._states[..., self.nstate, 2]

    @property
    def k(self):
        n = self.nstate
        return np.arange(-n, n + 1, dtype=float)[:, None] * self.kvalue

    def copy(self):
        new = object.__new__(StateMatrix)
        new._states = self._states.copy()
        new._density = self._density.copy()
        new.kvalue, new.tvalue = self.kvalue, self.tvalue
        new.options = dict(self.options)
        new.order1 = dict(self.order1)
        return new

    def __repr__(self):
        return f"StateMatrix(shape={self.shape}, nstate={self.nstate})"
