"""Host-side shape helpers.

The reference broadcasts operator parameters LEFT-ALIGNED against the simulation grid
(`append=True` in epgpy/common.py:273-334): axis 0 of a parameter array is axis 0 of the grid,
missing axes are appended on the right.  The lowering turns that rule into integer strides
(lowering.py), so only shape arithmetic lives here -- there is no array-module dispatch
(the reference's numpy|cupy switch, epgpy/common.py:21-74, has no equivalent: the only backend is
the sm_100a engine).
"""

import numpy as np

DEG = np.pi / 180.0


def isscalar(value):
    """True for anything without a length (epgpy/common.py:236-242)"""
    try:
        len(value)
        return False
    except TypeError:
        return True


def asparam(value, dtype=None):
    """parameter -> python/numpy scalar or ndarray (epgpy/common.py:139-153)"""
    if value is None or isscalar(value):
        return value
    return np.asarray(value) if dtype is None else np.asarray(value, dtype=dtype)


def get_shape(value):
    """shape of a scalar / nested sequence / array"""
    if value is None:
        return ()
    return tuple(np.shape(value))


def broadcast_shapes(*shapes, append=True):
    """common shape, new axes appended (default) or prepended (epgpy/common.py:290-303)"""
    shapes = [tuple(int(d) for d in s) for s in shapes]
    ndim = max([len(s) for s in shapes] + [0])
    if append:
        shapes = [s + (1,) * (ndim - len(s)) for s in shapes]
    else:
        shapes = [(1,) * (ndim - len(s)) + s for s in shapes]
    out = [1] * ndim
    for i in range(ndim):
        dims = {s[i] for s in shapes if s[i] != 1}
        if len(dims) > 1:
            raise ValueError(f"Incompatible shapes: {shapes}")
        if dims:
            out[i] = dims.pop()
    return tuple(out)


def broadcastable(*shapes, append=True):
    try:
        broadcast_shapes(*shapes, append=append)
        return True
    except ValueError:
        return False


def op_shape(*params):
    """left-aligned common shape of operator parameters; at least (1,) (reference: Operator.shape)"""
    shape = broadcast_shapes(*[get_shape(p) for p in params], (1,))
    return shape


def left(arr, ndim, tail=0):
    """append singleton axes so that `arr` (lead axes + `tail` trailing axes) has `ndim` lead axes"""
    arr = np.asarray(arr)
    lead = arr.ndim - tail
    if lead > ndim:
        raise ValueError(f"array with {lead} axes does not fit {ndim} grid axes")
    shape = arr.shape[:lead] + (1,) * (ndim - lead) + arr.shape[lead:]
    return arr.reshape(shape)


def expand_left(*arrays, tail=0):
    """left-aligned expansion of several parameter arrays to a common number of lead axes"""
    ndim = max([np.ndim(a) - tail for a in arrays] + [1])
    return tuple(left(a, ndim, tail) for a in arrays)


def repr_value(value, fmt=""):
    if value is None:
        return "None"
    if isscalar(value):
        try:
            return f"{value:{fmt}}"
        except (TypeError, ValueError):
            return str(value)
    return "(" + "x".join(map(str, get_shape(value))) + ")"


def repr_operator(cls, names, values, fmts=None):
    fmts = fmts or [""] * len(names)
    args = [f"{n}={repr_value(v, f)}" if n else repr_value(v, f) for n, v, f in zip(names, values, fmts) if v is not None]
    return f"{cls}({', '.join(args)})"


def set_axes(params, axes):
    """`axes=` keyword of the reference's operators (epgpy/common.py:336-347, opmatrix.py:148-150): move the
    (left-aligned, mutually broadcast) parameter axes of an operator to the given grid axes.
    `axes` is an int (first axis) or a tuple of ints (one per parameter axis)."""
    arrs = [None if p is None or isscalar(p) else np.asarray(p) for p in params]
    ndim = max([a.ndim for a in arrs if a is not None] + [0])
    if ndim == 0:
        return list(params)
    if isinstance(axes, (int, np.integer)):
        axes = tuple(range(int(axes), int(axes) + ndim))
    elif not isinstance(axes, tuple) or not all(isinstance(ax, (int, np.integer)) for ax in axes):
        raise ValueError(f"Invalid axes: {axes}")
    if len(axes) != ndim:
        raise ValueError(f"Invalid axes: {axes} for {ndim} parameter axes")
    newdims = tuple(i for i in range(max(axes)) if i not in axes)
    out = []
    for p, a in zip(params, arrs):
        out.append(p if a is None else np.expand_dims(left(a, ndim), newdims))
    return out
