"""simulate / get_adc_times / getshape / getnshift / modify -- the reference's driver API
(epgpy/functions.py:14-170, 251-369) on top of the sm_100a engine.

`simulate` lowers the sequence once (lowering.py), uploads the op-tape, launches ONE fused kernel per
device for the whole sequence and all atoms, and copies the ADC samples back.  The reference instead
runs a Python loop with 1-4 array passes per operator and a device->host copy per ADC
(functions.py:173-192, probe.py:63-66).
"""

import logging

import numpy as np

from . import common, engine, lowering
from . import operators as ops
from .lowering import flatten_sequence
from .statematrix import StateMatrix

LOGGER = logging.getLogger(__name__)


def getshape(sequence):
    """overall grid shape defined by a sequence (epgpy/functions.py:14-17)"""
    return common.broadcast_shapes(*[op.shape for op in flatten_sequence(sequence)], append=True)


def getnshift(sequence):
    """number of phase states required by a sequence (epgpy/functions.py:20-26)"""
    return sum(op.nshift for op in flatten_sequence(sequence))


def getkdim(sequence):
    return max([getattr(op, "kdim", 1) for op in flatten_sequence(sequence)] + [1])


def get_adc_times(sequence):
    """ADC opening times: running sum of the operators' durations (epgpy/functions.py:38-47)"""
    tim, times = 0, []
    for op in flatten_sequence(sequence):
        tim = tim + op.duration
        if isinstance(op, ops.Probe):
            times.append(tim)
    return times


def _devices(device):
    import torch

    if device is None:
        return [torch.cuda.current_device()]
    if isinstance(device, (list, tuple)):
        return [int(torch.device(d).index if not isinstance(d, int) else d) for d in device]
    if isinstance(device, int):
        return [device]
    d = torch.device(device)
    return [d.index if d.index is not None else torch.cuda.current_device()]


def _split(begin, count, n):
    per = -(-count // n)
    return [(begin + i * per, min(per, begin + count - (begin + i * per))) for i in range(n) if begin + i * per < begin + count]


def _host_buffer(shape, cdt):
    """output buffer of a run: PINNED host memory (the pitched D2H copies of `Plan.run_to_host` overlap with the
    kernels only into page-locked memory).  torch's caching host allocator keeps the pages when the returned array is
    dropped, so repeated simulations of the same size do not pin again.  Pageable memory is the fallback when the
    host refuses to pin that much."""
    import torch

    try:
        return torch.empty(shape, dtype=cdt, pin_memory=True)
    except RuntimeError:
        return torch.empty(shape, dtype=cdt)


def run_lowered(low, plan=None, device=None, atom_range=None, nchunk=None, keep_device=None, real_output=None):
    """run a lowered sequence on one or several devices of this process: the atom range is cut in one contiguous
    slab per device (no inter-device traffic) and every slab in column chunks, chunk i + 1 being computed while chunk i
    is copied into the pinned host result (engine.Plan.run_to_host).
    Returns (result, plan); result.sig_host / result.jac_host: host arrays complex [nadc][count][npool] and
    [njac][nvar][count][npool] (numpy views of pinned tensors; None when the tape has no such rows), result.parts: the
    device slabs [(device, begin, count, sig_dev, jac_dev)]."""
    import threading

    import torch

    import os
    import time

    engine.require_cuda()
    t0 = time.perf_counter()
    plan = plan or engine.Plan(low)
    t1 = time.perf_counter()
    devs = _devices(device)
    begin, count = atom_range if atom_range is not None else (0, low.natoms)
    if begin < 0 or count < 0 or begin + count > low.natoms:
        raise ValueError(f"atom range ({begin}, {count}) outside the grid of {low.natoms} atoms")
    cdt = torch.complex128 if low.dtype == "f64" else torch.complex64
    csz = 16 if low.dtype == "f64" else 8
    has_jac = bool(low.nvar and low.njac)
    sig_t = _host_buffer((low.nadc, count, low.npool), cdt) if low.nadc else None
    jac_t = _host_buffer((low.njac * low.nvar, count, low.npool), cdt) if has_jac else None
    t2 = time.perf_counter()
    slabs = _split(begin, count, len(devs)) if count else []
    if nchunk is None:  # chunks of >= 32 MB of output, at most 16 per device
        per_dev = low.nbytes_out(natoms=-(-count // max(1, len(slabs)))) if slabs else 0
        nchunk = int(min(16, max(1, per_dev // (32 << 20))))
    parts, errors = [None] * len(slabs), []

    # a real-valued signal crosses PCIe as reals and is widened to complex on the host (engine.Plan.run_to_host_real), unless
    # a row is reduced on the device (that needs the device slab) or the result is too small to matter
    # (with more than four ranks on one host the widening threads of all ranks compete for its memory system and the
    # plain complex copy wins: measured 231 ms against 209 ms per dictionary at 8 ranks, 259 against 284 ms at 2)
    crowded = int(os.environ.get("LOCAL_WORLD_SIZE", "1")) > 4 and real_output is None
    real_path = (real_output is not False and not crowded and not has_jac and plan.real_signal() and low.npool == 1 and sig_t is not None
                 and sig_t.is_pinned() and low.nbytes_out(natoms=count) >= (64 << 20)
                 and not any(r.kind == "sig" and r.reduce is not None for rows in low.rows for r in rows))

    def work(i, d, b, c):
        try:
            col = b - begin
            if real_path:
                plan.run_to_host_real(d, sig_t, b, c, nchunk=max(2 * nchunk, 8), host_col=col, host_atoms=count)
                parts[i] = (d, b, c, None, None)
                return
            sig_d, jac_d = plan.run_to_host(d, sig_t, b, c, nchunk=nchunk, out_jacobian=jac_t, host_col=col, host_atoms=count)
            parts[i] = (d, b, c, sig_d, jac_d)
        except BaseException as ex:  # re-raised in the calling thread
            errors.append(ex)

    if len(slabs) == 1:
        work(0, devs[0], *slabs[0])
    else:
        threads = [threading.Thread(target=work, args=(i, d, b, c)) for i, (d, (b, c)) in enumerate(zip(devs, slabs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    if os.environ.get("EPGX_TIMING"):
        print("epgx run: plan %.1f ms, host buffers %.1f ms, kernels + copies %.1f ms (%d chunks)" %
              (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (time.perf_counter() - t2), nchunk), flush=True)
    sig_host = sig_t.numpy() if sig_t is not None else None
    jac_host = jac_t.numpy().reshape(low.njac, low.nvar, count, low.npool) if jac_t is not None else None
    return RunResult(sig_host, jac_host, [p for p in parts if p is not None]), plan


class RunResult:
    def __init__(self, sig_host, jac_host, parts):
        self.sig_host, self.jac_host, self.parts = sig_host, jac_host, parts


def _assemble(low, res, count=None, asarray=True):
    """host slabs -> per-probe arrays shaped like the reference's output (nADC, *grid).  A probe whose rows are plain
    read-outs comes back as a reshaped VIEW of the [nadc][atoms] result buffer (no copy).  With `count` (a slab of the
    flattened grid, `simulate(shard=...)`) the grid axes are replaced by one flat atom axis."""
    if isinstance(res, list):  # one device slab [(device, begin, count, sig, jac)] (tests/tape_interp.py)
        (_, _, _, sig, jac), = res
        res = RunResult(None if sig is None else sig.cpu().numpy(), None if jac is None else jac.cpu().numpy(), res)
    sig_host, jac_host, parts = res.sig_host, res.jac_host, res.parts
    grid, ax, npool = tuple(low.grid), low.pool_axis, low.npool
    flat_mode = count is not None
    if flat_mode:
        store_shape = (count,) + ((npool,) if ax is not None else ())
        grid = store_shape
        ax = 1 if ax is not None else None
        order = [0] + ([1] if ax is not None else [])
    else:
        store_shape = tuple(d for i, d in enumerate(grid) if i != ax) + ((npool,) if ax is not None else ())
        order = [i for i in range(len(grid)) if i != ax] + ([ax] if ax is not None else [])  # grid axis of each storage axis
    cdt = np.complex128 if low.dtype == "f64" else np.complex64

    def to_grid(flat, lead=0):
        a = flat.reshape(flat.shape[:lead] + store_shape)
        if ax is not None:
            a = np.moveaxis(a, -1, lead + ax)
        return a.reshape(flat.shape[:lead] + grid)

    sig_dev = None
    if low.nadc and any(r.kind == "sig" and r.reduce is not None for rows in low.rows for r in rows):
        # reduction of rows on the device (probe.py:148-153)
        if flat_mode:
            raise NotImplementedError("Adc(reduce=...) over a shard of the grid")
        if len(parts) == 1:
            sig_dev = parts[0][3]
        else:
            import torch

            sig_dev = torch.cat([p[3].to(parts[0][3].device) for p in parts], dim=1)

    values = [[] for _ in range(low.nprobe)]
    plain = [all(rows[ip].kind == "sig" and rows[ip].reduce is None and rows[ip].post is None for rows in low.rows)
             for ip in range(low.nprobe)]
    for ip in range(low.nprobe):
        if not (asarray and plain[ip] and low.rows):
            continue
        idx = [rows[ip].index for rows in low.rows]
        step = idx[1] - idx[0] if len(idx) > 1 else 1
        if step > 0 and all(b - a == step for a, b in zip(idx, idx[1:])):
            values[ip] = to_grid(sig_host[idx[0]:idx[-1] + 1:step].reshape(len(idx), -1), lead=1)  # a view
        else:
            plain[ip] = False
    for rows in low.rows:
        for ip, row in enumerate(rows):
            if asarray and plain[ip]:
                continue
            if row.kind == "sig":
                if row.reduce is None:
                    arr = to_grid(sig_host[row.index].reshape(-1))
                else:
                    t = sig_dev[row.index].reshape(store_shape)
                    red = list(range(len(grid))) if row.reduce is True else sorted({r % len(grid) for r in row.reduce})
                    for s_ax in sorted((order.index(r) for r in red), reverse=True):
                        t = engine.device_reduce(t, s_ax)
                    arr = t.cpu().numpy()
                    keep = [g for g in order if g not in red]
                    arr = np.transpose(arr, np.argsort(keep)) if keep else arr.reshape(())[()]
                if row.post is not None:
                    arr = np.asarray(row.post(arr)).astype(cdt, copy=False)
                values[ip].append(arr)
            elif row.kind == "lin":
                # weighted sum of configuration rows (accumulated-time F0: statematrix.py:149-156)
                nrow, wts, scale = row.jac
                arr = sum(wts[i] * to_grid(sig_host[row.index + i].reshape(-1)) for i in range(nrow))
                if scale is not None:
                    arr = arr * common.left(np.atleast_1d(scale), arr.ndim)
                if row.post is not None:
                    arr = np.asarray(row.post(arr)).astype(cdt, copy=False)
                values[ip].append(np.asarray(arr).astype(cdt, copy=False))
            elif row.kind == "fourier":
                # DFT / Imaging (probe.py:168-219): F [*grid, nslot] read by the device, weights applied here
                nrow, kphys, tacc, probe, system = row.jac
                F = np.stack([to_grid(sig_host[row.index + i].reshape(-1)) for i in range(nrow)], axis=-1)
                val = probe.combine(F, kphys, tacc, system, low.kdim)
                values[ip].append(row.post(val) if row.post is not None else val)
            elif row.kind == "expr":
                val = row.jac._eval(to_grid(sig_host[row.index].reshape(-1)), to_grid(sig_host[row.index + 1].reshape(-1)))
                values[ip].append(row.post(val) if row.post is not None else val)
            elif row.kind == "hess":
                jrow, entries = row.jac
                out = np.zeros(grid + (len(entries), len(entries[0]) if entries else 0), dtype=cdt)
                for i1, ent in enumerate(entries):
                    for i2, vi in enumerate(ent):
                        if vi >= 0 and jrow >= 0:
                            out[..., i1, i2] = to_grid(jac_host[jrow, vi].reshape(-1))
                if row.post is not None:
                    out = np.asarray(row.post(out)).astype(cdt, copy=False)
                values[ip].append(out)
            else:
                jrow, cols = row.jac
                out = np.zeros(grid + (len(cols),), dtype=cdt)
                for ic, (kind, vi) in enumerate(cols):
                    if kind == "mag":
                        out[..., ic] = to_grid(sig_host[row.index].reshape(-1))
                    elif kind == "var":
                        out[..., ic] = to_grid(jac_host[jrow, vi].reshape(-1))
                if row.post is not None:
                    out = np.asarray(row.post(out)).astype(cdt, copy=False)
                values[ip].append(out)
    if asarray:
        return tuple(v if isinstance(v, np.ndarray) else np.asarray(v) for v in values)
    return tuple(tuple(v) for v in values)


def simulate(sequence, *, adc_time=False, init=None, squeeze=False, probe=None, callback=None, asarray=True,
             disp=False, dtype="float64", device=None, propagate_nondiff=False, shard=None, **options):
    """simulate a sequence; values are returned at the probe operators (epgpy/functions.py:50-170)

    Parameters (reference-compatible):
        sequence: (nested) list of operators
        init: None | 3-vector | (2n+1) x 3 array | StateMatrix
        adc_time: also return the ADC opening times
        probe: probe (or list of) superseding the in-sequence ones (None keeps the in-sequence probe)
        asarray: return one ndarray per probe
        **options: state-matrix options: max_nstate, kvalue
    Engine parameters:
        dtype: 'float64' (complex128 states, <= 1e-10 vs the reference) or 'float32' (complex64, ~1e-5)
        device: CUDA device index, or a list of indices: atoms are split in contiguous slabs
        propagate_nondiff: False reproduces the reference, whose D / X / SPOILER never touch the
            order-1 partial states (epgpy/operator.py:96-104); True applies them to the partials too
            (the exact chain rule).
        shard: (index, count): simulate only slab `index` of the `count` contiguous slabs of the flattened grid
            (one process per GPU, sharding.slab); the grid axes of the result are then ONE flat atom axis:
            (nADC, slab atoms).  sharding.simulate adds the final NCCL gather.
    Returns: values | (times, values) -- ndarray (nADC, *grid) per probe, a tuple if several probes
    """
    if squeeze:
        raise NotImplementedError("Automatic sequence squeezing not implemented yet")
    if callback and not isinstance(callback, ops.PartialsPruner):
        raise NotImplementedError("`callback` needs the state matrix on the host after every operator; "
                                  "the fused device path has no such hook (a PartialsPruner is accepted: see its docstring)")
    import os
    import time

    tm = [time.perf_counter()]
    low = lowering.lower(sequence, init=init, probe=probe, options=options, dtype=engine.norm_dtype(dtype),
                         propagate_nondiff=propagate_nondiff)
    tm.append(time.perf_counter())
    LOGGER.info("Simulate sequence: num. operators: %d, shape: %s, max order: %d", len(low.ops), low.grid, low.max_order)
    atom_range = None
    if shard is not None:
        from . import sharding

        atom_range = sharding.slab(low.natoms, int(shard[0]), int(shard[1]))
    res, plan = run_lowered(low, device=device, atom_range=atom_range)
    tm.append(time.perf_counter())
    values = _assemble(low, res, count=None if shard is None else atom_range[1], asarray=asarray)
    tm.append(time.perf_counter())
    if os.environ.get("EPGX_TIMING"):  # host-side breakdown of one call (ms): lowering | plan + upload + kernels + D2H | assembly
        print("epgx simulate: lower %.1f ms, run %.1f ms, assemble %.1f ms" % tuple(1e3 * (b - a) for a, b in zip(tm, tm[1:])), flush=True)
    times = np.asarray(low.times) if asarray else low.times
    if len(values) == 1:
        values = values[0]
    if adc_time:
        return times, values
    return values


def apply_operators(operators, sm, *, dtype="float64", device=None):
    """apply operators to a StateMatrix and return the new StateMatrix (the reference's `op(sm)`,
    epgpy/operator.py:96-104): a probe-less tape on the engine whose final state is read back
    (epgx_simulate_state; half storage -> the reference's [*grid, 2n+1, 3] by the symmetry of
    statematrix.py:418-421).  The result can seed a later `simulate(init=sm)` (functions.py:133-144).
    Order-1 partial states are not carried from one call to the next: differentiate inside one `simulate`."""
    from .operators import PD, DiffOperator

    seq = flatten_sequence(operators)
    if any(isinstance(op, DiffOperator) and (getattr(op, "order1", None) or {}) for op in seq):
        raise NotImplementedError("operators with order-1 variables applied outside `simulate`: the partial states stay on "
                                  "the device; put the operators and a Jacobian probe in one `simulate` call")
    if not isinstance(sm, StateMatrix):
        sm = StateMatrix(sm)
    low = lowering.lower(seq, init=sm, dtype=engine.norm_dtype(dtype), prune_unobservable=False, need_probe=False)
    plan = engine.Plan(low)
    dev = _devices(device)[0]
    _, _, state = plan.run_state(dev)
    half = state.cpu().numpy().astype(np.complex128)  # [atoms][npool][C][3]
    n = low.final_n
    half = half[:, :, :n + 1, :]
    grid, ax = tuple(low.grid), low.pool_axis
    half = half.reshape(tuple(low.atom_shape) + half.shape[1:])
    half = np.moveaxis(half, len(low.atom_shape), ax) if ax is not None else half[..., 0, :, :]
    half = half.reshape(grid + (n + 1, 3))
    full = np.zeros(grid + (2 * n + 1, 3), dtype=np.complex128)
    full[..., n:, :] = half
    if n:
        full[..., :n, 0] = half[..., :0:-1, 1].conj()
        full[..., :n, 1] = half[..., :0:-1, 0].conj()
        full[..., :n, 2] = half[..., :0:-1, 2].conj()
    density = sm.density
    for op in seq:  # PD replaces the equilibrium density (operator.py:315-341)
        if isinstance(op, PD):
            density = np.asarray(op.pd, dtype=float)
    new = StateMatrix(full, density=np.broadcast_to(common.left(np.atleast_1d(density), len(grid)), grid), kvalue=sm.kvalue,
                      tvalue=sm.tvalue, check=False, **sm.options)
    return new


# --------------------------------------------------------------------------------------------- #
# modify (epgpy/functions.py:251-347): pure host helper
# --------------------------------------------------------------------------------------------- #


def _with_b1(op, att):
    """RF pulse with its flip angle scaled by the B1 attenuation `att` (marked '#')"""
    if att is None or not isinstance(op, ops.T) or np.allclose(att, 1):
        return op
    return ops.T(op.alpha * att, op.phi, name=op.name + "#", duration=op.duration)


def _with_evolution(op, T1, T2, g):
    """timed operator followed by what its duration does to the spins: precession alone when only `g` is known,
    relaxation (missing times read as 1e10 ms, i.e. none) + precession otherwise (marked '*')"""
    if (T1 is None and T2 is None and g is None) or not np.any(np.asarray(op.duration) > 0):
        return op
    if T1 is None and T2 is None:
        follow = ops.P(op.duration, g, duration=0)
    else:
        follow = ops.E(op.duration, 1e10 if T1 is None else T1, 1e10 if T2 is None else T2, 0 if g is None else g, duration=0)
    out = op * follow
    out.name = out[0].name + "*"
    return out


def default_modifier(op, **kwargs):
    """what `modify` does to one operator by default (behaviour of epgpy/functions.py:310-347): keyword `att` scales the
    flip angle of RF pulses, keywords `T1` / `T2` / `g` append relaxation / precession over the operator's duration"""
    return _with_evolution(_with_b1(op, kwargs.get("att")), kwargs.get("T1"), kwargs.get("T2"), kwargs.get("g"))


def modify(sequence, modifier=None, *, expand=True, **params):
    """insert duration-dependent relaxation / precession after timed operators (epgpy/functions.py:251-307)"""
    shape = getshape(sequence)
    names = list(params)
    values = [common.asparam(v) for v in params.values()]
    if values:
        nd = max([np.ndim(v) for v in values] + [0])
        values = [v if common.isscalar(v) else common.left(v, nd) for v in values]
    if expand and (len(shape) > 1 or shape[0] > 1):
        dims = tuple(range(len(shape)))
        values = [v if common.isscalar(v) else np.expand_dims(v, dims) for v in values]
    params = dict(zip(names, values))
    if not modifier:
        modifier = default_modifier
        if not params:
            return sequence
    elif not callable(modifier):
        raise TypeError("`modifier` must be a callable")
    newseq, opdict = [], {}
    for op in flatten_sequence(sequence):
        if id(op) not in opdict:
            opdict[id(op)] = modifier(op, **params)
        newseq.append(opdict[id(op)])
    if isinstance(sequence, ops.MultiOperator):
        return ops.MultiOperator(newseq, name=sequence.name)
    return newseq
