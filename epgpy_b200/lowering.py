"""Lowering: flattened operator sequence -> op-tape for the sm_100a engine (include/epgx.h).

Replaces the reference's interpretation strategy (one numpy/cupy array expression per operator per
TR, epgpy/functions.py:173-192) by a one-off compilation on the host:

  * grid shape        left-aligned broadcast of all operator shapes (functions.py:14-17)
  * coefficient table one float64 block per operator parameter group, in the group's own
                      un-broadcast shape; the left-aligned broadcasting rule becomes per-axis strides
                      ("patterns")
  * records           one per operator application (+ derivative injections, diff.py:264-288)
  * segments          runs of records between two unit shifts; S(k) -> |k| unit shifts
                      (shift.py:86-101), max_nstate truncation included
  * order schedule    n_old / n_new / nact per segment, a pure function of the S operators.  Orders
                      that can no longer reach k = 0 before the last ADC are not updated
                      (`prune_unobservable`): their values never enter any returned sample.
"""

import numpy as np

from . import common, operators as ops_mod
from .exchange import X
from .operators import (PD, Adc, D, DiffOperator, EmptyOperator, Hessian, Jacobian, MultiOperator, Operator, Pair, Probe,
                        Reset, S, Spoiler, reduce_pulse)
from .statematrix import StateMatrix

# opcodes / flags (include/epgx.h)
(OP_NOP, OP_T_GEN, OP_T_RE, OP_T_IM, OP_E, OP_DIAG, OP_MATRIX, OP_D, OP_X, OP_SPOIL, OP_PD, OP_ADC, OP_FUSED,
 OP_CONT) = range(14)
F_BASE, F_PARTIALS, F_INJECT, F_G, F_AFFINE, F_Z0, F_SCALE, F_PRE, F_POST, F_IM, F_GEN, F_P1, F_P2, F_SLOT = (1 << i for i in range(14))
SEG_RESET, SEG_MASK_TOP, SEG_LATTICE = 1, 2, 4
MAX_DIMS, MAX_PATTERNS, MAX_POOLS = 8, 64, 4

OP_DTYPE = np.dtype([("code", "<u2"), ("flags", "<u2"), ("aux", "<i4"), ("off", "<u4", (3,)), ("pat", "u1", (3,)),
                     ("rsv", "u1"), ("aux1", "<i4"), ("rsv1", "<i4")])
SEG_DTYPE = np.dtype([("first", "<i4"), ("count", "<i4"), ("nact", "<i4"), ("shift", "<i4"), ("n_old", "<i4"),
                      ("n_new", "<i4"), ("flags", "<i4"), ("rsv", "<i4")])
assert OP_DTYPE.itemsize == 32 and SEG_DTYPE.itemsize == 32


def flatten_sequence(seq, flatten_multi=True):
    """flat list of operators from nested lists / MultiOperators (epgpy/functions.py:355-369)"""
    seq = [seq] if isinstance(seq, Operator) else seq
    out = []
    for item in seq:
        if isinstance(item, (list, tuple)):
            out.extend(flatten_sequence(item))
        elif flatten_multi and isinstance(item, MultiOperator):
            out.extend(flatten_sequence(item.operators))
        elif isinstance(item, Operator):
            out.append(item)
        else:
            raise ValueError(f"Invalid operator: {item}")
    return out


class Row:
    """one output row (one probe at one ADC)"""

    def __init__(self, kind, index, reduce=None, post=None, jac=None):
        self.kind, self.index, self.reduce, self.post, self.jac = kind, index, reduce, post, jac


class Lowered:
    """the lowered sequence + everything needed to assemble the outputs"""

    def nbytes_out(self, natoms=None, dtype=None):
        natoms = self.natoms if natoms is None else natoms
        csz = 16 if (dtype or self.dtype) == "f64" else 8
        return csz * natoms * self.npool * (self.nadc + self.njac * self.nvar)


class _Builder:
    """coefficient table, patterns and records of one lowering.  Records are plain tuples
    (code, flags, aux, off0, off1, off2, pat0, pat1, pat2, aux1) until `records_array` packs them: the 5 002-operator
    FISP sequence lowers in tens of milliseconds instead of the 0.6 s the per-record numpy scalars took."""

    def __init__(self, grid, pool_axis):
        self.grid = tuple(grid)
        self.pool_axis = pool_axis
        self.atom_axes = [i for i in range(len(grid)) if i != pool_axis]
        self.chunks, self.ncoef = [], 0
        self.patterns = {}
        self.records = []
        self.cache = {}
        self._shape_pat = {}  # block shape -> pattern index (blocks are C-contiguous)

    def pattern(self, strides, pool_stride):
        key = (tuple(int(s) for s in strides), int(pool_stride))
        if key not in self.patterns:
            if len(self.patterns) >= MAX_PATTERNS:
                raise NotImplementedError(f"more than {MAX_PATTERNS} distinct parameter broadcast patterns in one sequence")
            self.patterns[key] = len(self.patterns)
        return self.patterns[key]

    def _pattern_of_shape(self, shape):
        lead = shape[:-1]
        if len(lead) > len(self.grid):
            raise ValueError(f"Incompatible shapes: parameter {lead} vs grid {self.grid}")
        es, acc = [0] * len(shape), 1
        for i in range(len(shape) - 1, -1, -1):  # element strides of a C-contiguous array
            es[i] = acc
            acc *= shape[i]
        strides, pool_stride = [], 0
        for i in range(len(self.grid)):
            st = 0
            if i < len(lead) and lead[i] != 1:
                if lead[i] != self.grid[i]:
                    raise ValueError(f"Incompatible shapes: parameter {lead} vs grid {self.grid}")
                st = es[i]
            if i == self.pool_axis:
                pool_stride = st
            else:
                strides.append(st)
        return self.pattern(strides, pool_stride)

    def block(self, arr):
        """register a coefficient block `lead + (entry,)`; returns (offset, pattern)"""
        if not (type(arr) is np.ndarray and arr.dtype == np.float64 and arr.flags.c_contiguous):
            arr = np.ascontiguousarray(arr, dtype=np.float64)
        pat = self._shape_pat.get(arr.shape)
        if pat is None:
            pat = self._shape_pat[arr.shape] = self._pattern_of_shape(arr.shape)
        off = self.ncoef
        self.chunks.append(arr.reshape(-1))
        self.ncoef += arr.size
        return off, pat

    def record(self, code, flags, blocks=(), aux=0, aux1=0):
        off, pat = [0, 0, 0], [0, 0, 0]
        for i, b in enumerate(blocks):
            if b is not None:
                off[i], pat[i] = b
        self.records.append((code, flags, aux, off[0], off[1], off[2], pat[0], pat[1], pat[2], aux1))
        return len(self.records) - 1


# record tuple fields
R_CODE, R_FLAGS, R_AUX, R_OFF, R_PAT, R_AUX1 = 0, 1, 2, 3, 6, 9


def records_array(records):
    """tuple records -> structured array in the layout of include/epgx.h (epgx_op)"""
    out = np.zeros(len(records), dtype=OP_DTYPE)
    if records:
        a = np.array(records, dtype=np.int64).reshape(len(records), 10)
        out["code"], out["flags"], out["aux"], out["aux1"] = a[:, 0], a[:, 1], a[:, 2], a[:, 9]
        out["off"], out["pat"] = a[:, 3:6], a[:, 6:9]
    return out


def _scale_block(blk, coeff):
    coeff = np.asarray(coeff)
    if coeff.ndim == 0:
        return blk * float(coeff) if not np.iscomplexobj(coeff) else _cmul_block(blk, coeff)
    if np.iscomplexobj(coeff):
        raise NotImplementedError("complex chain-rule coefficients")
    nd = max(blk.ndim - 1, coeff.ndim)
    return common.left(blk, nd, tail=1) * common.left(coeff, nd)[..., None]


def _cmul_block(blk, z):
    raise NotImplementedError("complex chain-rule coefficients")


def _emit_form(bld, form, flags, aux=0, aux1=0):
    """emit one record for a coefficient form (aux / aux1: target variable and source set of an injection)"""
    kind = form[0]
    if kind == "tgen":
        form = reduce_pulse(form)
        kind = form[0]
    if kind == "tre":
        bld.record(OP_T_RE, flags, [bld.block(form[1])], aux, aux1)
    elif kind == "tim":
        bld.record(OP_T_IM, flags, [bld.block(form[1])], aux, aux1)
    elif kind == "tgen":
        bld.record(OP_T_GEN, flags, [bld.block(form[1])], aux, aux1)
    elif kind == "e":
        _, b0, b1, b2, affine = form
        fl = flags | (F_G if b2 is not None else 0) | (F_AFFINE if affine else 0)
        bld.record(OP_E, fl, [bld.block(b0), bld.block(b1), None if b2 is None else bld.block(b2)], aux, aux1)
    elif kind == "diag":
        _, blk, affine = form
        bld.record(OP_DIAG, flags | (F_AFFINE if affine else 0), [bld.block(blk)], aux, aux1)
    elif kind == "matrix":
        _, blk, blk0 = form
        bld.record(OP_MATRIX, flags | (F_AFFINE if blk0 is not None else 0),
                   [bld.block(blk), None if blk0 is None else bld.block(blk0)], aux, aux1)
    else:
        raise ValueError(kind)


def _scaled_form(form, coeff):
    kind = form[0]
    if np.ndim(coeff) == 0 and coeff == 1:
        return form
    if kind == "tgen":
        return ("tgen", _scale_block(form[1], coeff))
    if kind == "diag":
        return ("diag", _scale_block(form[1], coeff), form[2])
    if kind == "matrix":
        return ("matrix", _scale_block(form[1], coeff), None if form[2] is None else _scale_block(form[2], coeff))
    raise ValueError(kind)


def _unit_vector(vectors):
    """common integer base vector of collinear integer shift vectors"""
    base = None
    for v in vectors:
        v = np.asarray(v)
        if not np.issubdtype(v.dtype, np.integer):
            raise NotImplementedError(
                "float shifts (the reference's shift-merge / shift-prune methods, epgpy/shift.py:367-542) are outside the hot path"
            )
        if v.shape[:-1] not in ((), (1,)):
            raise NotImplementedError("per-atom shift vectors are outside the hot path")
        v = v.reshape(-1)
        if base is None:
            base = v // np.gcd.reduce(np.abs(v))
    return base


def _multiple(v, base):
    v = np.asarray(v).reshape(-1)
    if len(v) < len(base):
        v = np.pad(v, (0, len(base) - len(v)))
    elif len(v) > len(base):
        raise NotImplementedError("shift vectors of different dimensions in one sequence")
    i = int(np.argmax(np.abs(base)))
    m, r = divmod(int(v[i]), int(base[i]))
    if r or not np.array_equal(m * base, v):
        raise NotImplementedError(
            "non-collinear n-d shifts (the reference's general shift-nd method, epgpy/shift.py:297-364) are outside the hot path"
        )
    return m


def fuse_records(recs, segs):
    """peephole pass ("squeeze", a stub in the reference: epgpy/functions.py:350-352): runs of
    [E] [T_RE | T_IM] [E] without precession or derivatives become one FUSED + CONT record pair, applied
    in a single sweep over the orders.  An E that closes a segment acts identically on every order, so
    it commutes with the unit shift and is moved to the head of the next segment when a T waits there.
    recs: tuple records; segs: lists [first, count, nact, shift, n_old, n_new, flags] (updated in place)."""
    pulses = (OP_T_RE, OP_T_IM, OP_T_GEN)

    def pure_e(r):
        return r[0] == OP_E and (r[1] & ~F_AFFINE) == F_BASE

    def pulse(r):
        return r[0] in pulses and r[1] == F_BASE

    def diffusion(r):
        return r[0] == OP_D and r[1] == F_BASE

    out, carry = [], []
    nseg = len(segs)
    for i in range(nseg):
        seg = segs[i]
        rs = recs[seg[0]:seg[0] + seg[1]]
        if carry:  # a diagonal E commutes with the (diagonal) D records that open the segment: put it next to the pulse
            nd = 0
            while nd < len(rs) and diffusion(rs[nd]):
                nd += 1
            rs = rs[:nd] + carry + rs[nd:]
        carry = []
        if seg[3] != 0 and not (seg[6] & SEG_RESET) and i + 1 < nseg and rs and pure_e(rs[-1]):
            nxt = recs[segs[i + 1][0]:segs[i + 1][0] + segs[i + 1][1]]
            nd = 0
            while nd < len(nxt) and diffusion(nxt[nd]):
                nd += 1
            if nd < len(nxt) and pulse(nxt[nd]):
                carry = [rs.pop()]
        new, j, n = [], 0, len(rs)
        while j < n:
            pre = None
            if pure_e(rs[j]) and j + 1 < n and pulse(rs[j + 1]):
                pre, j = rs[j], j + 1
            if pulse(rs[j]) and (pre is not None or (j + 1 < n and pure_e(rs[j + 1]))):
                t = rs[j]
                post = rs[j + 1] if j + 1 < n and pure_e(rs[j + 1]) else None
                fl = F_BASE | (F_IM if t[0] == OP_T_IM else 0) | (F_GEN if t[0] == OP_T_GEN else 0) \
                    | (F_PRE if pre is not None else 0) | (F_POST if post is not None else 0)
                po, pp = (pre[3:5], pre[6:8]) if pre is not None else ((0, 0), (0, 0))
                qo, qp = (post[3:5], post[6:8]) if post is not None else ((0, 0), (0, 0))
                new.append((OP_FUSED, fl, 0, t[3], po[0], po[1], t[6], pp[0], pp[1], 0))
                new.append((OP_CONT, 0, 0, qo[0], qo[1], 0, qp[0], qp[1], 0, 0))
                j += 2 if post is not None else 1
            else:
                new.append(rs[j])
                j += 1
        seg[0], seg[1] = len(out), len(new)
        out += new
    return out, segs


class _Lattice:
    """the configuration lattice of general integer n-d shifts (the reference's `shift-nd` method,
    epgpy/shift.py:103-117, 297-364).  Shifts are the same for every atom, so the set of configurations k that can be
    populated after each operator is a pure function of the sequence: the host walks it ONCE, keeps one slot per
    lattice point (full storage: both k and -k) and turns every shift into three gather maps for the device
    (F+ moves to k + dk, F- to k - dk, Z stays; epgx.h EPGX_SEG_LATTICE).  Where the reference prunes rows whose values
    fall below a tolerance after every shift (data dependent, 1e-8), the lattice drops the points that are
    STRUCTURALLY empty -- never reached by any pathway -- which is exact; cropping at max_nstate is the reference's."""

    def __init__(self, kdim):
        self.kdim = kdim
        self.zero = (0,) * kdim
        self.coords = [self.zero]
        self.occ = {self.zero: [True, True, True]}  # per point: F+, F-, Z possibly non-zero

    def mix(self):  # a 3 x 3 operator couples the three components of a point
        for o in self.occ.values():
            if any(o):
                o[0] = o[1] = o[2] = True

    def spoil(self):
        for o in self.occ.values():
            o[0] = o[1] = False

    def reset(self):
        self.coords = [self.zero]
        self.occ = {self.zero: [False, False, True]}

    def shift(self, kvec, nmax):
        """apply S(kvec); returns (mapP, mapM, mapZ): source slot of every new slot (-1: empty)"""
        kvec = tuple(int(x) for x in kvec) + (0,) * (self.kdim - len(kvec))
        new = {self.zero: [False, False, False]}
        for c, (hp, hm, hz) in self.occ.items():
            if hp:
                new.setdefault(tuple(a + b for a, b in zip(c, kvec)), [False, False, False])[0] = True
            if hm:
                new.setdefault(tuple(a - b for a, b in zip(c, kvec)), [False, False, False])[1] = True
            if hz:
                new.setdefault(c, [False, False, False])[2] = True
        if nmax is not None:  # crop (shift.py:331-343): points with a component above nmax fall off
            new = {c: o for c, o in new.items() if all(abs(x) <= nmax for x in c)}
        index = {c: i for i, c in enumerate(self.coords)}
        coords = sorted(new)
        maps = ([], [], [])
        for c in coords:
            srcs = (tuple(a - b for a, b in zip(c, kvec)), tuple(a + b for a, b in zip(c, kvec)), c)
            for comp in range(3):
                src = srcs[comp]
                maps[comp].append(index[src] if (src in self.occ and self.occ[src][comp]) else -1)
        self.coords, self.occ = coords, new
        return maps

    @property
    def kzero(self):
        return self.coords.index(self.zero)


def lower(sequence, *, init=None, probe=None, options=None, dtype="f64", propagate_nondiff=False,
          prune_unobservable=True, fuse=True, pre_inject=True, need_probe=True):
    """sequence -> Lowered.  need_probe=False: a tape without read-out (functions.apply_operators reads the final
    state back instead; it also passes prune_unobservable=False so that every order is kept up to date)"""
    options = dict(options or {})
    seq = flatten_sequence(sequence)
    if need_probe and not any(isinstance(op, Probe) for op in seq):
        raise ValueError("Cannot simulate sequence without at least one Probe/ADC operator")
    # ---- initial state
    if init is None:
        init = [0, 0, 1]
    if isinstance(init, StateMatrix):
        sm = init
        options = {**sm.options, **options}
        kvalue = options.pop("kvalue", sm.kvalue)
    else:
        kvalue = options.pop("kvalue", 1.0)
        sm = StateMatrix(init)
    max_nstate = options.pop("max_nstate", None) or None
    lattice_opt = bool(options.pop("lattice", False))  # force the general lattice path (tests: it must reproduce the 1-d one)
    kgrid_opt = options.pop("kgrid", None)
    tvalue = options.pop("tvalue", None)
    if tvalue is None:
        tvalue = getattr(sm, "tvalue", 1.0)
    options.pop("prune", None)
    if options:
        raise TypeError(f"unknown simulate option(s): {sorted(options)}")

    # ---- probes (functions.py:118-127, 184-189)
    probes = None
    if probe:
        probes = list(probe) if isinstance(probe, (tuple, list)) else [probe]
        probes = [pb if isinstance(pb, (Probe, type(None))) else Probe(pb) for pb in probes]

    # ---- grid
    grid = common.broadcast_shapes(*[op.shape for op in seq], sm.shape, append=True)
    if len(grid) > MAX_DIMS:
        raise NotImplementedError(f"more than {MAX_DIMS} grid axes")
    xops = [op for op in seq if isinstance(op, X)]
    pool_axis, npool = None, 1
    if xops:
        axes = {(op.axis, op.ncomp) for op in xops}
        if len(axes) > 1:
            raise NotImplementedError("exchange operators along different axes in one sequence")
        pool_axis, npool = axes.pop()
        if npool > MAX_POOLS:
            raise NotImplementedError(f"exchange between more than {MAX_POOLS} compartments")
        if pool_axis >= len(grid) or grid[pool_axis] != npool:
            raise RuntimeError("Invalid state matrix shape")
    atom_shape = tuple(d for i, d in enumerate(grid) if i != pool_axis) or (1,)
    bld = _Builder(grid, pool_axis)

    # ---- derivative variables: only those a Jacobian / Hessian probe asks for
    jprobes = [pb for pb in (probes or []) if isinstance(pb, Jacobian)]
    hprobes = [pb for pb in (probes or []) if isinstance(pb, Hessian)]
    if probes is None or any(pb is None for pb in probes):  # a None entry keeps the in-sequence probes, Jacobians included
        jprobes += [op for op in seq if isinstance(op, Jacobian)]
        hprobes += [op for op in seq if isinstance(op, Hessian)]
    defined = []
    for op in seq:
        for var in getattr(op, "order1", None) or {}:
            if var not in defined:
                defined.append(var)
    wanted_pairs = {pair for pb in hprobes for pair in pb.pairs() if pair[0] in defined and pair[1] in defined}
    wanted = {v for pb in jprobes for v in pb.variables if v != "magnitude"}
    wanted |= {v for pb in hprobes for v in pb.first()} | {v for pair in wanted_pairs for v in pair}
    variables = [v for v in defined if v in wanted]
    vindex = {v: i for i, v in enumerate(variables)}
    nvar1 = len(variables)
    # order-2 partial states: one more state set per unordered pair of variables, numbered behind the order-1 ones
    # (epgpy/diff.py:290-378).  The shared-memory kernel keeps three partial states resident per tile: (a, b, ab)
    pairs = sorted(wanted_pairs, key=lambda pr: (vindex[pr[0]], vindex[pr[1]]))
    pindex = {pair: nvar1 + i for i, pair in enumerate(pairs)}
    nvar = nvar1 + len(pairs)
    tiles = None
    if pairs:
        tiles = [[vindex[a], vindex[b] if b != a else pindex[(a, b)], pindex[(a, b)] if b != a else -1] for a, b in pairs]
        covered = {v for pair in pairs for v in pair}
        rest = [vindex[v] for v in variables if v not in covered]
        tiles += [(rest[i:i + 3] + [-1, -1])[:3] for i in range(0, len(rest), 3)]

    # ---- shifts: 1-d integers, collinear integer vectors (exactly the 1-d problem) -- or a general integer lattice
    vecs = [op.k for op in seq if isinstance(op, S) and not common.isscalar(op.k)]
    lattice = lattice_opt or any(isinstance(pb, (ops_mod.DFT, ops_mod.Imaging)) for pb in list(seq) + list(probes or []))
    base = None
    # float shifts (the reference's `shift-merge`, shift.py:119-145, 367-444): wavenumbers quantised on a grid `kgrid`.
    # Shifts that are multiples of the grid never merge two different wavenumbers, and the method is then exactly the
    # integer lattice in grid units -- that is the part lowered here; other float shifts merge states approximately
    # (weighted mean wavenumbers, data dependent) and are refused.
    # The reference quantises WAVENUMBERS, i.e. shifts times ktvalue = (kvalue, kvalue, kvalue, tvalue)
    # (statematrix.py:203-211; shift.py:136-141): the grid units of a shift are k * ktvalue / kgrid.
    kscale = ktv = None
    if kgrid_opt is not None or any(not np.issubdtype(np.asarray(v).dtype, np.integer) for v in vecs):
        grids = [kgrid_opt if kgrid_opt is not None else op.kgrid for op in seq if isinstance(op, S)]
        if any(g is None for g in grids):
            raise AttributeError("kgrid not set")
        kdim_f = max([1] + [np.asarray(v).shape[-1] for v in vecs])
        kscale = np.broadcast_to(np.asarray(grids[0], dtype=float), (kdim_f,)).copy()
        if any(not np.allclose(np.broadcast_to(np.asarray(g, dtype=float), (kdim_f,)), kscale) for g in grids):
            raise NotImplementedError("different kgrid values in one sequence")
        kv3 = [float(kvalue)] * 3 if common.isscalar(kvalue) else [float(x) for x in list(kvalue)[:3]]
        ktv = np.asarray(kv3[:min(kdim_f, 3)] + [float(tvalue)] * (kdim_f == 4))
        lattice = True

    def grid_units(k):
        """shift in lattice (grid) units"""
        kv = np.atleast_1d(np.asarray(k if not common.isscalar(k) else [k])).reshape(-1)
        if kscale is None:
            return kv
        q = kv * ktv[:len(kv)] / kscale[:len(kv)]
        if not np.allclose(q, np.round(q), atol=1e-6):
            raise NotImplementedError("float shifts whose wavenumbers (k * kvalue) are not multiples of kgrid merge states "
                                      "approximately (shift-merge / shift-prune, epgpy/shift.py:367-542): outside the hot path")
        return np.round(q).astype(int)

    if vecs and not lattice:
        try:
            base = _unit_vector(vecs)
            for op in seq:
                if isinstance(op, S):
                    _multiple([op.k] if common.isscalar(op.k) else op.k, base)
        except NotImplementedError as ex:
            if "non-collinear" not in str(ex) and "different dimensions" not in str(ex):
                raise
            lattice, base = True, None
    if lattice:
        for v in vecs:
            if np.asarray(v).shape[:-1] not in ((), (1,)):
                raise NotImplementedError("lattice shifts need one vector per operator (per-atom shift vectors: shift-prune, "
                                          "epgpy/shift.py:478-542, are outside the hot path)")
            if kscale is None and not np.issubdtype(np.asarray(v).dtype, np.integer):
                raise AttributeError("kgrid not set")

    def shift_count(op):
        if base is None:
            return int(op.k)
        return _multiple([op.k] if common.isscalar(op.k) else op.k, base)

    # ---- lattice mode: walk the configuration lattice once (slot counts, gather maps of every shift)
    lat_steps, lat = [], None
    if lattice:
        if sm.nstate != 0:
            raise NotImplementedError("lattice shifts start from a state matrix without populated orders (nstate = 0)")
        kdim_l = max([1] + [np.asarray(op.k).shape[-1] for op in seq if isinstance(op, S) and not common.isscalar(op.k)])
        lat = _Lattice(kdim_l)
        nmax_pts = 1
        for op in seq:
            if isinstance(op, S):
                kv = grid_units(op.k)
                cap = (max_nstate or op.nmax or None) if kscale is None else None  # (shift-merge does not crop)
                kz_before, n_before = lat.kzero, len(lat.coords)
                maps = lat.shift(kv, cap)
                lat_steps.append((maps, kz_before, n_before, list(lat.coords)))
                nmax_pts = max(nmax_pts, len(lat.coords))
            elif isinstance(op, Reset) or (isinstance(op, PD) and op.reset):
                lat.reset()
            elif isinstance(op, Spoiler):
                lat.spoil()
            elif isinstance(op, DiffOperator) and not isinstance(op, (ops_mod.E, ops_mod.P, ops_mod.R, ops_mod.Phi, ops_mod.ScalarOp)):
                lat.mix()
        lat = _Lattice(kdim_l)  # replayed by the main walk

    # ---- order schedule: maximum order of the whole tape
    init_n = sm.nstate
    crop = 0
    if max_nstate and init_n > max_nstate and any(isinstance(op, S) for op in seq):
        # the first shift crops the state to max_nstate orders (shift.py:98, statematrix.resize); the orders above
        # never reach k = 0 before that, so cropping the initial state gives the same read-outs
        crop, init_n = init_n - max_nstate, max_nstate
    n, max_order = init_n, init_n
    for op in seq:
        if lattice:
            break
        if isinstance(op, S):
            cap = max_nstate or op.nmax or None
            n = n + abs(shift_count(op)) if cap is None else min(n + abs(shift_count(op)), max(cap, 0))
            max_order = max(max_order, n)
        elif isinstance(op, Reset):
            n = 0
    if init_n > max_order:
        max_order = init_n
    if lattice:
        max_order = nmax_pts - 1  # slots 0 .. max_order hold the lattice points

    # ---- init / equilibrium blocks (half storage: orders 0..init_n)
    st = sm.states[..., crop:sm.states.shape[-2] - crop, :]
    half = st[..., init_n:, :]
    init_blk = np.stack([half[..., 0].real, half[..., 0].imag, half[..., 1].real, half[..., 1].imag,
                         half[..., 2].real, half[..., 2].imag], axis=-1).reshape(half.shape[:-2] + (6 * (init_n + 1),))
    init_ref = bld.block(init_blk)
    m0_ref = bld.block(np.asarray(sm.density, dtype=float)[..., None])

    # ---- wavenumbers for D (diffusion.py:60-79): K(m) = m * base * kvalue [rad/m]
    kdim = 1 if base is None else len(base)
    bvec = np.ones(1) if base is None else base.astype(float)

    def diffusion_block(op, coords=None):
        nonlocal kdim
        m = np.arange(0, max_order + 1, dtype=float)
        tau = np.asarray(op.tau, dtype=float) * 1e-3
        Kp = m[:, None] * bvec[None, :] * kvalue * 1e-3   # order +m
        if coords is not None:  # lattice mode: one row per slot, the wavenumber of its lattice point (unused slots: 0)
            kdim = min(len(coords[0]), 3)  # (a fourth coordinate is accumulated time: no diffusion weighting)
            Kp = np.zeros((max_order + 1, kdim))
            cs = np.asarray(coords, dtype=float)[:, :kdim]
            # integer lattice: coordinates are shift counts (x kvalue); float lattice: grid units of WAVENUMBERS (x kgrid)
            Kp[:len(coords)] = cs * (kvalue if kscale is None else kscale[:kdim]) * 1e-3
        if op.k is None:
            sh = np.zeros(kdim)
        else:
            sh = np.asarray(op.k, dtype=float).reshape(-1)
            if coords is not None and len(sh) < kdim:
                sh = np.pad(sh, (0, kdim - len(sh)))
            if len(sh) != kdim:
                raise ValueError("Incompatible numbers of dimensions for k1 and k2")
            sh = sh * kvalue * 1e-3
        Dm = np.asarray(op.D, dtype=float)
        if Dm.ndim == 0:
            Dm = float(Dm) * np.eye(kdim)
        elif Dm.shape[-2:] != (kdim, kdim) or Dm.ndim != 2:
            raise NotImplementedError("D must be a scalar or one kdim x kdim matrix per operator")

        def trace_bd(k2, shv):
            """Tr(b D) / tau of a linear change (k2 - shv) -> k2  (diffusion.py:86-123)"""
            k1 = k2 - shv
            kd = k2 - k1
            q = np.einsum("mi,ij,mj->m", k1, Dm, k1)
            if np.allclose(kd, 0):
                return q
            return q + 0.5 * np.einsum("mi,ij,mj->m", k1, Dm, kd) + 0.5 * np.einsum("mi,ij,mj->m", kd, Dm, k1) \
                + np.einsum("mi,ij,mj->m", kd, Dm, kd) / 3

        tP = trace_bd(Kp, sh)     # F+ at order +m
        tM = trace_bd(-Kp, sh)    # F-(m) carries the F+ factor of order -m
        tL = np.einsum("mi,ij,mj->m", Kp, Dm, Kp)
        rows = np.stack([tP, tM, tL], axis=-1).reshape(-1)      # [3 * (max_order + 1)]
        tau = np.atleast_1d(tau)
        return np.exp(-tau[..., None] * rows)

    # ---- walk the sequence
    segs = []
    rows_out = []          # per ADC op: list of Row (one per probe)
    times, tic = [], 0
    nadc = njac = 0
    n = init_n
    alive = False          # any order-1 partial state non-zero so far
    alive2 = False         # any order-2 partial state non-zero so far
    alive_vars = set()     # order-1 variables injected so far
    seg_first = 0

    maps_all = []          # lattice mode: gather maps of all shifts, concatenated (F+ | F- | Z per shift)
    lat_i = 0

    def close_segment(shift, n_old, n_new, flags=0, rsv=0):
        nonlocal seg_first
        if lattice:  # the pass covers every slot; order 0 sits at slot kzero (bits 16.. of the flags)
            flags |= SEG_LATTICE | (lat.kzero << 16)
        segs.append([seg_first, len(bld.records) - seg_first, n_old, shift, n_old, n_new, flags, rsv])
        seg_first = len(bld.records)

    def part_flag():
        return F_PARTIALS if (nvar and (alive or alive2)) else 0

    def lattice_slots(zero_k=False):
        """slots of the current lattice that can hold F+ (all of them, or those with zero wavenumber), their
        wavenumbers (rad/m, first three coordinates) and accumulated times (fourth coordinate, x tvalue)"""
        if not lattice:
            raise NotImplementedError("this probe reads configurations other than k = 0: it needs the lattice path")
        if nvar:
            raise NotImplementedError("derivatives of probes over several configurations")
        cs = np.asarray(lat.coords, dtype=float).reshape(len(lat.coords), lat.kdim)
        unit = np.ones(lat.kdim)
        if kscale is not None:  # float lattice: grid units of wavenumbers / of t * tvalue
            unit = kscale[:lat.kdim].copy()
        else:
            unit[:min(lat.kdim, 3)] = (np.broadcast_to(np.asarray(kvalue, dtype=float), (3,)))[:min(lat.kdim, 3)]
            if lat.kdim == 4:
                unit[3] = float(tvalue)
        phys = cs * unit
        kphys = np.zeros((len(cs), 3))
        kphys[:, :min(lat.kdim, 3)] = phys[:, :3]
        tacc = phys[:, 3] if lat.kdim == 4 else np.zeros(len(cs))
        keep = [i for i, c in enumerate(lat.coords) if lat.occ[c][0] and (not zero_k or not any(c[:3]))]
        if not keep:
            keep = [lat.kzero]
        return keep, kphys[keep], tacc[keep]

    system = {}

    def system_now():
        return dict(system)

    for op in seq:
        if isinstance(op, Jacobian) and probes is not None:
            pass  # an in-sequence Jacobian is a probe like any other
        if isinstance(op, S) and lattice:
            maps, kz_before, n_before, coords_after = lat_steps[lat_i]
            lat_i += 1
            off = len(maps_all)
            for mp in maps:
                maps_all.extend(mp)
            assert lat.kzero == kz_before and len(lat.coords) == n_before
            close_segment(2, n_before - 1, len(coords_after) - 1, 0, rsv=off)  # (flags carry the kzero of the pass just closed)
            lat.shift(grid_units(op.k), (max_nstate or op.nmax or None) if kscale is None else None)
            n = len(lat.coords) - 1
        elif isinstance(op, S):
            m = shift_count(op)
            cap = max_nstate or op.nmax or None
            for _ in range(abs(m)):
                n_new = n + 1 if cap is None else min(n + 1, max(cap, 0))
                if n_new < n:
                    raise NotImplementedError(f"{op!r}: nmax={cap} is below the current number of states ({n}); a state matrix "
                                              "that shrinks in mid-sequence is not lowered (use the max_nstate option)")
                close_segment(1 if m > 0 else -1, n, n_new, SEG_MASK_TOP if n_new == n else 0)
                n = n_new
        elif isinstance(op, (X, D, Spoiler)):
            fl = F_BASE | (part_flag() if propagate_nondiff else 0)
            if isinstance(op, Spoiler):
                if lattice:
                    lat.spoil()
                bld.record(OP_SPOIL, fl)
            elif isinstance(op, D) and lattice:
                bld.record(OP_D, fl, [bld.block(diffusion_block(op, lat.coords))])  # the table follows the current lattice
            elif isinstance(op, D):
                key = (id(op), "D")
                if key not in bld.cache:
                    bld.cache[key] = bld.block(diffusion_block(op))
                bld.record(OP_D, fl, [bld.cache[key]])
            else:
                key = (id(op), "X")
                if key not in bld.cache:
                    # mat[..., dst(ax), src(ax+1), ..., 3] -> entry (mT[dst][src], mL[dst][src]) complex
                    mat = np.moveaxis(op.mat, (op.axis, op.axis + 1), (-3, -2))  # [..., dst, src, 3]
                    ent = np.stack([mat[..., 0].real, mat[..., 0].imag], axis=-1).reshape(mat.shape[:-3] + (-1,))
                    enl = np.stack([mat[..., 2].real, mat[..., 2].imag], axis=-1).reshape(mat.shape[:-3] + (-1,))
                    blk = np.concatenate([ent, enl], axis=-1)
                    # re-insert a singleton pool axis so that lead axes line up with the grid
                    blk = np.expand_dims(blk, op.axis) if blk.ndim - 1 >= op.axis else blk
                    bld.cache[key] = bld.block(blk)
                    dens = np.broadcast_to(common.left(sm.density, len(grid)), grid)
                    khi = np.asarray(op.khi, dtype=float)
                    chk = np.einsum("...i,...i->...", khi, np.moveaxis(dens[..., None], op.axis, -1))
                    if not np.allclose(chk, 0):
                        raise RuntimeError("Exchange matrix `khi` does not conserve total magnetization")
                bld.record(OP_X, fl, [bld.cache[key]])
        elif isinstance(op, PD):
            if nvar and alive and not propagate_nondiff:
                raise NotImplementedError("PD after a differentiated operator needs propagate_nondiff=True")
            bld.record(OP_PD, F_BASE, [bld.block(np.atleast_1d(np.asarray(op.pd, dtype=float))[..., None])])
            if op.reset:
                close_segment(0, n, n, SEG_RESET)
                if lattice:
                    lat.reset()
                    n = 0
        elif isinstance(op, Reset):
            if nvar and alive and not propagate_nondiff:
                raise NotImplementedError("RESET after a differentiated operator needs propagate_nondiff=True")
            close_segment(0, n, n, SEG_RESET)
            if lattice:
                lat.reset()
            n = 0
        elif isinstance(op, DiffOperator):
            if lattice and not isinstance(op, (ops_mod.E, ops_mod.P, ops_mod.R, ops_mod.Phi, ops_mod.ScalarOp)):
                lat.mix()
            key = (id(op), "form")
            if key not in bld.cache:
                bld.cache[key] = op._lowered_form()
            form = bld.cache[key]
            inj = [(vindex[var], param, coeff) for var, pc in op.order1.items() if var in vindex
                   for param, coeff in pc.items()] if nvar else []
            if pairs and not inj and not (getattr(op, "order2", None) or {}):
                _emit_or_reuse(bld, op, form, F_BASE | part_flag())  # no derivative of its own: one record for all state sets
            elif pairs:
                # order 2 (diff.py:290-378, every cross term): Op on the order-2 states, their injections from the
                # PRE-operator order-1 / base states, then the order-1 update, then the base state
                #   x_ab <- Op x_ab + sum_p c_ap dOp_p x_b + sum_p c_bp dOp_p x_a + sum_pq c_ap c_bq d2Op_pq x_0 + sum_p c2_ab,p dOp_p x_0
                if alive2:
                    _emit_or_reuse(bld, op, form, F_PARTIALS | F_P2)
                dforms = {}

                def dform(param):
                    if param not in dforms:
                        dforms[param] = op._dform(param)
                    return dforms[param]

                for (a, b) in pairs:
                    dst = pindex[(a, b)]
                    for src, var in ((b, a), (a, b)):  # a == b: both terms, the factor 2 of diff.py:349-362
                        if var in op.order1 and src in alive_vars:
                            for param, coeff in op.order1[var].items():
                                _emit_form(bld, _scaled_form(dform(param), coeff), F_INJECT, aux=dst, aux1=1 + vindex[src])
                                alive2 = True
                    if a in op.order1 and b in op.order1:
                        for p_, ca in op.order1[a].items():
                            for q_, cb in op.order1[b].items():
                                d2 = op._d2form(p_, q_)
                                if d2 is not None:
                                    _emit_form(bld, _scaled_form(_scaled_form(d2, ca), cb), F_INJECT, aux=dst)
                                    alive2 = True
                    for param, c2 in (getattr(op, "order2", None) or {}).get(Pair(a, b), {}).items():
                        _emit_form(bld, _scaled_form(dform(param), c2), F_INJECT, aux=dst)
                        alive2 = True
                if alive:
                    _emit_or_reuse(bld, op, form, F_PARTIALS | F_P1)
                for vi, param, coeff in inj:
                    _emit_form(bld, _scaled_form(dform(param), coeff), F_INJECT, aux=vi)
                    alive = True
                    alive_vars.add(variables[vi])
                _emit_or_reuse(bld, op, form, F_BASE)
            elif not inj:
                _emit_or_reuse(bld, op, form, F_BASE | part_flag())
            else:
                gens = [op._gform(param) for _, param, _ in inj] if pre_inject else [None]
                if all(g is not None for g in gens):
                    # pre-injection: x_v += c Op^-1 dOp x_0, then ONE record applies Op to every state set:
                    # Op (x_v + c Op^-1 dOp x_0) = Op x_v + c dOp x_0, the update of diff.py:264-288
                    for (vi, param, coeff), g in zip(inj, gens):
                        _emit_form(bld, _scaled_form(g, coeff), F_INJECT, aux=vi)
                    alive = True
                    _emit_or_reuse(bld, op, form, F_BASE | F_PARTIALS)
                else:
                    if alive:
                        _emit_or_reuse(bld, op, form, F_PARTIALS)
                    for vi, param, coeff in inj:
                        _emit_form(bld, _scaled_form(op._dform(param), coeff), F_INJECT, aux=vi)
                    alive = True
                    _emit_or_reuse(bld, op, form, F_BASE)
        elif isinstance(op, Probe):
            pass
        elif isinstance(op, ops_mod.System):
            if any(k in ("kvalue", "tvalue") for k in op.properties):
                raise NotImplementedError("System(kvalue=..., tvalue=...) in mid-sequence: pass them as simulate options")
            system.update(op.properties)
        elif isinstance(op, EmptyOperator):
            pass
        else:
            raise NotImplementedError(f"operator {op!r} ({type(op).__name__}) has no device implementation")

        tic = tic + op.duration
        if isinstance(op, Probe):
            rows = []
            for pb in (probes or [op]):
                eff = pb or op
                phase = getattr(op, "phase", None) if isinstance(op, Adc) else None
                custom_post = None
                if not isinstance(op, Adc) and getattr(op, "_post", None):
                    custom_post = op._post
                if isinstance(eff, Hessian):
                    fl = F_PARTIALS | (F_Z0 if eff.probe == "Z0" else 0)
                    blocks = []
                    if phase is not None:
                        ph = np.atleast_1d(np.exp(1j * np.asarray(phase, dtype=float) * common.DEG))
                        blocks = [bld.block(np.stack([ph.real, ph.imag], axis=-1))]
                        fl |= F_SCALE
                    entries = []
                    for a in eff.variables1:
                        row = []
                        for b in eff.variables2:
                            if a == "magnitude" or b == "magnitude":  # first derivatives w.r.t. the other one (diff.py:451-464)
                                v = b if a == "magnitude" else a
                                row.append(vindex.get(v, -1))
                            else:
                                row.append(pindex.get(Pair(a, b), -1))
                        entries.append(row)
                    jrow = -1
                    if nvar:
                        jrow, njac = njac, njac + 1
                        bld.record(OP_ADC, fl, blocks, aux=0, aux1=jrow)
                    rows.append(Row("hess", -1, post=custom_post, jac=(jrow, entries)))
                elif isinstance(eff, Jacobian):
                    attr = eff.probe
                    fl = (F_Z0 if attr == "Z0" else 0)
                    blocks = []
                    if phase is not None:
                        ph = np.exp(1j * np.asarray(phase, dtype=float) * common.DEG)
                        ph = np.atleast_1d(ph)
                        blocks = [bld.block(np.stack([ph.real, ph.imag], axis=-1))]
                        fl |= F_SCALE
                    need_mag = "magnitude" in eff.variables
                    cols = [("mag", None) if v == "magnitude" else ("var", vindex[v]) if v in vindex else ("zero", None)
                            for v in eff.variables]
                    srow = jrow = -1
                    if need_mag:
                        srow, nadc = nadc, nadc + 1
                        fl |= F_BASE
                    if nvar:
                        jrow, njac = njac, njac + 1
                        fl |= F_PARTIALS
                    if fl & (F_BASE | F_PARTIALS):
                        bld.record(OP_ADC, fl, blocks, aux=max(srow, 0), aux1=max(jrow, 0))
                    rows.append(Row("jac", srow, post=custom_post, jac=(jrow, cols)))
                elif getattr(eff, "expr", None) is not None:
                    # eval expression over F0 / Z0: read both on the device, evaluate on the host copy
                    bld.record(OP_ADC, F_BASE, [], aux=nadc)
                    bld.record(OP_ADC, F_BASE | F_Z0, [], aux=nadc + 1)
                    rows.append(Row("expr", nadc, post=(op._post if isinstance(op, Adc) else custom_post), jac=eff))
                    nadc += 2
                elif isinstance(eff, (ops_mod.DFT, ops_mod.Imaging)):
                    # Fourier probes (probe.py:168-219): every configuration that can hold transverse magnetisation is read
                    # into a row of its own (EPGX_FLAG_SLOT) and the host applies the probe's weights to the rows
                    slots, kphys, tacc = lattice_slots()
                    for i, sl in enumerate(slots):
                        bld.record(OP_ADC, F_BASE | F_SLOT, [], aux=nadc + i, aux1=sl)
                    rows.append(Row("fourier", nadc, post=custom_post, jac=(len(slots), kphys, tacc, eff, system_now())))
                    nadc += len(slots)
                elif lattice and lat.kdim == 4 and eff.attr in ("F0", "F0t"):
                    # configurations with an accumulated-time coordinate (shift.py:188-210): F0 is the sum of
                    # exp(-|t|) F over the configurations with zero wavenumber (statematrix.py:149-156)
                    if eff.reduce not in (None, False):
                        raise NotImplementedError("Adc(reduce=...) with accumulated-time coordinates")
                    if eff.attr == "F0t":
                        raise NotImplementedError("Adc('F0t'): one value per configuration; probe F0")
                    slots, _, tacc = lattice_slots(zero_k=True)
                    for i, sl in enumerate(slots):
                        bld.record(OP_ADC, F_BASE | F_SLOT, [], aux=nadc + i, aux1=sl)
                    wts = np.exp(-np.abs(tacc)).astype(complex)
                    scale = None if eff.weights is None else np.asarray(eff.weights, dtype=complex)
                    if phase is not None:
                        scale = np.exp(1j * np.asarray(phase, dtype=float) * common.DEG) * (1 if scale is None else scale)
                    rows.append(Row("lin", nadc, post=custom_post, jac=(len(slots), wts, scale)))
                    nadc += len(slots)
                else:
                    attr = eff.attr
                    if lattice and lat.kdim == 4:
                        raise NotImplementedError(f"Adc('{attr}') with accumulated-time coordinates: probe F0")
                    if attr not in ("F0", "Z0"):
                        raise NotImplementedError(f"Adc('{attr}'): only F0 and Z0 can be probed on the device")
                    red = eff.reduce
                    red = None if (red is None or red is False) else red
                    if red is not None and red is not True and isinstance(red, (int, np.integer)):
                        red = (int(red),)
                    scale = None
                    if eff.weights is not None:
                        scale = np.asarray(eff.weights, dtype=complex)
                    host_post = custom_post
                    if phase is not None:
                        if red is None:
                            ph = np.exp(1j * np.asarray(phase, dtype=float) * common.DEG)
                            if scale is None:
                                scale = ph
                            else:
                                a, b = common.expand_left(np.atleast_1d(scale), np.atleast_1d(ph))
                                scale = a * b
                        else:
                            host_post = op._post
                    fl = F_BASE | (F_Z0 if attr == "Z0" else 0)
                    blocks = []
                    if scale is not None:
                        scale = np.atleast_1d(scale)
                        blocks = [bld.block(np.stack([scale.real, scale.imag], axis=-1))]
                        fl |= F_SCALE
                    bld.record(OP_ADC, fl, blocks, aux=nadc)
                    rows.append(Row("sig", nadc, reduce=red, post=host_post))
                    nadc += 1
            rows_out.append(rows)
            times.append(tic)

    close_segment(0, n, n)

    # ---- prune orders that cannot reach k = 0 before the last read-out
    recs = bld.records
    if fuse and not nvar:
        recs, segs = fuse_records(recs, segs)
    if prune_unobservable and not lattice:
        reach = -1
        for i in range(len(segs) - 1, -1, -1):
            sg = segs[i]
            if sg[6] & SEG_RESET:
                reach = -1
            elif reach >= 0:
                reach += abs(sg[3])
            for r in recs[sg[0]:sg[0] + sg[1]]:
                if r[0] == OP_ADC:
                    reach = max(reach, 0)
                    break
            sg[2] = min(sg[4], reach)
        # orders above the highest observable one need no storage either: clamp the order schedule
        # like a max_nstate truncation would (what is cut off could never come back to k = 0 in time)
        cap_eff = max(max((sg[2] for sg in segs), default=0), init_n, 0)
        if cap_eff < max_order:
            for sg in segs:
                sg[4], sg[5] = min(sg[4], cap_eff), min(sg[5], cap_eff)
            max_order = cap_eff
    recs = records_array(recs)
    sega = np.zeros(len(segs), dtype=SEG_DTYPE)
    if segs:
        a = np.array(segs, dtype=np.int64)
        for j, name in enumerate(("first", "count", "nact", "shift", "n_old", "n_new", "flags", "rsv")):
            sega[name] = a[:, j]
    segs = sega

    low = Lowered()
    low.dtype = dtype
    low.grid, low.atom_shape, low.pool_axis, low.npool = grid, atom_shape, pool_axis, npool
    low.natoms = int(np.prod(atom_shape))
    low.patterns = sorted(bld.patterns, key=bld.patterns.get)
    low.ops, low.segs = recs, segs
    low.coef = np.concatenate(bld.chunks) if bld.chunks else np.zeros(1)
    low.init_ref, low.m0_ref, low.init_n = init_ref, m0_ref, init_n
    low.nadc, low.njac, low.nvar, low.max_order = nadc, njac, nvar, max_order
    low.variables = variables
    low.nvar1, low.pairs = nvar1, pairs
    low.tiles = np.array(tiles, dtype=np.int32).reshape(-1, 3) if tiles else np.zeros((0, 3), dtype=np.int32)
    low.final_n = n  # order count when the tape ends
    low.lattice = lattice
    low.kdim = lat.kdim if lattice else 1
    low.maps = np.asarray(maps_all, dtype=np.int32)
    low.coords = None if lat is None else np.asarray(lat.coords, dtype=np.int64)
    low.rows, low.times = rows_out, times
    low.nprobe = len(probes) if probes else 1
    low.keepalive = seq
    return low


def _emit_or_reuse(bld, op, form, flags):
    """records of the same operator object reuse its coefficient blocks (operators are reusable
    objects in the reference too, docs/basics.md:111)"""
    key = (id(op), "rec")
    rec = bld.cache.get(key)
    if rec is not None:
        bld.records.append((rec[0], (rec[1] & ~(F_BASE | F_PARTIALS | F_INJECT | F_P1 | F_P2)) | flags) + rec[2:])
        return
    _emit_form(bld, form, flags)
    bld.cache[key] = bld.records[-1]
