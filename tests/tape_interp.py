"""Numpy interpreter of the op-tape (include/epgx.h) -- TEST INFRASTRUCTURE ONLY.

Executes a `lowering.Lowered` with the semantics the CUDA kernels implement, vectorised over atoms,
so that the host lowering (patterns, records, segments, order schedule, pruning) can be checked on a
machine without a GPU against the oracle and the golden vectors.  It is never imported by the
product package.
"""

import numpy as np

from epgpy_b200 import lowering as L


def run(low, stream=None):
    """-> signal [nadc, natoms, npool], jac [njac, nvar, natoms, npool] (complex128).
    With `stream` (Plan.stream(): the merged record stream of the register kernels, csrc/epgx_common.cuh) the
    interpreter walks that stream instead of the tape's segments: segment markers, whole-TR records (TR / TRC) and the
    whole-TR derivative groups (TRJ), the latter through the fused five-coefficient algebra of epgx_realjac.cuh
    (trj_assemble) -- a host-side check of the stream builder and of that algebra."""
    coef = low.coef
    ashape, npool = low.atom_shape, low.npool
    natoms = low.natoms
    idx = np.stack(np.unravel_index(np.arange(natoms), ashape), axis=-1)  # [natoms, ndim]
    poff = []
    for strides, ps in low.patterns:
        base = (idx[:, : len(strides)] * np.asarray(strides)[None, :]).sum(-1)
        poff.append(base[:, None] + ps * np.arange(npool)[None, :])  # [natoms, npool]

    def blk(off, pat, n, extra=0):
        """[natoms, npool, n] reals of a block"""
        o = off + poff[pat][..., None] + extra + np.arange(n)
        return coef[o]

    C = low.max_order + 1
    nset = 1 + low.nvar
    P = np.zeros((natoms, npool, nset, C), dtype=complex)
    M = np.zeros_like(P)
    Z = np.zeros_like(P)
    ib = blk(low.init_ref[0], low.init_ref[1], 6 * (low.init_n + 1)).reshape(natoms, npool, low.init_n + 1, 6)
    P[:, :, 0, : low.init_n + 1] = ib[..., 0] + 1j * ib[..., 1]
    M[:, :, 0, : low.init_n + 1] = ib[..., 2] + 1j * ib[..., 3]
    Z[:, :, 0, : low.init_n + 1] = ib[..., 4] + 1j * ib[..., 5]
    m0 = blk(low.m0_ref[0], low.m0_ref[1], 1)[..., 0]  # [natoms, npool]
    sig = np.zeros((low.nadc, natoms, npool), dtype=complex)
    jac = np.zeros((low.njac, low.nvar, natoms, npool), dtype=complex)

    def cplx(b, i):
        return b[..., i] + 1j * b[..., i + 1]

    def expand(records):
        """FUSED + CONT -> the E / T / E records they stand for (independent check of the fusion)"""
        out, it = [], iter(records)
        for rec in it:
            if int(rec["code"]) != L.OP_FUSED:
                out.append(rec)
                continue
            cont = next(it)
            fl = int(rec["flags"])
            if fl & L.F_PRE:
                e = np.zeros((), dtype=L.OP_DTYPE)
                e["code"], e["flags"] = L.OP_E, L.F_BASE | L.F_AFFINE
                e["off"][0:2], e["pat"][0:2] = rec["off"][1:3], rec["pat"][1:3]
                out.append(e)
            t = np.zeros((), dtype=L.OP_DTYPE)
            t["code"], t["flags"] = (L.OP_T_GEN if fl & L.F_GEN else L.OP_T_IM if fl & L.F_IM else L.OP_T_RE), L.F_BASE
            t["off"][0], t["pat"][0] = rec["off"][0], rec["pat"][0]
            out.append(t)
            if fl & L.F_POST:
                e = np.zeros((), dtype=L.OP_DTYPE)
                e["code"], e["flags"] = L.OP_E, L.F_BASE | L.F_AFFINE
                e["off"][0:2], e["pat"][0:2] = cont["off"][0:2], cont["pat"][0:2]
                out.append(e)
        return out

    st = {"m0": m0, "kz": 0}  # kz: slot of the order k = 0 (lattice mode: epgx.h EPGX_SEG_LATTICE; else 0)

    def apply(rec, na):
        if True:
            m0 = st["m0"]
            code, flags, aux, aux1 = int(rec["code"]), int(rec["flags"]), int(rec["aux"]), int(rec["aux1"])
            off, pat = rec["off"], rec["pat"]
            sets = []
            nvar1 = getattr(low, "nvar1", low.nvar)
            if flags & L.F_INJECT:
                sets = ["inject"]
            else:
                if flags & L.F_BASE:
                    sets.append(0)
                if flags & L.F_PARTIALS:  # P1 / P2: order-1 / order-2 partial states only
                    sets += [s for s in range(1, nset) if not ((flags & L.F_P1) and s - 1 >= nvar1)
                             and not ((flags & L.F_P2) and s - 1 < nvar1)]
            if na <= 0 and code != L.OP_PD:
                return
            p, m, z = P[..., :na], M[..., :na], Z[..., :na]  # views [natoms, npool, nset, na]

            def linear(fn, affine):
                """fn(p, m, z) -> (p', m', z') on [natoms, npool, na]; affine: (ap, am, az) at k = 0"""
                for s in sets:
                    src = aux1 if s == "inject" else s  # injection source: 0 = base state, 1 + v = partial state v
                    o = fn(p[:, :, src], m[:, :, src], z[:, :, src])
                    o = [np.array(x) for x in o]
                    if affine is not None and src == 0 and (flags & L.F_AFFINE):
                        for x, a in zip(o, affine):
                            x[..., st["kz"]] += a
                    if s == "inject":
                        p[:, :, aux + 1] += o[0]; m[:, :, aux + 1] += o[1]; z[:, :, aux + 1] += o[2]
                    else:
                        p[:, :, s], m[:, :, s], z[:, :, s] = o

            if code in (L.OP_T_GEN, L.OP_T_RE, L.OP_T_IM):
                if code == L.OP_T_GEN:
                    b = blk(off[0], pat[0], 6)
                    a, w, B, U = b[..., 0], b[..., 1], cplx(b, 2), cplx(b, 4)
                else:
                    b = blk(off[0], pat[0], 4)
                    a, w, B = b[..., 0], b[..., 1], b[..., 2] + 0j
                    U = b[..., 3] + 0j if code == L.OP_T_RE else -1j * b[..., 3]
                a, w, B, U = (x[..., None] for x in (a, w, B, U))
                linear(lambda p_, m_, z_: (a * p_ + B * m_ + U * z_, B.conj() * p_ + a * m_ + U.conj() * z_,
                                           -0.5 * (U.conj() * p_ + U * m_) + w * z_), None)
            elif code == L.OP_E:
                b0 = blk(off[0], pat[0], 2)
                e1, r0 = b0[..., 0, None], b0[..., 1]
                e2 = blk(off[1], pat[1], 1)[..., 0, None]
                f = e2 + 0j
                if flags & L.F_G:
                    b2 = blk(off[2], pat[2], 2)
                    f = e2 * cplx(b2, 0)[..., None]
                linear(lambda p_, m_, z_: (f * p_, f.conj() * m_, e1 * z_), (0, 0, r0 * m0))
            elif code == L.OP_DIAG:
                b = blk(off[0], pat[0], 8)
                aP, aM, aZ, a0 = (cplx(b, i) for i in (0, 2, 4, 6))
                linear(lambda p_, m_, z_: (aP[..., None] * p_, aM[..., None] * m_, aZ[..., None] * z_), (0, 0, a0 * m0))
            elif code == L.OP_MATRIX:
                b = blk(off[0], pat[0], 18)
                mm = (b[..., 0::2] + 1j * b[..., 1::2]).reshape(b.shape[:-1] + (3, 3))[..., None]
                aff = None
                if flags & L.F_AFFINE:
                    b1 = blk(off[1], pat[1], 6)
                    aff = tuple(cplx(b1, i) * m0 for i in (0, 2, 4))
                linear(lambda p_, m_, z_: tuple(mm[..., i, 0, :] * p_ + mm[..., i, 1, :] * m_ + mm[..., i, 2, :] * z_
                                                for i in range(3)), aff)
            elif code == L.OP_D:
                b = blk(off[0], pat[0], 3 * C).reshape(natoms, npool, C, 3)[:, :, :na]
                for s in sets:
                    p[:, :, s] *= b[..., 0]; m[:, :, s] *= b[..., 1]; z[:, :, s] *= b[..., 2]
            elif code == L.OP_X:
                n2 = npool * npool * 2
                b = coef[off[0] + poff[pat[0]][:, 0, None] + np.arange(2 * n2)]  # [natoms, 2*n2]
                mt = (b[:, 0:n2:2] + 1j * b[:, 1:n2:2]).reshape(natoms, npool, npool)
                ml = (b[:, n2::2] + 1j * b[:, n2 + 1::2]).reshape(natoms, npool, npool)
                for s in sets:
                    eq = np.zeros((natoms, npool, na), dtype=complex)
                    if s == 0:
                        eq[..., st["kz"]] = m0
                    p[:, :, s] = np.einsum("aij,ajk->aik", mt, p[:, :, s])
                    m[:, :, s] = np.einsum("aij,ajk->aik", mt.conj(), m[:, :, s])
                    z[:, :, s] = np.einsum("aij,ajk->aik", ml, z[:, :, s] - eq) + eq
            elif code == L.OP_SPOIL:
                for s in sets:
                    p[:, :, s] = 0; m[:, :, s] = 0
            elif code == L.OP_PD:
                st["m0"] = blk(off[0], pat[0], 1)[..., 0]
            elif code == L.OP_ADC:
                f = 1.0
                if flags & L.F_SCALE:
                    f = cplx(blk(off[0], pat[0], 2), 0)
                src = Z if flags & L.F_Z0 else P
                if flags & L.F_BASE:
                    sig[aux] = src[:, :, 0, aux1 if flags & L.F_SLOT else st["kz"]] * f
                if flags & L.F_PARTIALS:
                    for v in range(low.nvar):
                        jac[aux1, v] = src[:, :, 1 + v, st["kz"]] * f
    def close(sh, n_old, n_new, sflags, rsv=0):
        if sflags & L.SEG_RESET:
            P[:] = 0; M[:] = 0; Z[:] = 0
            Z[:, :, 0, 0] = st["m0"]
        elif sh == 2:  # lattice gather: new slot j <- old slot map[j] (-1: empty), per component
            nn = n_new + 1
            for arr, mp in zip((P, M, Z), (low.maps[rsv:rsv + nn], low.maps[rsv + nn:rsv + 2 * nn], low.maps[rsv + 2 * nn:rsv + 3 * nn])):
                old = arr.copy()
                arr[:] = 0
                ok = mp >= 0
                arr[..., np.nonzero(ok)[0]] = old[..., mp[ok]]
        elif sh:
            up, dn = (P, M) if sh > 0 else (M, P)
            new0 = dn[..., 1].conj() if (n_old >= 1 and C > 1) else 0 * dn[..., 0]
            up[..., 1:] = up[..., :-1].copy()
            up[..., 0] = new0
            dn[..., :-1] = dn[..., 1:].copy()
            dn[..., -1] = 0
            dn[..., n_old:] = 0
            up[..., n_new + 1:] = 0

    if stream is None:
        for seg in low.segs:
            st["kz"] = (int(seg["flags"]) >> 16) if int(seg["flags"]) & L.SEG_LATTICE else 0
            for rec in expand(low.ops[seg["first"]: seg["first"] + seg["count"]]):
                apply(rec, int(seg["nact"]) + 1)
            close(int(seg["shift"]), int(seg["n_old"]), int(seg["n_new"]), int(seg["flags"]), int(seg["rsv"]))
        return sig, jac

    # ---- the merged stream (internal codes: csrc/epgx_common.cuh)
    OP_SEG, OP_TR, OP_TRC, OP_TRJ = 64, 65, 66, 67

    def mkrec(code, flags, offs=(), pats=(), aux=0, aux1=0):
        r = np.zeros((), dtype=L.OP_DTYPE)
        r["code"], r["flags"], r["aux"], r["aux1"] = code, flags, aux, aux1
        for i, (o, q) in enumerate(zip(offs, pats)):
            r["off"][i], r["pat"][i] = o, q
        return r

    def close_from(cont):
        w = int(cont["flags"])
        close((w & 3) - 1, int(cont["off"][2]) >> 16, int(cont["off"][2]) & 0xffff, w >> 2)
        return int(cont["aux1"])

    nact, i, n = -1, 0, len(stream)
    while i < n:
        rec = stream[i]
        code, flags = int(rec["code"]), int(rec["flags"]) & 0x0fff  # bits 12..15 of a window's first record: window flags
        if code == OP_SEG:
            sh = int(rec["off"][0])
            close(sh - (1 << 32) if sh >= 1 << 31 else sh, int(rec["off"][1]), int(rec["off"][2]), int(rec["aux1"]))
            nact = int(rec["aux"])
            i += 1
        elif code in (OP_TR, OP_TRC):
            cont = stream[i + 1]
            if code == OP_TRC:
                c2 = stream[i + 2]
                if int(c2["flags"]) & 2:
                    apply(mkrec(L.OP_D, L.F_BASE, [c2["off"][1]], [c2["pat"][1]]), nact + 1)
            f = rec.copy()
            f["code"], f["flags"] = L.OP_FUSED, flags
            for r in expand([f, mkrec(L.OP_CONT, 0, cont["off"][:2], cont["pat"][:2])]):
                apply(r, nact + 1)
            adc = mkrec(L.OP_ADC, L.F_BASE, aux=int(cont["aux"]))
            if code == OP_TRC and int(stream[i + 2]["flags"]) & 1:
                adc = mkrec(L.OP_ADC, L.F_BASE | L.F_SCALE, [stream[i + 2]["off"][0]], [stream[i + 2]["pat"][0]], aux=int(cont["aux"]))
            apply(adc, nact + 1)
            nact = close_from(cont)
            i += 2 if code == OP_TR else 3
        elif code == OP_TRJ:
            g = stream[i:i + 5]
            na = nact + 1
            one = np.ones((natoms, npool))
            tb = blk(g[0]["off"][0], g[0]["pat"][0], 4)
            ta, tw, tbb, tu = (tb[..., k] for k in range(4))
            th = -0.5 * tu
            al1 = al2 = be1 = be2 = one
            ra = rb = 0 * one
            if flags & L.F_PRE:
                b0 = blk(g[0]["off"][1], g[0]["pat"][1], 2)
                al1, ra, al2 = b0[..., 0], b0[..., 1], blk(g[0]["off"][2], g[0]["pat"][2], 1)[..., 0]
            if flags & L.F_POST:
                b0 = blk(g[1]["off"][0], g[1]["pat"][0], 2)
                be1, rb, be2 = b0[..., 0], b0[..., 1], blk(g[1]["off"][1], g[1]["pat"][1], 1)[..., 0]
            F = dict(a=be2 * al2 * ta, b=be2 * al2 * tbb, u=be2 * al1 * tu, h=be1 * al2 * th, w=be1 * al1 * tw,
                     f=be2 * tu * ra, z=be1 * tw * ra + rb)
            x = [P[:, :, 0, :na].copy(), M[:, :, 0, :na].copy(), Z[:, :, 0, :na].copy()]
            m0 = st["m0"]

            def lin(c, p_, m_, z_):
                a, b, u, h, w = (c[k][..., None] for k in "abuhw")
                return a * p_ + b * m_ + u * z_, a * m_ + b * p_ + u * z_, w * z_ + h * (p_ + m_)

            for s in range(nset):
                o = lin(F, P[:, :, s, :na], M[:, :, s, :na], Z[:, :, s, :na])
                if s == 0:
                    J = F
                else:
                    v = s - 1
                    pp = pz = p0 = ga = gw = gb = gu = qq = qz = q0 = 0 * one
                    if int(g[2]["flags"]) >> v & 1:
                        c = blk(g[2]["off"][v], g[2]["pat"][v], 8); pp, pz, p0 = c[..., 0], c[..., 4], c[..., 6]
                    if int(g[3]["flags"]) >> v & 1:
                        c = blk(g[3]["off"][v], g[3]["pat"][v], 4); ga, gw, gb, gu = (c[..., k] for k in range(4))
                    if int(g[4]["flags"]) >> v & 1:
                        c = blk(g[4]["off"][v], g[4]["pat"][v], 8); qq, qz, q0 = c[..., 0], c[..., 4], c[..., 6]
                    gh = -0.5 * gu
                    A = ta * ga + tbb * gb + tu * gh
                    B = ta * gb + tbb * ga + tu * gh
                    U = (ta + tbb) * gu + tu * gw
                    H = th * (ga + gb) + tw * gh
                    W = 2 * th * gu + tw * gw
                    J = dict(a=F["a"] * (pp + qq) + be2 * al2 * A, w=F["w"] * (pz + qz) + be1 * al1 * W,
                             b=F["b"] * (pp + qq) + be2 * al2 * B, u=F["u"] * (pz + qq) + be2 * al1 * U,
                             h=F["h"] * (pp + qz) + be1 * al2 * H, f=F["u"] * p0 + be2 * (U + qq * tu) * ra,
                             z=F["w"] * p0 + be1 * ((W + qz * tw) * ra + q0))
                    o = [a_ + b_ for a_, b_ in zip(o, lin(J, *x))]
                o = [np.array(t) for t in o]
                if na > 0:
                    o[0][..., 0] += J["f"] * m0; o[1][..., 0] += J["f"] * m0; o[2][..., 0] += J["z"] * m0
                P[:, :, s, :na], M[:, :, s, :na], Z[:, :, s, :na] = o
            sig[int(g[1]["aux"])] = P[:, :, 0, 0]
            if flags & L.F_PARTIALS:
                for v in range(low.nvar):
                    jac[int(g[1]["rsv1"]), v] = P[:, :, 1 + v, 0]
            nact = close_from(g[1])
            i += 5
        elif code == L.OP_FUSED:
            for r in expand([rec, stream[i + 1]]):
                apply(r, nact + 1)
            i += 2
        else:
            if code != L.OP_NOP:
                r = rec.copy()
                r["flags"] = flags
                apply(r, nact + 1)
            i += 1
    return sig, jac


def simulate(epg_mod, sequence, **kw):
    """like functions.simulate, but executed by this interpreter (host only)"""
    from epgpy_b200 import functions

    init = kw.pop("init", None)
    probe = kw.pop("probe", None)
    propagate = kw.pop("propagate_nondiff", False)
    prune = kw.pop("prune_unobservable", True)
    asarray = kw.pop("asarray", True)
    extra = {k: kw.pop(k) for k in ("fuse", "pre_inject") if k in kw}
    low = L.lower(sequence, init=init, probe=probe, options=kw, propagate_nondiff=propagate, prune_unobservable=prune,
                  **extra)
    sig, jac = run(low)

    class _T:  # minimal stand-in for a device tensor
        def __init__(self, a):
            self.a = a

        def cpu(self):
            return self

        def numpy(self):
            return self.a

        def __getitem__(self, i):
            return _T(self.a[i])

        def reshape(self, shape):
            return _T(self.a.reshape(shape))

    saved = functions.engine.device_reduce
    functions.engine.device_reduce = lambda t, axis: _T(t.a.sum(axis=axis))
    try:
        values = functions._assemble(low, [(0, 0, low.natoms, _T(sig), _T(jac))], asarray=asarray)
    finally:
        functions.engine.device_reduce = saved
    return values[0] if len(values) == 1 else values
