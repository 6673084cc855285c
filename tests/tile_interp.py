"""NumPy interpreter of the TILED evaluation (test infrastructure, CPU tier).

Restates, on the CPU, what eval_tiled_kernel (qas_b200/csrc/kernels_t.cuh) does with a grouped tape: tile passes
planned by the C planner (qas_plan_tile_pass_host -- the code the kernel runs before every pass), ops translated to
pass coordinates (m' = A^-1 m by reduction against the pivot table, r' = A^T r), every pass executed tile by tile
with the kernel's coset enumeration (an echelon basis of the common kernel of the group's parity masks from the C
routine qas_tile_group_basis, tile-number parities as role flips / constant controls), slab written back through the frame, cross
matrices accumulated over the tiles.  It is not part of the product.
"""
import ctypes

import numpy as np

from tape_interp import _parity, table_entry

MAXP = 32


def plan_tile_pass(n, ops, start, direction, tb, clow, max_ops=MAXP):
    from qas_b200 import _lib
    lib = _lib.load()
    ops = np.ascontiguousarray(ops, dtype=np.uint32)
    col = np.zeros(16, dtype=np.uint32)
    piv = np.zeros(16, dtype=np.int32)
    cnt = ctypes.c_int(0)
    rc = lib.qas_plan_tile_pass_host(n, ops.ctypes.data, len(ops), int(start), int(direction), tb, clow, max_ops,
                                     col.ctypes.data, piv.ctypes.data, ctypes.byref(cnt))
    assert rc == 0
    return cnt.value, [int(x) for x in col[:n]], [int(x) for x in piv[:n]]


def _translate(o, n, tb, clow, col, piv):
    m, r, rc = int(o[0]), int(o[1]), int(o[2])
    low = (1 << clow) - 1
    x, mp = m & ~low, m & low
    for p in range(n - 1, clow - 1, -1):
        if (x >> p) & 1:
            ci = piv[p]
            assert 0 <= ci < tb, "mask of a pass op outside the tile span"
            x ^= col[ci]
            mp ^= 1 << ci
    assert x == 0
    rp = sum(_parity(col[i] & r) << i for i in range(n))
    rcp = sum(_parity(col[i] & rc) << i for i in range(n))
    return mp, rp, rcp


def _phys(n, tb, col, tau):
    """physical index of every tile coordinate t of tile tau"""
    t = np.arange(1 << tb)
    out = np.zeros(1 << tb, dtype=np.int64)
    for i in range(tb):
        out ^= np.where((t >> i) & 1, col[i], 0)
    for i in range(n - tb):
        if (tau >> i) & 1:
            out ^= col[tb + i]
    return out


def _insert_zero(t, b):
    return ((t >> b) << (b + 1)) | (t & ((1 << b) - 1))


_par = np.vectorize(_parity)


def group_basis(g, tb, ms, rls):
    """the C routine the kernel runs once per group and pass (qas_tile_group_basis)"""
    from qas_b200 import _lib
    lib = _lib.load()
    m = np.array(ms, dtype=np.uint32)
    r = np.array(rls, dtype=np.uint32)
    kb = np.zeros(16, dtype=np.uint32)
    dim = ctypes.c_int(0)
    rc = lib.qas_tile_group_basis_host(g, m.ctypes.data, r.ctypes.data, tb, kb.ctypes.data, ctypes.byref(dim))
    assert rc == 0 and dim.value == tb - g
    return [int(x) for x in kb[:dim.value]]


def _cosets(g, tb, ms, rls, flips):
    """representatives Pp (logical bits of the group's gates are 0) in the kernel's enumeration order -- coset c =
    role-flip offset xor the basis vectors of K selected by the bits of c -- and the slot masks xm"""
    kb = group_basis(g, tb, ms, rls)
    for v in kb:
        assert all(_parity(v & rls[q]) == 0 for q in range(g))
    off = 0
    for q in range(g):
        if flips[q]:
            off ^= ms[q]
    c = np.arange(1 << (tb - g))
    pp = np.full(1 << (tb - g), off, dtype=np.int64)
    for i, v in enumerate(kb):
        pp ^= np.where((c >> i) & 1, v, 0)
    xm = [0, ms[0]] if g == 1 else [0, ms[0], ms[1], ms[0] ^ ms[1]]
    return pp, xm


def run_tiled(n, ops, init, init_index, pool, c, params, terms, tb, clow, want_grad=True):
    """ops: grouped tape (meta bits 24-27 set by the pass planner).  -> (energy, grads, passes_fwd, passes_adj)"""
    from oracle import sim
    dim, lowtb = 1 << n, (1 << tb) - 1
    slab = np.zeros(dim, dtype=np.complex128)
    lslab = np.zeros(dim, dtype=np.complex128)
    nops = len(ops)
    ntiles = 1 << (n - tb)

    def pass_ops(j0, cnt, col, piv):
        return [(_translate(ops[j0 + i], n, tb, clow, col, piv), int(ops[j0 + i][3]), int(ops[j0 + i][4])) for i in range(cnt)]

    # ---- forward sweep
    e, first, npass_f = nops - 1, True, 0
    while first or e >= 0:
        cnt, col, piv = plan_tile_pass(n, ops, e, -1, tb, clow)
        assert cnt >= 1 or nops == 0
        for q in range(clow):
            assert col[q] == 1 << q
        assert all(cv & ((1 << clow) - 1) == 0 for cv in col[clow:])
        j0 = e - cnt + 1
        tops = pass_ops(j0, cnt, col, piv)
        seen = np.zeros(dim, dtype=bool)
        for tau in range(ntiles):
            ga = _phys(n, tb, col, tau)
            assert not seen[ga].any()
            seen[ga] = True
            if first:
                if init[0] == "plus":
                    tile = np.full(1 << tb, 2.0 ** (-0.5 * n), dtype=np.complex128)
                else:
                    tile = (ga == (init_index if init[0] == "basis" else 0)).astype(np.complex128)
            else:
                tile = slab[ga].copy()
            i = cnt - 1
            while i >= 0:
                g = max(1, (tops[i][2] >> 26) & 3)
                f = i - g + 1
                kind = (tops[f][2] >> 16) & 0xF
                if kind == 0:
                    ms = [tops[f + q][0][0] for q in range(g)]
                    rls = [tops[f + q][0][1] & lowtb for q in range(g)]
                    flips = [bool(_parity(tau & (tops[f + q][0][1] >> tb))) for q in range(g)]
                    pp, xm = _cosets(g, tb, ms, rls, flips)
                    if g == 1 and tops[f][0][2]:
                        rc = tops[f][0][2]
                        fire = (_par(pp & (rc & lowtb)) == 1) != bool(_parity(tau & (rc >> tb)))
                        pp = pp[fire]
                    a = [tile[pp ^ x].copy() for x in xm]
                    for q in range(g - 1, -1, -1):
                        u, _ = table_entry(tops[f + q][1], pool, c, params)
                        for s in range(1 << g):
                            if not (s >> q) & 1:
                                s1 = s | (1 << q)
                                a[s], a[s1] = u[0, 0] * a[s] + u[0, 1] * a[s1], u[1, 0] * a[s] + u[1, 1] * a[s1]
                    for s, x in enumerate(xm):
                        tile[pp ^ x] = a[s]
                else:
                    u, _ = table_entry(tops[f][1], pool, c, params)
                    t = np.arange(1 << tb)
                    hi = (_par(t & (tops[f][0][1] & lowtb)) == 1) != bool(_parity(tau & (tops[f][0][1] >> tb)))
                    tile = np.where(hi, u[1, 1], u[0, 0]) * tile
                i = f - 1
            slab[ga] = tile
        assert seen.all()
        e -= cnt
        first = False
        npass_f += 1
    # ---- H application (dense, physical coordinates)
    idx = np.arange(dim)
    for coeff, word in terms:
        x, z = sim.word_to_masks(word)
        ny = bin(x & z).count("1")
        j = idx ^ x
        lslab += coeff * (1j ** ny) * np.where(_par(j & z) == 1, -1.0, 1.0) * slab[j]
    energy = float(np.vdot(slab, lslab).real)
    grads, npass_a = {}, 0
    if not want_grad:
        return energy, grads, npass_f, npass_a
    # ---- adjoint sweep
    j = 0
    while j < nops:
        cnt, col, piv = plan_tile_pass(n, ops, j, +1, tb, clow)
        assert cnt >= 1
        tops = pass_ops(j, cnt, col, piv)
        Ms = [np.zeros((2, 2), dtype=np.complex128) for _ in range(cnt)]
        for tau in range(ntiles):
            ga = _phys(n, tb, col, tau)
            tile, ltile = slab[ga].copy(), lslab[ga].copy()
            f = 0
            while f < cnt:
                g = max(1, (tops[f][2] >> 24) & 3)
                kind = (tops[f][2] >> 16) & 0xF
                if kind == 0:
                    ms = [tops[f + q][0][0] for q in range(g)]
                    rls = [tops[f + q][0][1] & lowtb for q in range(g)]
                    flips = [bool(_parity(tau & (tops[f + q][0][1] >> tb))) for q in range(g)]
                    pp, xm = _cosets(g, tb, ms, rls, flips)
                    if g == 1 and tops[f][0][2]:
                        rc = tops[f][0][2]
                        fire = (_par(pp & (rc & lowtb)) == 1) != bool(_parity(tau & (rc >> tb)))
                        pp = pp[fire]
                    a = [tile[pp ^ x].copy() for x in xm]
                    l = [ltile[pp ^ x].copy() for x in xm]
                    for q in range(g):
                        u, _ = table_entry(tops[f + q][1], pool, c, params)
                        ud = u.conj().T
                        for s in range(1 << g):
                            if not (s >> q) & 1:
                                s1 = s | (1 << q)
                                a[s], a[s1] = ud[0, 0] * a[s] + ud[0, 1] * a[s1], ud[1, 0] * a[s] + ud[1, 1] * a[s1]
                                Ms[f + q] += np.array([[np.vdot(l[s], a[s]), np.vdot(l[s], a[s1])],
                                                       [np.vdot(l[s1], a[s]), np.vdot(l[s1], a[s1])]])
                                l[s], l[s1] = ud[0, 0] * l[s] + ud[0, 1] * l[s1], ud[1, 0] * l[s] + ud[1, 1] * l[s1]
                    for s, x in enumerate(xm):
                        tile[pp ^ x] = a[s]
                        ltile[pp ^ x] = l[s]
                else:
                    u, _ = table_entry(tops[f][1], pool, c, params)
                    t = np.arange(1 << tb)
                    hi = (_par(t & (tops[f][0][1] & lowtb)) == 1) != bool(_parity(tau & (tops[f][0][1] >> tb)))
                    d = np.where(hi, u[1, 1], u[0, 0]).conj()
                    tile = d * tile
                    Ms[f][0, 0] += np.vdot(ltile[~hi], tile[~hi])
                    Ms[f][1, 1] += np.vdot(ltile[hi], tile[hi])
                    ltile = d * ltile
                f += g
            slab[ga], lslab[ga] = tile, ltile
        for i in range(cnt):
            meta = tops[i][2]
            depth, npar = meta & 0xFFFF, (meta >> 20) & 3
            if npar:
                _, du = table_entry(tops[i][1], pool, c, params)
                for jj in range(npar):
                    grads[(depth, jj)] = 2.0 * float(np.sum(du[jj] * Ms[i]).real)
        j += cnt
        npass_a += 1
    return energy, grads, npass_f, npass_a
