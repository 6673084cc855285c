"""NumPy interpreter of support-compressed tapes ("c-tapes") -- TEST INFRASTRUCTURE.

Executes, on the CPU, the algorithm of the compressed evaluation kernel (qas_b200/csrc/kernels_c.cuh,
evalc_kernel) on the records the C c-tape compiler emits (qas_compile_ctape_host, tape.h): a 2^d-amplitude
register in the time-adapted basis of the support span, passes that enumerate the cosets of a group by
inserting zeros at the pivots of its masks inside the prefix 2^{d_j} and project the representative onto
the common kernel of the parity masks, lambda = H psi by sign-table look-ups at the physical index
i(t) = xor_k t_k B_k, and the 2x2 cross-matrix adjoint sweep.  Not part of the product.
"""
import ctypes

import numpy as np

from oracle import sim
from qas_b200 import _lib
from qas_b200.qml_models.qml_gate_ops import pool_table
from tape_interp import _parity, table_entry

CT_FULL, CT_TOO_BIG = -1, -2


def compile_ctape(n, op_pool, k, xmasks, gmax=2, max_d=12, variant=0):
    """-> dict(d, ops uint32[nops,5], ngroups, basis [d], slices [(x', s)]) ; d < 0: not compressed
    (ops is then the grouped classic tape)."""
    lib = _lib.load()
    table = pool_table(op_pool)
    pool_arr = (_lib.PoolEntry * len(table))(*[_lib.PoolEntry(*e) for e in table])
    k = np.ascontiguousarray(np.asarray(k, dtype=np.int32))
    cap = max(1, 10 * len(k))
    ops = np.zeros((cap, 5), dtype=np.uint32)
    xs = np.ascontiguousarray(np.asarray(xmasks, dtype=np.uint32))
    basis = np.zeros(n, dtype=np.uint32)
    lst = np.zeros(max(1, len(xs)), dtype=np.uint32)
    nops, ng, d, nl = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    rc = lib.qas_compile_ctape_host(n, pool_arr, len(table), len(k), k.ctypes.data, gmax, xs.ctypes.data, len(xs), max_d,
                                    ops.ctypes.data, cap, ctypes.byref(nops), ctypes.byref(ng), ctypes.byref(d),
                                    basis.ctypes.data, lst.ctypes.data, ctypes.byref(nl), variant)
    assert rc == 0, lib.qas_last_error(None).decode()
    return {"d": d.value, "ops": ops[:nops.value].copy(), "ngroups": ng.value,
            "basis": [int(b) for b in basis[:max(d.value, 0)]],
            "slices": [(int(v) & 0xFFFF, int(v) >> 16) for v in lst[:nl.value]]}


def sign_tables(n, terms):
    """X-mask slices of a Pauli sum as the library builds them: -> (xmasks [S], imag [S], V [S, 2^n]) with
    (H psi)[i] = sum_s (i if imag_s) V_s[i] psi[i ^ x_s]."""
    dim = 1 << n
    idx = np.arange(dim)
    par = np.vectorize(_parity)
    slices = {}
    for coeff, word in terms:
        x, z = sim.word_to_masks(word)
        ny = bin(x & z).count("1")
        sgn = -1.0 if (ny & 2) else 1.0
        slices.setdefault((x, ny & 1), []).append((z, sgn * coeff))
    keys = sorted(slices, key=lambda kv: (kv[1], kv[0]))
    V = np.zeros((len(keys), dim))
    for s, (x, im) in enumerate(keys):
        for z, cf in slices[(x, im)]:
            V[s] += cf * np.where(par((idx ^ x) & z) == 1, -1.0, 1.0)
    return [x for x, _ in keys], [im for _, im in keys], V


def _insert_zero(t, b):
    return ((t >> b) << (b + 1)) | (t & ((1 << b) - 1))


def group_cosets(grp, dj):
    """Representatives of the cosets of span(m_a) inside the prefix 2^dj, as the kernel enumerates them."""
    g = len(grp)
    ms = [int(o[0]) for o in grp]
    rs = [int(o[1]) for o in grp]
    # pivots: highest set bit of the masks after echelonising them against each other
    piv, ech = [], []
    for m in ms:
        x = m
        for q, e in zip(piv, ech):
            if (x >> q) & 1:
                x ^= e
        assert x, "masks of a group must be independent"
        piv.append(x.bit_length() - 1)
        ech.append(x)
    assert all(q < dj for q in piv)
    xm = [0] * (1 << g)
    for c in range(1, 1 << g):
        q = c.bit_length() - 1
        xm[c] = xm[c & ~(1 << q)] ^ ms[q]
    Ps = []
    for c in range(1 << (dj - g)):
        t = c
        for q in sorted(piv):
            t = _insert_zero(t, q)
        P = t
        for m, r in zip(ms, rs):
            if _parity(t & r):
                P ^= m
        for r in rs:
            assert _parity(P & r) == 0       # slot bits are logical bits
        Ps.append(P)
    return Ps, xm


def run_ctape(ct, n, pool, c, params, xmasks, imag, V, want_grad=True, poison=True, adjoint_gates=2):
    """-> (energy, {(depth, j): dE/dtheta}) of a compressed record, following evalc_kernel.
    adjoint_gates = 1: the adjoint sweep un-applies the gates of a two-gate group in two one-gate passes, tape order, the
    second on its own support d_j (evalc_kernel<..., AG = 1>); the forward sweep keeps the groups."""
    d, ops = ct["d"], ct["ops"]
    assert d >= 0
    dim = 1 << d
    psi = np.zeros(dim, dtype=np.complex128)
    psi[0] = 1.0
    groups, j = [], 0
    while j < len(ops):
        g = (int(ops[j][4]) >> 24) & 3
        assert g >= 1 and (int(ops[j + g - 1][4]) >> 26) & 3 == g
        groups.append((j, g))
        j += g
    assert len(groups) == ct["ngroups"]
    par = np.vectorize(_parity)

    def apply_group(gi, adjoint, lam=None, grads=None, sub=None):
        nonlocal psi
        j, g = groups[gi] if sub is None else sub
        grp = [ops[j + q] for q in range(g)]
        dj = int(grp[0][4]) >> 28                  # support dimension after the group
        sup = 1 << dj
        if poison:
            assert np.all(psi[sup:] == 0) or adjoint, "amplitude outside the tracked support"
        mats = [table_entry(int(o[3]), pool, c, params) for o in grp]
        if (int(grp[0][4]) >> 16) & 0xF != 0:      # parity-diagonal op on the prefix
            assert g == 1
            r, meta = int(grp[0][1]), int(grp[0][4])
            u, du = mats[0]
            idx = np.arange(sup)
            hi = par(idx & r) == 1
            if not adjoint:
                psi[:sup] = np.where(hi, u[1, 1], u[0, 0]) * psi[:sup]
                return
            dd = np.where(hi, u[1, 1], u[0, 0]).conj()
            psi[:sup] = dd * psi[:sup]
            M = np.zeros((2, 2), dtype=np.complex128)
            M[0, 0] = np.vdot(lam[:sup][~hi], psi[:sup][~hi])
            M[1, 1] = np.vdot(lam[:sup][hi], psi[:sup][hi])
            lam[:sup] = dd * lam[:sup]
            for jj in range((meta >> 20) & 3):
                grads[(meta & 0xFFFF, jj)] = 2.0 * float(np.sum(du[jj] * M).real)
            return
        rc = int(grp[0][2])
        assert rc == 0 or g == 1
        Ps, xm = group_cosets(grp, dj)
        seen = np.zeros(sup, dtype=np.int64)
        order = range(g) if adjoint else range(g - 1, -1, -1)
        Ms = [np.zeros((2, 2), dtype=np.complex128) for _ in range(g)]
        for P in Ps:
            slots = [P ^ x for x in xm]
            for s_ in slots:
                seen[s_] += 1
            if rc and not _parity(P & rc):
                continue
            a = [psi[s_] for s_ in slots]
            l = [lam[s_] for s_ in slots] if adjoint else None
            for q in order:
                u = mats[q][0]
                ud = u.conj().T
                for cc in range(1 << g):
                    if (cc >> q) & 1:
                        continue
                    c1 = cc | (1 << q)
                    if not adjoint:
                        a[cc], a[c1] = u[0, 0] * a[cc] + u[0, 1] * a[c1], u[1, 0] * a[cc] + u[1, 1] * a[c1]
                    else:
                        a[cc], a[c1] = ud[0, 0] * a[cc] + ud[0, 1] * a[c1], ud[1, 0] * a[cc] + ud[1, 1] * a[c1]
                        Ms[q][0, 0] += np.conj(l[cc]) * a[cc]
                        Ms[q][0, 1] += np.conj(l[cc]) * a[c1]
                        Ms[q][1, 0] += np.conj(l[c1]) * a[cc]
                        Ms[q][1, 1] += np.conj(l[c1]) * a[c1]
                        l[cc], l[c1] = ud[0, 0] * l[cc] + ud[0, 1] * l[c1], ud[1, 0] * l[cc] + ud[1, 1] * l[c1]
            for s_, v in zip(slots, a):
                psi[s_] = v
            if adjoint:
                for s_, v in zip(slots, l):
                    lam[s_] = v
        assert np.all(seen == 1), "cosets of a group must tile the support prefix exactly once"
        if adjoint:
            for q in range(g):
                meta = int(grp[q][4])
                du = mats[q][1]
                for jj in range((meta >> 20) & 3):
                    grads[(meta & 0xFFFF, jj)] = 2.0 * float(np.sum(du[jj] * Ms[q]).real)

    for gi in range(len(groups) - 1, -1, -1):
        apply_group(gi, False)
    # lambda = H psi on the support: sign-table look-ups at the physical index
    phys = np.zeros(dim, dtype=np.int64)
    for t in range(dim):
        for kk, b in enumerate(ct["basis"]):
            if (t >> kk) & 1:
                phys[t] ^= b
    lam = np.zeros(dim, dtype=np.complex128)
    tt = np.arange(dim)
    for xp, s in ct["slices"]:
        lam += (1j if imag[s] else 1.0) * V[s][phys] * psi[tt ^ xp]
        assert all((int(phys[t]) ^ int(xmasks[s])) == int(phys[t ^ xp]) for t in range(dim))
    energy = float(np.vdot(psi, lam).real)
    grads = {}
    if want_grad:
        for gi in range(len(groups)):
            j, g = groups[gi]
            if adjoint_gates == 1 and g == 2 and (int(ops[j][4]) >> 16) & 0xF == 0:
                apply_group(gi, True, lam, grads, sub=(j, 1))
                apply_group(gi, True, lam, grads, sub=(j + 1, 1))
            else:
                apply_group(gi, True, lam, grads)
    return energy, grads
