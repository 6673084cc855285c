"""NumPy interpreter of the generalised-gate tape (test infrastructure).

Executes, on the CPU, exactly the algorithm the CUDA evaluation kernel runs on the tape the
C tape compiler emits (qas_b200/csrc/tape.h, kernels.cuh): CNOT-folded index frame, pairs
(P, P^m) selected by parity masks, X-mask grouped Hamiltonian, 2x2 cross-matrix adjoint.
It lets the CPU-only test tier check the tape compiler and the kernel's algorithm against
the oracle without a GPU.  It is not part of the product.
"""
import numpy as np

from oracle import gates as G
from oracle import sim

NCONST = 12
_R = 1.0 / np.sqrt(2.0)
_CONST = {
    0: G.matrix("Hadamard"), 1: G.matrix("PauliX"), 2: G.matrix("PauliY"), 3: G.matrix("PauliZ"),
    4: G.matrix("S"), 5: G.matrix("Sdg"), 6: G.matrix("T"), 7: G.matrix("Tdg"), 8: G.matrix("SX"),
    9: G.matrix("S") @ G.matrix("Hadamard"), 10: G.matrix("Hadamard") @ G.matrix("Sdg"),
    11: np.diag([np.exp(1j * np.pi / 8), np.exp(-1j * np.pi / 8)]),
}
_INNER = {"CRX": "RX", "CRY": "RY", "CRZ": "RZ", "CRot": "Rot", "CPhase": "PhaseShift",
          "ControlledPhaseShift": "PhaseShift", "IsingXX": "RZ", "IsingYY": "RZ", "IsingZZ": "RZ"}


def _parity(x):
    return bin(int(x)).count("1") & 1


def table_entry(tab, pool, c, params):
    """(U 2x2, [dU_j]) for a tape table index."""
    if tab < NCONST:
        return _CONST[tab], []
    i, key = divmod(tab - NCONST, c)
    name, _ = sim._entry(pool, key)
    name = _INNER.get(name, name)
    theta = [float(params[i, key, j]) for j in range(G.num_params(name))]
    return G.matrix(name, theta), [sim.dmatrix(name, theta, j) for j in range(len(theta))]


def run_tape(n, ops, init, init_index, pool, c, params, terms, want_grad=True):
    """-> (energy, {(depth, j): dE/dtheta}) following the kernel's algorithm."""
    dim = 1 << n
    kind = init[0]
    if kind == "plus":
        psi = np.full(dim, 2.0 ** (-0.5 * n), dtype=np.complex128)
    else:
        psi = np.zeros(dim, dtype=np.complex128)
        psi[init_index if kind == "basis" else 0] = 1.0
    idx = np.arange(dim)
    par = np.vectorize(_parity)

    def pairs(m, r, rc):
        p0 = idx[par(idx & r) == 0]
        if rc:
            p0 = p0[par(p0 & rc) == 1]
        return p0, p0 ^ m

    for o in ops[::-1]:                      # forward sweep: tape is stored last-gate-first
        m, r, rc, tab, meta = (int(x) for x in o)
        u, _ = table_entry(tab, pool, c, params)
        if (meta >> 16) & 0xF == 0:
            p0, p1 = pairs(m, r, rc)
            a0, a1 = psi[p0].copy(), psi[p1].copy()
            psi[p0] = u[0, 0] * a0 + u[0, 1] * a1
            psi[p1] = u[1, 0] * a0 + u[1, 1] * a1
        else:
            psi = np.where(par(idx & r) == 1, u[1, 1], u[0, 0]) * psi
    # X-mask grouped Hamiltonian
    lam = np.zeros(dim, dtype=np.complex128)
    for coeff, word in terms:
        x, z = sim.word_to_masks(word)
        ny = _parity(0)  # placeholder to keep names local
        ny = bin(x & z).count("1")
        j = idx ^ x
        lam += coeff * (1j ** ny) * np.where(par(j & z) == 1, -1.0, 1.0) * psi[j]
    energy = float(np.vdot(psi, lam).real)
    grads = {}
    if want_grad:
        for o in ops:                        # adjoint sweep: reverse time order
            m, r, rc, tab, meta = (int(x) for x in o)
            u, du = table_entry(tab, pool, c, params)
            depth, npar = meta & 0xFFFF, (meta >> 20) & 3
            M = np.zeros((2, 2), dtype=np.complex128)
            if (meta >> 16) & 0xF == 0:
                p0, p1 = pairs(m, r, rc)
                ud = u.conj().T
                a0, a1 = psi[p0].copy(), psi[p1].copy()
                q0, q1 = ud[0, 0] * a0 + ud[0, 1] * a1, ud[1, 0] * a0 + ud[1, 1] * a1
                psi[p0], psi[p1] = q0, q1
                l0, l1 = lam[p0].copy(), lam[p1].copy()
                M[0, 0], M[0, 1] = np.vdot(l0, q0), np.vdot(l0, q1)
                M[1, 0], M[1, 1] = np.vdot(l1, q0), np.vdot(l1, q1)
                lam[p0], lam[p1] = ud[0, 0] * l0 + ud[0, 1] * l1, ud[1, 0] * l0 + ud[1, 1] * l1
            else:
                hi = par(idx & r) == 1
                d = np.where(hi, u[1, 1], u[0, 0]).conj()
                psi = d * psi
                M[0, 0] = np.vdot(lam[~hi], psi[~hi])
                M[1, 1] = np.vdot(lam[hi], psi[hi])
                lam = d * lam
            for j in range(npar):
                grads[(depth, j)] = 2.0 * float(np.sum(du[j] * M).real)
    return energy, grads


# ---------------------------------------------------------------------------------------------
# Planned (register-fused) passes: restates GroupGeom / fwd_group / adj_group of kernels.cuh.
def plan_passes(ops, gmax):
    """Run the C pass planner (qas_plan_passes_host) on a copy of a compiled tape."""
    import ctypes
    from qas_b200 import _lib
    lib = _lib.load()
    out = np.ascontiguousarray(ops, dtype=np.uint32).copy()
    ng = ctypes.c_int(0)
    rc = lib.qas_plan_passes_host(out.ctypes.data, len(out), int(gmax), ctypes.byref(ng))
    assert rc == 0
    return out, ng.value


def _insert_zero(t, b):
    return ((t >> b) << (b + 1)) | (t & ((1 << b) - 1))


def group_cosets(n, grp):
    """coset representatives P and slot masks xm of one group, exactly as the kernel enumerates them"""
    g = len(grp)
    R = [int(o[1]) for o in grp]
    pv = [r.bit_length() - 1 for r in R]
    assert all(pv[q] > pv[q + 1] for q in range(g - 1)), "pivots must be strictly descending"
    for q in range(g):
        for q2 in range(g):
            if q != q2:
                assert not (R[q2] >> pv[q]) & 1, "pivot column must have a single 1"
    xm = [0] * (1 << g)
    for c in range(1, 1 << g):
        q = c.bit_length() - 1
        xm[c] = xm[c & ~(1 << q)] ^ int(grp[q][0])
    Ps = []
    for t in range(1 << (n - g)):
        ins = t
        for q in range(g - 1, -1, -1):
            ins = _insert_zero(ins, pv[q])
        P = ins
        for q in range(g):
            P |= _parity(ins & R[q]) << pv[q]
        Ps.append(P)
    return Ps, xm


def run_tape_planned(n, ops, init, init_index, pool, c, params, terms, want_grad=True):
    """Same contract as run_tape, on a tape that went through the pass planner."""
    dim = 1 << n
    kind = init[0]
    if kind == "plus":
        psi = np.full(dim, 2.0 ** (-0.5 * n), dtype=np.complex128)
    else:
        psi = np.zeros(dim, dtype=np.complex128)
        psi[init_index if kind == "basis" else 0] = 1.0
    idx = np.arange(dim)
    par = np.vectorize(_parity)
    # group boundaries from the FIRST-op marks, cross-checked with the LAST-op marks
    groups, j = [], 0
    while j < len(ops):
        g = (int(ops[j][4]) >> 24) & 3
        assert g >= 1
        assert (int(ops[j + g - 1][4]) >> 26) & 3 == g
        groups.append((j, g))
        j += g

    def apply_group(j, g, adjoint, lam=None, grads=None):
        grp = [ops[j + q] for q in range(g)]
        k0 = (int(grp[0][4]) >> 16) & 0xF
        if k0 != 0 or int(grp[0][2]) != 0:
            assert g == 1
            return False
        Ps, xm = group_cosets(n, grp)
        seen = np.zeros(dim, dtype=np.int64)
        order = range(g) if adjoint else range(g - 1, -1, -1)
        Ms = [np.zeros((2, 2), dtype=np.complex128) for _ in range(g)]
        mats = [table_entry(int(o[3]), pool, c, params) for o in grp]
        for P in Ps:
            slots = [P ^ x for x in xm]
            for s_ in slots:
                seen[s_] += 1
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
        assert np.all(seen == 1), "cosets of a group must tile the index space exactly once"
        if adjoint:
            for q in range(g):
                meta = int(grp[q][4])
                du = mats[q][1]
                for jj in range((meta >> 20) & 3):
                    grads[(meta & 0xFFFF, jj)] = 2.0 * float(np.sum(du[jj] * Ms[q]).real)
        return True

    for j, g in groups[::-1]:
        if not apply_group(j, g, False):
            m, r, rc, tab, meta = (int(x) for x in ops[j])
            u, _ = table_entry(tab, pool, c, params)
            if (meta >> 16) & 0xF == 0:
                p0 = idx[par(idx & r) == 0]
                if rc:
                    p0 = p0[par(p0 & rc) == 1]
                p1 = p0 ^ m
                a0, a1 = psi[p0].copy(), psi[p1].copy()
                psi[p0] = u[0, 0] * a0 + u[0, 1] * a1
                psi[p1] = u[1, 0] * a0 + u[1, 1] * a1
            else:
                psi = np.where(par(idx & r) == 1, u[1, 1], u[0, 0]) * psi
    lam = np.zeros(dim, dtype=np.complex128)
    for coeff, word in terms:
        x, z = sim.word_to_masks(word)
        ny = bin(x & z).count("1")
        jx = idx ^ x
        lam += coeff * (1j ** ny) * np.where(par(jx & z) == 1, -1.0, 1.0) * psi[jx]
    energy = float(np.vdot(psi, lam).real)
    grads = {}
    if want_grad:
        for j, g in groups:
            if apply_group(j, g, True, lam, grads):
                continue
            m, r, rc, tab, meta = (int(x) for x in ops[j])
            u, du = table_entry(tab, pool, c, params)
            depth, npar = meta & 0xFFFF, (meta >> 20) & 3
            M = np.zeros((2, 2), dtype=np.complex128)
            if (meta >> 16) & 0xF == 0:
                p0 = idx[par(idx & r) == 0]
                if rc:
                    p0 = p0[par(p0 & rc) == 1]
                p1 = p0 ^ m
                ud = u.conj().T
                a0, a1 = psi[p0].copy(), psi[p1].copy()
                q0, q1 = ud[0, 0] * a0 + ud[0, 1] * a1, ud[1, 0] * a0 + ud[1, 1] * a1
                psi[p0], psi[p1] = q0, q1
                l0, l1 = lam[p0].copy(), lam[p1].copy()
                M[0, 0], M[0, 1] = np.vdot(l0, q0), np.vdot(l0, q1)
                M[1, 0], M[1, 1] = np.vdot(l1, q0), np.vdot(l1, q1)
                lam[p0], lam[p1] = ud[0, 0] * l0 + ud[0, 1] * l1, ud[1, 0] * l0 + ud[1, 1] * l1
            else:
                hi = par(idx & r) == 1
                d = np.where(hi, u[1, 1], u[0, 0]).conj()
                psi = d * psi
                M[0, 0] = np.vdot(lam[~hi], psi[~hi])
                M[1, 1] = np.vdot(lam[hi], psi[hi])
                lam = d * lam
            for jj in range(npar):
                grads[(depth, jj)] = 2.0 * float(np.sum(du[jj] * M).real)
    return energy, grads
