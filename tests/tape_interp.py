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
# Planned (register-fused, support-restricted) passes: restates GroupGeom / fwd_group / adj_group
# of kernels.cuh on the descriptors the C pass planner emits.
GD_HDR = 2


def plan_passes(n, ops, gmax, init=("zero", 0), init_index=0, swizzled=True):
    """Run the C pass planner (qas_plan_passes_host) on a copy of a compiled tape ->
    (ops with group marks, number of groups, descriptors uint32[ngroups, 2 + n])."""
    import ctypes
    from qas_b200 import _lib
    lib = _lib.load()
    out = np.ascontiguousarray(ops, dtype=np.uint32).copy()
    stride = GD_HDR + n
    gd = np.zeros(max(1, len(out)) * stride, dtype=np.uint32)
    ng = ctypes.c_int(0)
    kind = {"zero": 0, "basis": 1, "plus": 2}[init[0]]
    rc = lib.qas_plan_passes_host(n, out.ctypes.data, len(out), int(gmax), kind, int(init_index), int(bool(swizzled)),
                                  gd.ctypes.data, len(gd), ctypes.byref(ng))
    assert rc == 0
    return out, ng.value, gd[:ng.value * stride].reshape(ng.value, stride).copy()


def _swz(i):
    f = i >> 3
    f ^= f >> 3
    f ^= f >> 6
    f ^= f >> 12
    return i ^ (f & 7)


def group_cosets(n, grp, gd, swizzled=True):
    """coset representatives (in planner order) and slot masks of one group, as the kernel enumerates
    them: P(t) = P0 ^ xor of X_i over the bits of t; returns (Ps, xm, nsup) in INDEX coordinates
    (a swizzled plan stores slot coordinates; the swizzle is a linear involution)"""
    g = len(grp)
    cv = _swz if swizzled else (lambda x: x)
    P0, dc = cv(int(gd[0])), int(gd[1])
    X = [cv(int(x)) for x in gd[GD_HDR:GD_HDR + n - g]]
    xm = [0] * (1 << g)
    for c in range(1, 1 << g):
        q = c.bit_length() - 1
        xm[c] = xm[c & ~(1 << q)] ^ cv(int(grp[q][0]))
    # every representative lies in the common kernel of the r's: slot bits are logical bits
    for x in X + [P0]:
        for o in grp:
            assert _parity(x & int(o[1])) == 0
    Ps = []
    for t in range(1 << (n - g)):
        P = P0
        for i in range(n - g):
            if (t >> i) & 1:
                P ^= X[i]
        Ps.append(P)
    return Ps, xm, 1 << dc


def quarter_warp_conflicts(n, grp, gd, swizzled=True):
    """worst bank-group multiplicity over the first quarter-warp (coset counter 0..7) and all slots"""
    Ps, xm, _ = group_cosets(n, grp, gd, swizzled)
    worst = 1
    for x in xm:
        seen = {}
        for P in Ps[:8]:
            a = (_swz(P ^ x) if swizzled else (P ^ x)) & 7
            seen[a] = seen.get(a, 0) + 1
        worst = max(worst, max(seen.values()))
    return worst


def run_tape_planned(n, ops, gds, init, init_index, pool, c, params, terms, want_grad=True, swizzled=True):
    """Same contract as run_tape, on a tape that went through the pass planner; gds = descriptors."""
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
    assert len(groups) == len(gds)
    state = {"psi": psi}

    def apply_group(gi, adjoint, lam=None, grads=None):
        j, g = groups[gi]
        psi = state["psi"]
        grp = [ops[j + q] for q in range(g)]
        if (int(grp[0][4]) >> 16) & 0xF != 0:
            return False
        rc = int(grp[0][2])
        assert rc == 0 or g == 1
        Ps, xm, nsup = group_cosets(n, grp, gds[gi], swizzled)
        seen = np.zeros(dim, dtype=np.int64)
        order = range(g) if adjoint else range(g - 1, -1, -1)
        Ms = [np.zeros((2, 2), dtype=np.complex128) for _ in range(g)]
        mats = [table_entry(int(o[3]), pool, c, params) for o in grp]
        for t, P in enumerate(Ps):
            slots = [P ^ x for x in xm]
            for s_ in slots:
                seen[s_] += 1
            ins = t < nsup
            if not ins:
                # the planner's claim: nothing lives outside the first 2^dc cosets.  Forward sweep:
                # exactly zero (never written).  Adjoint sweep: un-applying a gate leaves rounding
                # residue where the support shrank; it is never read again.
                # lambda is not propagated outside either: every later (earlier-in-time) gate only
                # reads it inside its own, smaller support (kernels.cuh adj_group).
                tol = 1e-13 if adjoint else 0.0
                assert all(abs(psi[s_]) <= tol for s_ in slots), "amplitude outside the tracked support"
                continue
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
                        if ins:
                            a[cc], a[c1] = ud[0, 0] * a[cc] + ud[0, 1] * a[c1], ud[1, 0] * a[cc] + ud[1, 1] * a[c1]
                            Ms[q][0, 0] += np.conj(l[cc]) * a[cc]
                            Ms[q][0, 1] += np.conj(l[cc]) * a[c1]
                            Ms[q][1, 0] += np.conj(l[c1]) * a[cc]
                            Ms[q][1, 1] += np.conj(l[c1]) * a[c1]
                        l[cc], l[c1] = ud[0, 0] * l[cc] + ud[0, 1] * l[c1], ud[1, 0] * l[cc] + ud[1, 1] * l[c1]
            if ins:
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

    for gi in range(len(groups) - 1, -1, -1):
        if not apply_group(gi, False):
            j, g = groups[gi]
            m, r, rc, tab, meta = (int(x) for x in ops[j])
            u, _ = table_entry(tab, pool, c, params)
            state["psi"] = np.where(par(idx & r) == 1, u[1, 1], u[0, 0]) * state["psi"]
    psi = state["psi"]
    lam = np.zeros(dim, dtype=np.complex128)
    for term in terms:
        if len(term) == 2:                       # (coeff, Pauli word)
            coeff, word = term
            x, z = sim.word_to_masks(word)
            ny = bin(x & z).count("1")
        else:                                    # (x mask, z mask, coeff, number of Y's): compressed registers
            x, z, coeff, ny = term
        jx = idx ^ x
        lam += coeff * (1j ** ny) * np.where(par(jx & z) == 1, -1.0, 1.0) * psi[jx]
    energy = float(np.vdot(psi, lam).real)
    grads = {}
    if want_grad and groups and (int(ops[groups[0][0]][4]) >> 16) & 0xF == 0:
        # the kernel computes lambda only on the support of the final state: poison the rest
        j0, g0 = groups[0]
        Ps, xm, nsup = group_cosets(n, [ops[j0 + q] for q in range(g0)], gds[0], swizzled)
        inside = np.zeros(dim, dtype=bool)
        for t, P in enumerate(Ps):
            if t < nsup:
                for x in xm:
                    inside[P ^ x] = True
        lam = np.where(inside, lam, 1.0e3 * (1.0 + 1.0j))
    if want_grad:
        for gi in range(len(groups)):
            if apply_group(gi, True, lam, grads):
                continue
            j, g = groups[gi]
            psi = state["psi"]
            m, r, rc, tab, meta = (int(x) for x in ops[j])
            u, du = table_entry(tab, pool, c, params)
            depth, npar = meta & 0xFFFF, (meta >> 20) & 3
            M = np.zeros((2, 2), dtype=np.complex128)
            hi = par(idx & r) == 1
            d = np.where(hi, u[1, 1], u[0, 0]).conj()
            psi = d * psi
            M[0, 0] = np.vdot(lam[~hi], psi[~hi])
            M[1, 1] = np.vdot(lam[hi], psi[hi])
            lam *= d
            state["psi"] = psi
            for jj in range(npar):
                grads[(depth, jj)] = 2.0 * float(np.sum(du[jj] * M).real)
    return energy, grads
