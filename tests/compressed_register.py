"""Prototype of the support-compressed register -- TEST INFRASTRUCTURE.

From a basis state |b0> a circuit only populates b0 ^ D, D = span of the xor masks of its CNOT-folded
tape.  Writing the index as i(t) = b0 ^ xor_k t_k B_k for a basis B of D turns the evaluation into that
of a d = dim D qubit register:

  * a 2x2 op (m, r, rc): pair mask m' = coordinates of m; logical bit parity(i & r) = parity(t & r') ^ f
    with r'_k = parity(B_k & r), f = parity(b0 & r) (f = 1 swaps the roles of the two amplitudes); the
    control parity likewise (rc', fc);
  * a parity-diagonal op: the same (r', f);
  * a Pauli term (x, z, c) with x in D: x' = coordinates of x, z'_k = parity(B_k & z), and the constant
    (-1)^parity(b0 & z) goes into the coefficient; terms with x outside D connect the support to
    amplitudes that are exactly zero and drop out of the energy AND of every derivative (lambda is only
    needed on the support, kernels.cuh adj_group).

`compress` does that translation on a tape from the C tape compiler; `run` executes the compressed
problem with the evaluation kernel's algorithm (forward sweep, lambda = H psi, 2x2 cross-matrix adjoint
sweep) on 2^d amplitudes.  tests/test_host.py checks energies and gradients against the oracle."""
import numpy as np

from oracle import sim
from tape_interp import table_entry


def _par(x):
    return bin(int(x)).count("1") & 1


def _basis(vectors):
    """Echelon basis (highest-bit pivots) of the span -> list of (pivot, vector), pivots descending."""
    basis = []
    for x in vectors:
        x = int(x)
        for pv, v in basis:
            if (x >> pv) & 1:
                x ^= v
        if x:
            basis.append((x.bit_length() - 1, x))
            basis.sort(reverse=True)
    return basis


def _coords(basis, x):
    """Coordinates of x in the basis (bit k <-> basis[k]); None if x is outside the span."""
    t = 0
    x = int(x)
    for k, (pv, v) in enumerate(basis):
        if (x >> pv) & 1:
            x ^= v
            t |= 1 << k
    return t if x == 0 else None


def compress(ops, b0, terms):
    """-> (d, ops' [(m', r', f, rc', fc, has_control, tab, meta)], terms' [(x', z', coeff, ny)])."""
    is_u1 = [((int(o[4]) >> 16) & 0xF) == 0 for o in ops]
    raw = _basis([int(o[0]) for o, u in zip(ops, is_u1) if u])
    # express every basis vector as a combination of ... itself: coordinates are taken w.r.t. `raw`
    d = len(raw)
    B = [v for _, v in raw]

    def mask(r):
        return sum(_par(B[k] & r) << k for k in range(d))

    cops = []
    for o, u in zip(ops, is_u1):
        m, r, rc, tab, meta = (int(x) for x in o)
        mp = _coords(raw, m) if u else 0
        assert mp is not None
        cops.append((mp, mask(r), _par(b0 & r), mask(rc), _par(b0 & rc), bool(rc), tab, meta))
    cterms = []
    for coeff, word in terms:
        x, z = sim.word_to_masks(word)
        xp = _coords(raw, x)
        if xp is None:
            continue
        ny = bin(x & z).count("1")
        cterms.append((xp, mask(z), coeff * (-1.0 if _par(b0 & z) else 1.0), ny))
    return d, cops, cterms


def run(d, cops, cterms, pool, c, params):
    """Energy and {(depth, j): dE/dtheta} of the compressed problem."""
    dim = 1 << d
    idx = np.arange(dim)
    par = np.vectorize(_par)
    psi = np.zeros(dim, dtype=np.complex128)
    psi[0] = 1.0

    def pairs(mp, rp, f, rcp, fc, has_c):
        p0 = idx[(par(idx & rp) ^ f) == 0]            # logical bit 0
        if has_c:
            p0 = p0[(par(p0 & rcp) ^ fc) == 1]
        return p0, p0 ^ mp

    for mp, rp, f, rcp, fc, has_c, tab, meta in cops[::-1]:
        u, _ = table_entry(tab, pool, c, params)
        if (meta >> 16) & 0xF == 0:
            p0, p1 = pairs(mp, rp, f, rcp, fc, has_c)
            a0, a1 = psi[p0].copy(), psi[p1].copy()
            psi[p0] = u[0, 0] * a0 + u[0, 1] * a1
            psi[p1] = u[1, 0] * a0 + u[1, 1] * a1
        else:
            psi = np.where((par(idx & rp) ^ f) == 1, u[1, 1], u[0, 0]) * psi
    lam = np.zeros(dim, dtype=np.complex128)
    for xp, zp, coeff, ny in cterms:
        j = idx ^ xp
        lam += coeff * (1j ** ny) * np.where(par(j & zp) == 1, -1.0, 1.0) * psi[j]
    energy = float(np.vdot(psi, lam).real)
    grads = {}
    for mp, rp, f, rcp, fc, has_c, tab, meta in cops:
        u, du = table_entry(tab, pool, c, params)
        depth, npar = meta & 0xFFFF, (meta >> 20) & 3
        M = np.zeros((2, 2), dtype=np.complex128)
        if (meta >> 16) & 0xF == 0:
            p0, p1 = pairs(mp, rp, f, rcp, fc, has_c)
            ud = u.conj().T
            a0, a1 = psi[p0].copy(), psi[p1].copy()
            q0, q1 = ud[0, 0] * a0 + ud[0, 1] * a1, ud[1, 0] * a0 + ud[1, 1] * a1
            psi[p0], psi[p1] = q0, q1
            l0, l1 = lam[p0].copy(), lam[p1].copy()
            M[0, 0], M[0, 1] = np.vdot(l0, q0), np.vdot(l0, q1)
            M[1, 0], M[1, 1] = np.vdot(l1, q0), np.vdot(l1, q1)
            lam[p0], lam[p1] = ud[0, 0] * l0 + ud[0, 1] * l1, ud[1, 0] * l0 + ud[1, 1] * l1
        else:
            hi = (par(idx & rp) ^ f) == 1
            dd = np.where(hi, u[1, 1], u[0, 0]).conj()
            psi = dd * psi
            M[0, 0] = np.vdot(lam[~hi], psi[~hi])
            M[1, 1] = np.vdot(lam[hi], psi[hi])
            lam = dd * lam
        for j in range(npar):
            grads[(depth, j)] = 2.0 * float(np.sum(du[j] * M).real)
    return energy, grads
