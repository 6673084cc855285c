"""Minimal restricted Hartree-Fock over STO-6G (s and p shells) in NumPy, used only by
tools/make_hamiltonians.py to rebuild the H2O operator of the reference
(qas/qml_models/qml_models_legacy.py:1245-1256: qml.qchem.molecular_hamiltonian(..., basis="sto-6g",
active_electrons=4, active_orbitals=4)), which upstream comes from PySCF/OpenFermion/PennyLane --
none of them installable here.

Pinning (tests/test_hamiltonian_provider.py):
  * the RHF energy must reproduce `converged SCF energy = -75.1645758157887` and the full-space
    FCI energy `E(FCI) = -75.491788196432` quoted at search-scripts-legacy/search_H2O_near_cx.py:45-49;
  * the same code run on LiH must reproduce the molecular-orbital integrals the reference ships in
    results_and_plots-legacy/molecule_pyscf_sto-6g.hdf5 (up to the sign of each orbital).

Integrals: McMurchie-Davidson (Hermite Gaussians), vectorised over the 6x6(x6x6) primitives of a
contracted pair (quartet).  STO-6G: Hehre, Stewart, Pople, J. Chem. Phys. 51, 2657 (1969) --
expansions of Slater 1s / 2sp functions with unit exponent, scaled by zeta^2.
"""
import itertools

import numpy as np
from scipy.special import gammainc, gamma

BOHR_PER_ANGSTROM = 1.0 / 0.529177210903

# Slater exponent 1 expansions (exponents, contraction coefficients of normalised primitives)
_STO6G_1S = ([2.310303149e+01, 4.235915534e+00, 1.185056519e+00, 4.070988982e-01, 1.580884151e-01,
              6.510953954e-02],
             [9.163596280e-03, 4.936149294e-02, 1.685383049e-01, 3.705627997e-01, 4.164915298e-01,
              1.303340841e-01])
_STO6G_2SP = ([1.030869372e+01, 2.040359519e+00, 6.341422177e-01, 2.439773685e-01, 1.059595374e-01,
               4.856900860e-02],
              [-1.325278809e-02, -4.699171014e-02, -3.378537151e-02, 2.502417861e-01, 5.951172526e-01,
               2.407061763e-01],
              [3.759696623e-03, 3.767936984e-02, 1.738967435e-01, 4.180364347e-01, 4.258595477e-01,
               1.017082955e-01])
_STO3G_1S = ([2.227660584e+00, 4.057711562e-01, 1.098175104e-01],
             [1.543289673e-01, 5.353281423e-01, 4.446345422e-01])
_STO3G_2SP = ([9.942027296e-01, 2.310313333e-01, 7.513856000e-02],
              [-9.996722919e-02, 3.995128261e-01, 7.001154689e-01],
              [1.559162750e-01, 6.076837186e-01, 3.919573931e-01])
_SETS = {"sto-6g": (_STO6G_1S, _STO6G_2SP), "sto-3g": (_STO3G_1S, _STO3G_2SP)}
# standard molecular Slater exponents of the STO-nG sets: (zeta_1s, zeta_2sp)
_ZETA = {"H": (1.24, None), "Li": (2.69, 0.80), "O": (7.66, 2.25)}
_Z = {"H": 1, "Li": 3, "O": 8}


class BasisFunction:
    __slots__ = ("center", "lmn", "alpha", "coef")

    def __init__(self, center, lmn, alpha, coef):
        self.center = np.asarray(center, dtype=float)
        self.lmn = tuple(lmn)
        self.alpha = np.asarray(alpha, dtype=float)
        l = sum(lmn)
        norm = (2 * self.alpha / np.pi) ** 0.75 * (4 * self.alpha) ** (l / 2.0)   # l <= 1
        self.coef = np.asarray(coef, dtype=float) * norm


def sto_ng_basis(symbols, coords_bohr, basis="sto-6g"):
    """Cartesian basis functions in PySCF's order: per atom 1s, 2s, 2px, 2py, 2pz."""
    s1, s2 = _SETS[basis]
    coords = np.asarray(coords_bohr, dtype=float).reshape(-1, 3)
    out = []
    for sym, r in zip(symbols, coords):
        z1, z2 = _ZETA[sym]
        out.append(BasisFunction(r, (0, 0, 0), np.array(s1[0]) * z1 ** 2, s1[1]))
        if z2 is not None:
            a = np.array(s2[0]) * z2 ** 2
            out.append(BasisFunction(r, (0, 0, 0), a, s2[1]))
            for lmn in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
                out.append(BasisFunction(r, lmn, a, s2[2]))
    return out


# ------------------------------------------------------------------ Hermite machinery
def _hermite_E(i, j, a, b, xab):
    """E_t^{ij}, t = 0..i+j, of one Cartesian direction; a, b, xab broadcast arrays."""
    p = a + b
    mu = a * b / p
    xpa = -b / p * xab
    xpb = a / p * xab
    tab = {(0, 0): [np.exp(-mu * xab * xab)]}

    def get(ii, jj, t):
        e = tab[(ii, jj)]
        return e[t] if 0 <= t < len(e) else 0.0

    for ii in range(1, i + 1):
        tab[(ii, 0)] = [get(ii - 1, 0, t - 1) / (2 * p) + xpa * get(ii - 1, 0, t) + (t + 1) * get(ii - 1, 0, t + 1)
                        for t in range(ii + 1)]
    for jj in range(1, j + 1):
        tab[(i, jj)] = [get(i, jj - 1, t - 1) / (2 * p) + xpb * get(i, jj - 1, t) + (t + 1) * get(i, jj - 1, t + 1)
                        for t in range(i + jj + 1)]
    return tab[(i, j)]


def _boys(n, x):
    x = np.asarray(x, dtype=float)
    small = x < 1e-9
    xs = np.where(small, 1.0, x)
    big = 0.5 * gamma(n + 0.5) * gammainc(n + 0.5, xs) / xs ** (n + 0.5)
    return np.where(small, 1.0 / (2 * n + 1) - x / (2 * n + 3), big)


def _hermite_R(tmax, umax, vmax, alpha, rx, ry, rz):
    """R_{tuv} (order-0 Hermite Coulomb integrals) for t<=tmax, u<=umax, v<=vmax."""
    nmax = tmax + umax + vmax
    r2 = rx * rx + ry * ry + rz * rz
    base = [(-2 * alpha) ** n * _boys(n, alpha * r2) for n in range(nmax + 1)]
    memo = {}

    def R(t, u, v, n):
        if t < 0 or u < 0 or v < 0:
            return 0.0
        key = (t, u, v, n)
        if key in memo:
            return memo[key]
        if t == u == v == 0:
            val = base[n]
        elif t > 0:
            val = (t - 1) * R(t - 2, u, v, n + 1) + rx * R(t - 1, u, v, n + 1)
        elif u > 0:
            val = (u - 1) * R(t, u - 2, v, n + 1) + ry * R(t, u - 1, v, n + 1)
        else:
            val = (v - 1) * R(t, u, v - 2, n + 1) + rz * R(t, u, v - 1, n + 1)
        memo[key] = val
        return val

    return {(t, u, v): R(t, u, v, 0) for t in range(tmax + 1) for u in range(umax + 1) for v in range(vmax + 1)}


def _pair(f, g):
    a = f.alpha[:, None]
    b = g.alpha[None, :]
    cc = f.coef[:, None] * g.coef[None, :]
    ab = f.center - g.center
    p = a + b
    P = (a[..., None] * f.center + b[..., None] * g.center) / p[..., None]
    return a, b, cc, ab, p, P


def overlap(f, g, shift=(0, 0, 0)):
    """<f | g with its Cartesian powers raised by `shift`> (contracted, coefficients of g kept)."""
    a, b, cc, ab, p, _ = _pair(f, g)
    s = cc * (np.pi / p) ** 1.5
    for d in range(3):
        j = g.lmn[d] + shift[d]
        if j < 0:
            return 0.0
        s = s * _hermite_E(f.lmn[d], j, a, b, ab[d])[0]
    return float(np.sum(s))


def kinetic(f, g):
    a, b, cc, ab, p, _ = _pair(f, g)

    def s1(d, j):
        if j < 0:
            return 0.0
        return _hermite_E(f.lmn[d], j, a, b, ab[d])[0]

    tot = 0.0
    for d in range(3):
        j = g.lmn[d]
        td = -0.5 * (j * (j - 1) * s1(d, j - 2) - 2 * b * (2 * j + 1) * s1(d, j) + 4 * b * b * s1(d, j + 2))
        term = td
        for o in range(3):
            if o != d:
                term = term * s1(o, g.lmn[o])
        tot = tot + term
    return float(np.sum(cc * (np.pi / p) ** 1.5 * tot))


def nuclear(f, g, centers, charges):
    a, b, cc, ab, p, P = _pair(f, g)
    E = [_hermite_E(f.lmn[d], g.lmn[d], a, b, ab[d]) for d in range(3)]
    tot = 0.0
    for C, Z in zip(centers, charges):
        pc = P - C
        R = _hermite_R(len(E[0]) - 1, len(E[1]) - 1, len(E[2]) - 1, p, pc[..., 0], pc[..., 1], pc[..., 2])
        acc = 0.0
        for t, u, v in R:
            acc = acc + E[0][t] * E[1][u] * E[2][v] * R[(t, u, v)]
        tot = tot - Z * acc
    return float(np.sum(cc * 2 * np.pi / p * tot))


def eri(f, g, h, k):
    """(fg|hk), chemists' notation."""
    a, b, cab, ab, p, P = _pair(f, g)
    c, d, ccd, cd, q, Q = _pair(h, k)
    sh1 = (slice(None), slice(None), None, None)
    sh2 = (None, None, slice(None), slice(None))
    E1 = [[e[sh1] if np.ndim(e) else e for e in _hermite_E(f.lmn[x], g.lmn[x], a, b, ab[x])] for x in range(3)]
    E2 = [[e[sh2] if np.ndim(e) else e for e in _hermite_E(h.lmn[x], k.lmn[x], c, d, cd[x])] for x in range(3)]
    p4, q4 = p[sh1], q[sh2]
    alpha = p4 * q4 / (p4 + q4)
    pq = P[:, :, None, None, :] - Q[None, None, :, :, :]
    n1 = [len(E1[x]) - 1 for x in range(3)]
    n2 = [len(E2[x]) - 1 for x in range(3)]
    R = _hermite_R(n1[0] + n2[0], n1[1] + n2[1], n1[2] + n2[2], alpha, pq[..., 0], pq[..., 1], pq[..., 2])
    tot = 0.0
    for t, u, v in itertools.product(range(n1[0] + 1), range(n1[1] + 1), range(n1[2] + 1)):
        e1 = E1[0][t] * E1[1][u] * E1[2][v]
        for tt, uu, vv in itertools.product(range(n2[0] + 1), range(n2[1] + 1), range(n2[2] + 1)):
            sign = -1.0 if (tt + uu + vv) & 1 else 1.0
            tot = tot + sign * e1 * E2[0][tt] * E2[1][uu] * E2[2][vv] * R[(t + tt, u + uu, v + vv)]
    pref = 2 * np.pi ** 2.5 / (p4 * q4 * np.sqrt(p4 + q4))
    return float(np.sum(cab[sh1] * ccd[sh2] * pref * tot))


# ------------------------------------------------------------------ SCF
def ao_integrals(symbols, coords_bohr, basis="sto-6g"):
    """-> (S, Hcore, (pq|rs), E_nuc); contracted functions renormalised to <f|f> = 1 as PySCF does."""
    coords = np.asarray(coords_bohr, dtype=float).reshape(-1, 3)
    charges = [_Z[s] for s in symbols]
    bas = sto_ng_basis(symbols, coords, basis)
    for f in bas:
        f.coef = f.coef / np.sqrt(overlap(f, f))
    m = len(bas)
    S = np.zeros((m, m))
    T = np.zeros((m, m))
    V = np.zeros((m, m))
    for i in range(m):
        for j in range(i + 1):
            S[i, j] = S[j, i] = overlap(bas[i], bas[j])
            T[i, j] = T[j, i] = kinetic(bas[i], bas[j])
            V[i, j] = V[j, i] = nuclear(bas[i], bas[j], coords, charges)
    g = np.zeros((m, m, m, m))
    for i in range(m):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            for k in range(m):
                for l in range(k + 1):
                    if k * (k + 1) // 2 + l > ij:
                        continue
                    v = eri(bas[i], bas[j], bas[k], bas[l])
                    for a_, b_, c_, d_ in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k),
                                           (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                        g[a_, b_, c_, d_] = v
    e_nuc = 0.0
    for i in range(len(symbols)):
        for j in range(i):
            e_nuc += charges[i] * charges[j] / np.linalg.norm(coords[i] - coords[j])
    return S, T + V, g, e_nuc


def rhf(S, h, g, n_electrons, e_nuc, max_iter=200, tol=1e-12):
    """Closed-shell SCF with DIIS.  -> (E_total, C [ao, mo], orbital energies)."""
    nocc = n_electrons // 2
    s_val, s_vec = np.linalg.eigh(S)
    X = s_vec @ np.diag(s_val ** -0.5) @ s_vec.T

    def diag(F):
        e, c = np.linalg.eigh(X.T @ F @ X)
        return e, X @ c

    e_orb, C = diag(h)
    D = 2 * C[:, :nocc] @ C[:, :nocc].T
    fs, errs = [], []
    e_old = 0.0
    for _ in range(max_iter):
        J = np.einsum("pqrs,rs->pq", g, D)
        K = np.einsum("prqs,rs->pq", g, D)
        F = h + J - 0.5 * K
        e_tot = 0.5 * np.sum(D * (h + F)) + e_nuc
        err = X.T @ (F @ D @ S - S @ D @ F) @ X
        fs.append(F)
        errs.append(err)
        fs, errs = fs[-8:], errs[-8:]
        if len(fs) > 1:
            n = len(fs)
            B = -np.ones((n + 1, n + 1))
            B[n, n] = 0.0
            for i in range(n):
                for j in range(n):
                    B[i, j] = np.sum(errs[i] * errs[j])
            rhs = np.zeros(n + 1)
            rhs[n] = -1.0
            try:
                w = np.linalg.solve(B, rhs)[:n]
                F = sum(wi * fi for wi, fi in zip(w, fs))
            except np.linalg.LinAlgError:
                pass
        e_orb, C = diag(F)
        D = 2 * C[:, :nocc] @ C[:, :nocc].T
        if abs(e_tot - e_old) < tol and np.max(np.abs(err)) < 1e-9:
            break
        e_old = e_tot
    # deterministic orbital phases: largest-magnitude AO coefficient positive
    for j in range(C.shape[1]):
        if C[np.argmax(np.abs(C[:, j])), j] < 0:
            C[:, j] = -C[:, j]
    return e_tot, C, e_orb


def mo_integrals(h, g, C):
    """-> (one_body [p,q], two_body in OpenFermion order two[p,q,r,s] = (ps|qr))."""
    one = C.T @ h @ C
    mo = np.einsum("pqrs,pa,qb,rc,sd->abcd", g, C, C, C, C, optimize=True)
    return one, np.ascontiguousarray(mo.transpose(0, 2, 3, 1))


# ------------------------------------------------------------------ sector diagonalisation
def sector_ground_energy(n_qubits, words, n_alpha, n_beta):
    """Lowest eigenvalue of a Pauli sum [(coeff, 'IXYZ..')] (wire q = spin orbital q, even = alpha)
    restricted to the determinants with the given electron numbers."""
    idx = []
    for i in range(1 << n_qubits):
        occ = [(i >> (n_qubits - 1 - q)) & 1 for q in range(n_qubits)]
        if sum(occ[0::2]) == n_alpha and sum(occ[1::2]) == n_beta:
            idx.append(i)
    idx = np.array(idx, dtype=np.int64)
    pos = -np.ones(1 << n_qubits, dtype=np.int64)
    pos[idx] = np.arange(len(idx))
    H = np.zeros((len(idx), len(idx)), dtype=complex)
    popc = np.array([bin(i).count("1") & 1 for i in range(1 << n_qubits)], dtype=np.int64)
    for c, w in words:
        x = sum(1 << (n_qubits - 1 - q) for q, ch in enumerate(w) if ch in "XY")
        z = sum(1 << (n_qubits - 1 - q) for q, ch in enumerate(w) if ch in "ZY")
        ny = sum(ch == "Y" for ch in w)
        j = idx ^ x
        ok = pos[j] >= 0
        # P|i> = i^{ny} (-1)^{popc(i & z)} |i ^ x>
        amp = c * (1j) ** ny * (1 - 2 * popc[idx & z])
        np.add.at(H, (pos[j[ok]], np.arange(len(idx))[ok]), amp[ok])
    assert np.allclose(H, H.conj().T, atol=1e-10)
    return float(np.linalg.eigvalsh(H)[0])
