"""Second-quantised Hamiltonian -> Pauli sum (Jordan-Wigner), plus a minimal reader for the
OpenFermion MolecularData HDF5 files shipped with the reference
(results_and_plots-legacy/molecule_pyscf_sto-{3g,6g}.hdf5).

Used only by tools/make_hamiltonians.py to generate the data files under qas_b200/data/.
h5py is not installed, so the reader scans for the zlib streams of the float64 datasets and
identifies them by size and by their mathematical properties.
"""
import zlib

import numpy as np


# ----------------------------------------------------------------------------- HDF5 scan
def scan_zlib_float64(path):
    """Return every zlib stream in the file that inflates to a whole number of doubles."""
    raw = open(path, "rb").read()
    out = []
    i = 0
    while i < len(raw) - 2:
        if raw[i] == 0x78 and raw[i + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            try:
                d = zlib.decompressobj()
                data = d.decompress(raw[i:])
                if d.eof and len(data) >= 8 and len(data) % 8 == 0:
                    out.append((i, np.frombuffer(data, dtype="<f8").copy()))
                    i += len(raw[i:]) - len(d.unused_data)
                    continue
            except zlib.error:
                pass
        i += 1
    return out


def read_molecular_integrals(path):
    """-> (one_body [m,m], two_body [m,m,m,m] in OpenFermion order h[p,q,r,s])."""
    streams = scan_zlib_float64(path)
    two = [a for _, a in streams if round(a.size ** 0.25) ** 4 == a.size and a.size >= 16]
    assert two, "no two-body tensor found"
    two_body = max(two, key=lambda a: a.size)
    m = int(round(two_body.size ** 0.25))
    two_body = two_body.reshape(m, m, m, m)
    sq = [a.reshape(m, m) for _, a in streams if a.size == m * m]
    overlap = [a for a in sq if np.allclose(np.diag(a), 1.0) and np.allclose(a, a.T)]
    assert len(overlap) == 1
    s = overlap[0]
    rest = [a for a in sq if a is not s]
    orbitals = [a for a in rest if np.allclose(a.T @ s @ a, np.eye(m), atol=1e-6)]
    assert len(orbitals) == 1
    one = [a for a in rest if a is not orbitals[0] and np.allclose(a, a.T, atol=1e-10)]
    assert len(one) == 1
    return one[0], two_body


def active_space(one_body, two_body, occupied, active):
    """OpenFermion get_active_space_integrals: fold doubly occupied core orbitals."""
    const = 0.0
    for i in occupied:
        const += 2 * one_body[i, i]
        for j in occupied:
            const += 2 * two_body[i, j, j, i] - two_body[i, j, i, j]
    ob = one_body.copy()
    for u in active:
        for v in active:
            for i in occupied:
                ob[u, v] += 2 * two_body[i, u, v, i] - two_body[i, u, i, v]
    ix = np.ix_(active, active)
    return const, ob[ix], two_body[np.ix_(active, active, active, active)]


def spinorb_from_spatial(one_body, two_body):
    """Interleaved spin orbitals (even = alpha, odd = beta), OpenFermion convention."""
    m = one_body.shape[0]
    n = 2 * m
    h1 = np.zeros((n, n))
    h2 = np.zeros((n, n, n, n))
    for p in range(m):
        for q in range(m):
            h1[2 * p, 2 * q] = h1[2 * p + 1, 2 * q + 1] = one_body[p, q]
            for r in range(m):
                for s in range(m):
                    v = two_body[p, q, r, s]
                    h2[2 * p, 2 * q + 1, 2 * r + 1, 2 * s] = v
                    h2[2 * p + 1, 2 * q, 2 * r, 2 * s + 1] = v
                    h2[2 * p, 2 * q, 2 * r, 2 * s] = v
                    h2[2 * p + 1, 2 * q + 1, 2 * r + 1, 2 * s + 1] = v
    return h1, h2


# ----------------------------------------------------------------------------- Pauli algebra
# A Pauli string on n qubits is (x, z) bit masks over qubit index (bit q = qubit q) with the
# operator i^{popcount(x&z)} X^x Z^z, i.e. Y = iXZ.  Products pick up extra phases.
def _mul(a, b):
    """(xa,za,ca) * (xb,zb,cb) in the X^x Z^z normal form (phases tracked in c)."""
    xa, za, ca = a
    xb, zb, cb = b
    # X^xa Z^za X^xb Z^zb = (-1)^{|za & xb|} X^{xa^xb} Z^{za^zb}
    sign = -1.0 if bin(za & xb).count("1") & 1 else 1.0
    return (xa ^ xb, za ^ zb, ca * cb * sign)


def _ladder(p, dagger):
    """JW a_p = Z_{<p} (X_p + iY_p)/2 ; a_p^dag = Z_{<p} (X_p - iY_p)/2, in X^xZ^z form."""
    zlow = (1 << p) - 1
    bit = 1 << p
    # X_p Z_{<p}:   x=bit, z=zlow (X^xZ^z with disjoint support: no phase)
    # Y_p = i X_p Z_p -> i * (x=bit, z=bit|zlow)
    sgn = -1.0 if dagger else 1.0
    return [(bit, zlow, 0.5 + 0j), (bit, zlow | bit, 0.5 * sgn * 1j * 1j)]
    # note: (X + sgn*iY)/2 with Y = iXZ -> (X + sgn*i*i XZ)/2 = (X - sgn XZ)/2


def jordan_wigner(const, h1, h2, tol=1e-12):
    """H = const + sum h1[p,q] a+_p a_q + 1/2 sum h2[p,q,r,s] a+_p a+_q a_r a_s
    -> dict {(x, z): complex coeff of X^x Z^z} (qubit p = spin orbital p)."""
    n = h1.shape[0]
    acc = {(0, 0): complex(const)}
    cr = [_ladder(p, True) for p in range(n)]
    an = [_ladder(p, False) for p in range(n)]

    def add(ops, coeff):
        terms = [(0, 0, coeff)]
        for op in ops:
            terms = [_mul(t, o) for t in terms for o in op]
        for x, z, c in terms:
            acc[(x, z)] = acc.get((x, z), 0.0) + c

    for p in range(n):
        for q in range(n):
            if abs(h1[p, q]) > tol:
                add([cr[p], an[q]], h1[p, q])
    nz = np.argwhere(np.abs(h2) > tol)
    for p, q, r, s in nz:
        add([cr[p], cr[q], an[r], an[s]], 0.5 * h2[p, q, r, s])
    return {k: v for k, v in acc.items() if abs(v) > tol}


def to_words(n, acc, tol=1e-10):
    """{(x,z): coeff of X^xZ^z} -> [(real coeff, 'IXYZ..' word)], wire w = qubit w = string
    position w (wire 0 first / most significant).  Y = iXZ so X^xZ^z = (-i)^{|x&z|} * word."""
    out = []
    for (x, z), c in sorted(acc.items()):
        ny = bin(x & z).count("1")
        c = c * (-1j) ** ny
        assert abs(c.imag) < 1e-9, (x, z, c)
        if abs(c.real) < tol:
            continue
        word = "".join("IXZY"[((x >> q) & 1) | (((z >> q) & 1) << 1)] for q in range(n))
        out.append((float(c.real), word))
    return out
