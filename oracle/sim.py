"""NumPy statevector oracle for the QAS hot path (test infrastructure only).

Follows, function by function, the reference code it restates:

  extract_param_indices  <- qas/qml_models/utils.py:11-25 (dup qml_models_legacy.py:201-216)
  lower                  <- backboneCirc, qas/qml_models/hamiltonian_model.py:22-41
                            (PlaceHolder skipped, parameters consumed in order)
  to_list                <- toList, hamiltonian_model.py:93-116
  energy                 <- costFunc / getLoss, hamiltonian_model.py:43-63 and
                            qml_models_legacy.py:1295-1315 (HF basis-state start),
                            :2377-2393 (QAOA: Hadamards first, -sum 1/2 w (1 - <ZZ>))
  grad_*                 <- getGradient, hamiltonian_model.py:75-91 (dense zeros(p,c,l) with the
                            touched entries filled; zeros when nothing is trainable)
  batch_mean_gradient    <- search(), qas/mcts/mcts.py:340-347

The state is a ``[2]*n`` complex128 array updated gate by gate with ``tensordot`` and the
Hamiltonian expectation is accumulated term by term -- the same call pattern as
PennyLane 0.29 ``default.qubit`` (third party, not in the reference tree), so timing this
module is a conservative stand-in for the reference's CPU path.
Wire 0 is the most significant bit of the flat index.
"""
import numpy as np

from . import gates as G

_PAULI = {
    "I": np.eye(2, dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


# ----------------------------------------------------------------------------- lowering
def _entry(pool, key):
    d = pool[key] if key in pool else pool[str(key)]
    assert len(d.keys()) == 1
    name = list(d.keys())[0]
    return name, list(d[name])


def extract_param_indices(k, pool):
    """[(depth i, key k_i, slot j)] for every parameter of every parameterised gate."""
    assert min(k) >= 0
    assert max(k) < len(pool)
    out = []
    for i, key in enumerate(k):
        name, _ = _entry(pool, key)
        for j in range(G.num_params(name)):
            out.append((i, int(key), j))
    return out


def lower(k, pool, params):
    """Ordered gate list [(name, wires, [theta...] or None)]; PlaceHolders are skipped."""
    params = np.asarray(params)
    gates = []
    for i, key in enumerate(k):
        name, wires = _entry(pool, key)
        if name == "PlaceHolder":
            continue
        npar = G.num_params(name)
        theta = [float(params[i, key, j]) for j in range(npar)] if npar > 0 else None
        gates.append((name, wires, theta))
    return gates


def to_list(k, pool, params):
    return lower(k, pool, params)


# ----------------------------------------------------------------------------- simulation
def init_state(n, init=("zero", 0)):
    """init = ("zero", _) | ("basis", flat index) | ("plus", _)   (complex128, shape [2]*n)."""
    kind, arg = init
    if kind == "plus":
        st = np.full(2 ** n, 1.0 / np.sqrt(2.0 ** n), dtype=np.complex128)
    else:
        st = np.zeros(2 ** n, dtype=np.complex128)
        st[arg if kind == "basis" else 0] = 1.0
    return st.reshape([2] * n)


def apply_matrix(state, mat, wires):
    """state <- (mat on `wires`) state, via tensordot + moveaxis."""
    n = state.ndim
    kq = len(wires)
    m = np.asarray(mat, dtype=np.complex128).reshape([2] * (2 * kq))
    out = np.tensordot(m, state, axes=(list(range(kq, 2 * kq)), list(wires)))
    return np.moveaxis(out, list(range(kq)), list(wires))


def run(n, gate_list, init=("zero", 0)):
    st = init_state(n, init)
    for name, wires, theta in gate_list:
        st = apply_matrix(st, G.matrix(name, theta or ()), wires)
    return st


def apply_pauli_word(state, word):
    """word: string over I/X/Y/Z of length n (wire 0 first)."""
    out = state
    for w, ch in enumerate(word):
        if ch != "I":
            out = apply_matrix(out, _PAULI[ch], [w])
    return out


def apply_hamiltonian(state, terms):
    """terms: [(coeff, word)] -> H|state> term by term."""
    acc = np.zeros_like(state)
    for coeff, word in terms:
        acc = acc + coeff * apply_pauli_word(state, word)
    return acc


def expval(state, terms):
    """sum_t c_t <state|P_t|state>, accumulated term by term (real part)."""
    e = 0.0
    for coeff, word in terms:
        e += coeff * np.vdot(state, apply_pauli_word(state, word)).real
    return float(e)


def energy(n, k, pool, params, terms, init=("zero", 0)):
    return expval(run(n, lower(k, pool, params), init), terms)


# ----------------------------------------------------------------------------- derivatives
def dmatrix(name, theta, j):
    """Analytic d(gate matrix)/d(theta_j)."""
    Z2, Y2, X2 = -0.5j * _PAULI["Z"], -0.5j * _PAULI["Y"], -0.5j * _PAULI["X"]
    if name == "Rot":
        phi, th, om = theta
        if j == 0:
            return G.rot(phi, th, om) @ Z2
        if j == 1:
            return G.rz(om) @ Y2 @ G.ry(th) @ G.rz(phi)
        return Z2 @ G.rot(phi, th, om)
    if name == "RX":
        return X2 @ G.rx(theta[0])
    if name == "RY":
        return Y2 @ G.ry(theta[0])
    if name == "RZ":
        return Z2 @ G.rz(theta[0])
    if name in ("PhaseShift", "U1"):
        return np.array([[0, 0], [0, 1j * np.exp(1j * theta[0])]], dtype=np.complex128)
    if name == "U2":
        phi, de = theta
        s = 1.0 / np.sqrt(2.0)
        if j == 0:
            return s * np.array([[0, 0], [1j * np.exp(1j * phi), 1j * np.exp(1j * (phi + de))]])
        return s * np.array([[0, -1j * np.exp(1j * de)], [0, 1j * np.exp(1j * (phi + de))]])
    if name == "U3":
        th, phi, de = theta
        c, s = np.cos(th / 2), np.sin(th / 2)
        if j == 0:
            return 0.5 * np.array([[-s, -np.exp(1j * de) * c],
                                   [np.exp(1j * phi) * c, -np.exp(1j * (phi + de)) * s]])
        if j == 1:
            return np.array([[0, 0], [1j * np.exp(1j * phi) * s, 1j * np.exp(1j * (phi + de)) * c]])
        return np.array([[0, -1j * np.exp(1j * de) * s], [0, 1j * np.exp(1j * (phi + de)) * c]])
    if name in ("CRX", "CRY", "CRZ", "CRot", "ControlledPhaseShift", "CPhase"):
        inner = {"CRX": "RX", "CRY": "RY", "CRZ": "RZ", "CRot": "Rot",
                 "ControlledPhaseShift": "PhaseShift", "CPhase": "PhaseShift"}[name]
        m = np.zeros((4, 4), dtype=np.complex128)
        m[2:, 2:] = dmatrix(inner, theta, j)
        return m
    if name in ("IsingXX", "IsingYY", "IsingZZ"):
        p = _PAULI[name[-1]]
        return (-0.5j * np.kron(p, p)) @ G.matrix(name, theta)
    raise ValueError("gate %s has no parameters" % name)


def _scatter(shape, idx, vals):
    g = np.zeros(shape, dtype=np.float64)
    for t, v in zip(idx, vals):
        g[t] = v
    return g


def grad_adjoint(n, k, pool, params, terms, init=("zero", 0)):
    """Exact gradient by the adjoint method: lambda = H psi, reverse sweep."""
    params = np.asarray(params, dtype=np.float64)
    idx = extract_param_indices(k, pool)
    if not idx:
        return np.zeros(params.shape)
    gl = lower(k, pool, params)
    psi = run(n, gl, init)
    lam = apply_hamiltonian(psi, terms)
    vals_rev = []
    for name, wires, theta in reversed(gl):
        u = G.matrix(name, theta or ())
        ud = u.conj().T
        psi = apply_matrix(psi, ud, wires)            # state before the gate
        if theta:
            for j in reversed(range(len(theta))):
                dpsi = apply_matrix(psi, dmatrix(name, theta, j), wires)
                vals_rev.append(2.0 * np.vdot(lam, dpsi).real)
        lam = apply_matrix(lam, ud, wires)
    return _scatter(params.shape, idx, list(reversed(vals_rev)))


def grad_param_shift(n, k, pool, params, terms, init=("zero", 0)):
    """Two-term parameter-shift rule (what diff_method='parameter-shift' does for the LiH/H2O
    models, qml_models_legacy.py:1520,1631). Valid for gates in gates.TWO_TERM_SHIFT."""
    params = np.asarray(params, dtype=np.float64)
    idx = extract_param_indices(k, pool)
    vals = []
    for (i, key, j) in idx:
        name, _ = _entry(pool, key)
        assert name in G.TWO_TERM_SHIFT, name
        pp = params.copy()
        pp[i, key, j] += np.pi / 2
        ep = energy(n, k, pool, pp, terms, init)
        pp[i, key, j] -= np.pi
        em = energy(n, k, pool, pp, terms, init)
        vals.append(0.5 * (ep - em))
    return _scatter(params.shape, idx, vals)


def grad_fd(n, k, pool, params, terms, init=("zero", 0), h=1e-5):
    params = np.asarray(params, dtype=np.float64)
    idx = extract_param_indices(k, pool)
    vals = []
    for t in idx:
        pp = params.copy()
        pp[t] += h
        ep = energy(n, k, pool, pp, terms, init)
        pp[t] -= 2 * h
        em = energy(n, k, pool, pp, terms, init)
        vals.append((ep - em) / (2 * h))
    return _scatter(params.shape, idx, vals)


def batch_mean_gradient(grads, avg=True):
    """stack -> nan_to_num -> mean (zeros included) or sum; qas/mcts/mcts.py:345-347."""
    g = np.nan_to_num(np.stack(grads, axis=0))
    return np.mean(g, axis=0) if avg else np.sum(g, axis=0)


# ----------------------------------------------------------------------------- Pauli helpers
def word_to_masks(word):
    """Pauli word (wire 0 first = MSB) -> (xmask, zmask): X->x, Z->z, Y->x and z."""
    n = len(word)
    x = z = 0
    for w, ch in enumerate(word):
        bit = 1 << (n - 1 - w)
        if ch in "XY":
            x |= bit
        if ch in "ZY":
            z |= bit
    return x, z


def hamiltonian_matrix(n, terms):
    """Dense 2^n x 2^n matrix (small n only), for eigenvalue cross-checks."""
    dim = 2 ** n
    h = np.zeros((dim, dim), dtype=np.complex128)
    for coeff, word in terms:
        m = np.array([[1.0 + 0j]])
        for ch in word:
            m = np.kron(m, _PAULI[ch])
        h += coeff * m
    return h


# ----------------------------------------------------------------------------- large registers
def run_flat(n, gate_list, init=("zero", 0)):
    """Same evolution as `run`, on a flat 2^n vector with strided views (no tensordot copies);
    used to check the large-state mode at 20-24 qubits in reasonable time."""
    kind, arg = init
    if kind == "plus":
        st = np.full(2 ** n, 2.0 ** (-0.5 * n), dtype=np.complex128)
    else:
        st = np.zeros(2 ** n, dtype=np.complex128)
        st[arg if kind == "basis" else 0] = 1.0
    for name, wires, theta in gate_list:
        m = G.matrix(name, theta or ())
        if len(wires) == 1:
            b = n - 1 - wires[0]
            v = st.reshape(-1, 2, 1 << b)
            a0, a1 = v[:, 0, :].copy(), v[:, 1, :]
            v[:, 0, :] = m[0, 0] * a0 + m[0, 1] * a1
            v[:, 1, :] = m[1, 0] * a0 + m[1, 1] * a1
        elif name in ("CNOT", "CZ", "CY", "CRX", "CRY", "CRZ", "CRot", "CPhase", "ControlledPhaseShift"):
            u = m[2:, 2:]
            bc, bt = n - 1 - wires[0], n - 1 - wires[1]
            hi, lo = max(bc, bt), min(bc, bt)
            v = st.reshape(-1, 2, 1 << (hi - lo - 1), 2, 1 << lo)
            if bc == hi:
                sub0, sub1 = v[:, 1, :, 0, :], v[:, 1, :, 1, :]
            else:
                sub0, sub1 = v[:, 0, :, 1, :], v[:, 1, :, 1, :]
            a0, a1 = sub0.copy(), sub1.copy()
            sub0[...] = u[0, 0] * a0 + u[0, 1] * a1
            sub1[...] = u[1, 0] * a0 + u[1, 1] * a1
        else:
            st = apply_matrix(st.reshape([2] * n), m, wires).reshape(-1)
    return st


def expval_diag_zz_chain(st, n):
    """<sum_i Z_i Z_{i+1}> on a flat vector (wire 0 = MSB)."""
    p = np.abs(st) ** 2
    idx = np.arange(st.size, dtype=np.int64)
    e = 0.0
    for i in range(n - 1):
        b0, b1 = n - 1 - i, n - 2 - i
        sign = 1.0 - 2.0 * (((idx >> b0) ^ (idx >> b1)) & 1)
        e += float(np.dot(p, sign))
    return e
