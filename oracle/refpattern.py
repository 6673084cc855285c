"""CPU baseline: the oracle arranged in the reference's CALL PATTERN (test/bench infrastructure).

The literal reference cannot run here (PennyLane is not installed), so the
CPU baseline times this restatement of what one reference evaluation does:

  * a new model object per evaluation (qas/mcts/mcts.py:103, :340), which re-extracts the
    parameter indices (utils.py:11-25);
  * per-gate tensordot on a [2]*n complex128 array (PennyLane default.qubit);
  * the expectation the model's device computes: term-by-term Pauli expectation for the
    qml.Hamiltonian models, a scipy sparse mat-vec for the models that wrap
    qml.SparseHamiltonian (LiH / H2O, qml_models_legacy.py:1494,1524,1605,1634);
  * the gradient by the method class the model selects: 2 * (#parameters) shifted executions
    for LiH / H2O (diff_method="parameter-shift", :1520,:1631), one reverse-mode sweep
    for the others (autograd backprop / lightning adjoint).

It omits PennyLane's device/QNode construction and autograd tracing, so it is a
CONSERVATIVE (too fast) stand-in for the reference.
"""
import numpy as np

from . import sim


def sparse_hamiltonian(n, terms):
    """scipy CSR matrix of the Pauli sum (what qml.utils.sparse_hamiltonian builds)."""
    import scipy.sparse as sp
    dim = 1 << n
    idx = np.arange(dim)
    rows, cols, vals = [], [], []
    for coeff, word in terms:
        x, z = sim.word_to_masks(word)
        ny = bin(x & z).count("1")
        par = np.zeros(dim, dtype=np.int64)
        zz = z
        while zz:
            b = zz & -zz
            par ^= (idx & b) != 0
            zz ^= b
        # <i|P|j> nonzero for i = j ^ x
        rows.append(idx ^ x)
        cols.append(idx)
        vals.append(coeff * (1j ** ny) * np.where(par == 1, -1.0, 1.0))
    m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim))
    return m.tocsr()


class RefPatternEvaluator:
    """One (n, pool, Hamiltonian, init) family; `evaluate_one` = what the reference does per circuit."""

    def __init__(self, n, pool, terms, init=("zero", 0), grad_method="adjoint", sparse=False):
        self.n, self.pool, self.terms, self.init = n, pool, terms, init
        self.grad_method = grad_method
        self.hs = sparse_hamiltonian(n, terms) if sparse else None

    def _energy(self, k, params):
        _ = sim.extract_param_indices(k, self.pool)          # model construction
        st = sim.run(self.n, sim.lower(k, self.pool, params), self.init)
        if self.hs is not None:
            v = st.reshape(-1)
            return float(np.vdot(v, self.hs @ v).real)
        return sim.expval(st, self.terms)

    def reward(self, k, params):
        return -self._energy(k, params)

    def gradient(self, k, params):
        params = np.asarray(params, dtype=np.float64)
        idx = sim.extract_param_indices(k, self.pool)
        g = np.zeros(params.shape)
        if not idx:
            return g
        if self.grad_method == "parameter-shift":
            for t in idx:
                pp = params.copy()
                pp[t] += np.pi / 2
                ep = self._energy(k, pp)
                pp[t] -= np.pi
                em = self._energy(k, pp)
                g[t] = 0.5 * (ep - em)
            return g
        if self.hs is None:
            return sim.grad_adjoint(self.n, k, self.pool, params, self.terms, self.init)
        # reverse-mode with the sparse operator
        gl = sim.lower(k, self.pool, params)
        psi = sim.run(self.n, gl, self.init)
        lam = (self.hs @ psi.reshape(-1)).reshape(psi.shape)
        from . import gates as G
        vals = []
        for name, wires, theta in reversed(gl):
            ud = G.matrix(name, theta or ()).conj().T
            psi = sim.apply_matrix(psi, ud, wires)
            if theta:
                for j in reversed(range(len(theta))):
                    vals.append(2.0 * np.vdot(lam, sim.apply_matrix(psi, sim.dmatrix(name, theta, j), wires)).real)
            lam = sim.apply_matrix(lam, ud, wires)
        for t, v in zip(idx, reversed(vals)):
            g[t] = v
        return g

    def evaluate_one(self, k, params):
        """One candidate evaluation = energy + gradient (the unit of BASELINE.json's metric)."""
        k = [int(x) for x in k]
        return self._energy(k, params), self.gradient(k, params)


# method class per reference model
REF_METHOD = {
    "FourQubitH2": ("adjoint", False), "FourQubitH2_VaccumInitial": ("adjoint", False),
    "QAOAVQCDemo": ("adjoint", False), "QAOAWeightedVQCDemo": ("adjoint", False),
    "TwoQubitH2": ("adjoint", False),
    "LiH": ("parameter-shift", True), "H2O": ("parameter-shift", True),
    "H2STO3G": ("adjoint", False), "LiHSTO3G": ("adjoint", False), "H2OSTO3G": ("adjoint", False),
    "KagomeHeisenberg": ("adjoint", False),
}
