"""CPU restatement of the reference's VQLS cost exactly as it computes it -- TEST INFRASTRUCTURE ONLY:
one Hadamard-test circuit on n + 1 qubits per (l, l', j, Re/Im), qml_models_legacy.py:1880-1951
(4-qubit problem) and :2144-2215 (5-qubit problem).  Used to check the two-observable form of
qas_b200/qml_models/vqls_model.py."""
import numpy as np

from oracle import gates as G
from oracle import sim


def _controlled_word(st, n, word):
    """Apply the Pauli word on the n system wires controlled by the ancilla (wire n): CNOT / CZ per letter
    (CA, :1838-1856)."""
    for w, ch in enumerate(word):
        if ch == "X":
            st = sim.apply_matrix(st, G.matrix("CNOT"), [n, w])
        elif ch == "Z":
            st = sim.apply_matrix(st, G.matrix("CZ"), [n, w])
        elif ch != "I":
            raise ValueError("the reference's problems only hold X and Z letters")
    return st


def hadamard_test(n, gate_list, word_l, word_lp, j, part):
    """<Z_ancilla> of local_hadamard_test (:1881-1918)."""
    st = np.zeros(2 ** (n + 1), dtype=np.complex128)
    st[0] = 1.0
    st = st.reshape([2] * (n + 1))
    H = G.matrix("Hadamard")
    st = sim.apply_matrix(st, H, [n])
    if part == "Im":
        st = sim.apply_matrix(st, G.matrix("PhaseShift", [-np.pi / 2]), [n])
    for w in range(n):                                   # backboneCirc starts with U_b (:1858-1860)
        st = sim.apply_matrix(st, H, [w])
    for name, wires, theta in gate_list:
        st = sim.apply_matrix(st, G.matrix(name, theta or ()), wires)
    st = _controlled_word(st, n, word_l)
    for w in range(n):
        st = sim.apply_matrix(st, H, [w])
    if j != -1:
        st = sim.apply_matrix(st, G.matrix("CZ"), [n, j])
    for w in range(n):
        st = sim.apply_matrix(st, H, [w])
    st = _controlled_word(st, n, word_lp)
    st = sim.apply_matrix(st, H, [n])
    prob = np.abs(st.reshape(-1)) ** 2
    return float(np.sum(prob[0::2]) - np.sum(prob[1::2]))      # ancilla is the last wire: lowest index bit


def cost_loc(n, problem, k, pool, params):
    """cost_loc (:1939-1951): 0.5 - 0.5 |sum mu(l,l',j)| / (n |sum mu(l,l',-1)|)."""
    gl = sim.lower(k, pool, params)

    def mu(l, lp, j):
        re = hadamard_test(n, gl, problem[l][1], problem[lp][1], j, "Re")
        im = hadamard_test(n, gl, problem[l][1], problem[lp][1], j, "Im")
        return re + 1j * im

    L = len(problem)
    norm = sum(problem[l][0] * np.conj(problem[lp][0]) * mu(l, lp, -1) for l in range(L) for lp in range(L))
    mu_sum = sum(problem[l][0] * np.conj(problem[lp][0]) * mu(l, lp, j)
                 for l in range(L) for lp in range(L) for j in range(n))
    return 0.5 - 0.5 * abs(mu_sum) / (n * abs(norm))
