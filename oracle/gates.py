"""Gate matrices of the QAS gate set (oracle side; test infrastructure only).

The reference maps gate names to PennyLane operator classes in
``qas/qml_models/qml_gate_ops.py:66-105`` (``SUPPORTED_OPS_DICT``).  PennyLane 0.29.1 is
not installed, so the matrices below restate its documented definitions.  Conventions
that matter were verified against the reference's saved results:
``Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi)``, ``U3`` as in the docstring of
``u3``, ``CNOT(wires=[control, target])``, wire 0 = most significant bit.

Every entry of GATES is ``(num_wires, num_params, matrix_fn)`` where ``matrix_fn(*params)``
returns the 2x2 or 4x4 complex128 matrix in the computational basis ordered with the
first listed wire as the more significant bit.
"""
import numpy as np

_I2 = np.eye(2, dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2.0)


def rx(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def ry(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def rz(t):
    return np.array([[np.exp(-0.5j * t), 0], [0, np.exp(0.5j * t)]], dtype=np.complex128)


def phase_shift(t):
    return np.array([[1, 0], [0, np.exp(1j * t)]], dtype=np.complex128)


def rot(phi, theta, omega):
    """Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi)."""
    return rz(omega) @ ry(theta) @ rz(phi)


def u2(phi, delta):
    return np.array([[1, -np.exp(1j * delta)],
                     [np.exp(1j * phi), np.exp(1j * (phi + delta))]], dtype=np.complex128) / np.sqrt(2.0)


def u3(theta, phi, delta):
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[c, -np.exp(1j * delta) * s],
                     [np.exp(1j * phi) * s, np.exp(1j * (phi + delta)) * c]], dtype=np.complex128)


def _const(m):
    m = np.array(m, dtype=np.complex128)
    return lambda: m.copy()


def controlled(u):
    """|0><0| (x) I + |1><1| (x) u, control = first wire."""
    m = np.eye(4, dtype=np.complex128)
    m[2:, 2:] = u
    return m


def ising_xx(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    m = np.eye(4, dtype=np.complex128) * c
    m[0, 3] = m[1, 2] = m[2, 1] = m[3, 0] = -1j * s
    return m


def ising_yy(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    m = np.eye(4, dtype=np.complex128) * c
    m[0, 3] = m[3, 0] = 1j * s
    m[1, 2] = m[2, 1] = -1j * s
    return m


def ising_zz(t):
    e, f = np.exp(-0.5j * t), np.exp(0.5j * t)
    return np.diag(np.array([e, f, f, e], dtype=np.complex128))


def multi_rz_1(t):
    return rz(t)


_SQ = 1.0 / np.sqrt(2.0)
_SWAP = [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]
_ISWAP = [[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]
_SISWAP = [[1, 0, 0, 0], [0, _SQ, 1j * _SQ, 0], [0, 1j * _SQ, _SQ, 0], [0, 0, 0, 1]]

# name -> (num_wires, num_params, matrix function)
GATES = {
    "Hadamard": (1, 0, _const(_H)),
    "PauliX": (1, 0, _const(_X)),
    "PauliY": (1, 0, _const(_Y)),
    "PauliZ": (1, 0, _const(_Z)),
    "S": (1, 0, _const([[1, 0], [0, 1j]])),
    "Sdg": (1, 0, _const([[1, 0], [0, -1j]])),
    "T": (1, 0, _const([[1, 0], [0, np.exp(0.25j * np.pi)]])),
    "Tdg": (1, 0, _const([[1, 0], [0, np.exp(-0.25j * np.pi)]])),
    "SX": (1, 0, _const(0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]))),
    "PlaceHolder": (1, 0, _const(_I2)),
    "Rot": (1, 3, rot),
    "RX": (1, 1, rx),
    "RY": (1, 1, ry),
    "RZ": (1, 1, rz),
    "PhaseShift": (1, 1, phase_shift),
    "U1": (1, 1, phase_shift),
    "U2": (1, 2, u2),
    "U3": (1, 3, u3),
    "CNOT": (2, 0, _const(controlled(_X))),
    "CZ": (2, 0, _const(controlled(_Z))),
    "CY": (2, 0, _const(controlled(_Y))),
    "SWAP": (2, 0, _const(_SWAP)),
    "ISWAP": (2, 0, _const(_ISWAP)),
    "SISWAP": (2, 0, _const(_SISWAP)),
    "SQISW": (2, 0, _const(_SISWAP)),
    "CRX": (2, 1, lambda t: controlled(rx(t))),
    "CRY": (2, 1, lambda t: controlled(ry(t))),
    "CRZ": (2, 1, lambda t: controlled(rz(t))),
    "CRot": (2, 3, lambda a, b, c: controlled(rot(a, b, c))),
    "ControlledPhaseShift": (2, 1, lambda t: controlled(phase_shift(t))),
    "CPhase": (2, 1, lambda t: controlled(phase_shift(t))),
    "IsingXX": (2, 1, ising_xx),
    "IsingYY": (2, 1, ising_yy),
    "IsingZZ": (2, 1, ising_zz),
}

# Parameters for which the two-term parameter-shift rule (shift pi/2, factor 1/2) is exact:
# the parameter enters through exp(-i t G) with G having two eigenvalues +-1/2.
TWO_TERM_SHIFT = {"Rot", "RX", "RY", "RZ", "PhaseShift", "U1", "U2", "U3",
                  "IsingXX", "IsingYY", "IsingZZ"}


def num_params(name):
    return GATES[name][1]


def num_wires(name):
    return GATES[name][0]


def matrix(name, params=()):
    nw, npar, fn = GATES[name]
    assert len(params) == npar, (name, params)
    return fn(*params)


def dmatrix_fd(name, params, j, h=1e-6):
    """Central-difference derivative of the gate matrix w.r.t. parameter j (used only as a
    cross-check of the analytic adjoint path)."""
    pp = list(params)
    pp[j] = params[j] + h
    a = matrix(name, pp)
    pp[j] = params[j] - h
    b = matrix(name, pp)
    return (a - b) / (2 * h)
