"""Vectorised sampler of legal structure lists (synthetic candidate batches).

Draws B structure lists the way the reference's warm-up phase does on a fresh tree
(``MCTSController.randomSample``, qas/mcts/mcts.py:107-114: expand with a uniformly random
legal, not-yet-expanded action) under the legality rules of ``QMLStateBasicGates``
(qas/mcts/qml_mcts_state.py:58-134), but for the whole batch at once with
NumPy: per depth step one [B, c] legality mask, then one uniform draw per row.
Used by bench.py, the tests and smoke(); the object-per-node form of the same rules is
``qas_b200.mcts.qml_mcts_state.QMLStateBasicGates``.
"""
import numpy as np

from qas_b200.qml_models.qml_gate_ops import SUPPORTED_OPS_DICT

_PAULI_LIKE = ("PauliX", "PauliY", "PauliZ", "Hadamard")
_CANCEL = {"T": "Tdg", "Tdg": "T", "S": "Sdg", "Sdg": "S"}
_DOUBLE = {"T": "S", "Tdg": "Sdg", "S": "PauliZ", "Sdg": "PauliZ"}


def legal_mask(entries, single_names, last_op, touched, counts, limits):
    """entries: [(name, wires)] per key; last_op [B, n] (last non-PlaceHolder key stacked on each
    qubit, -1 if none); touched [B, n] bool; counts {name: [B]}; -> bool [B, c]."""
    B = last_op.shape[0]
    c = len(entries)
    ok = np.ones((B, c), dtype=bool)
    name_of = [e[0] for e in entries]
    for key, (name, wires) in enumerate(entries):
        spec = SUPPORTED_OPS_DICT[name]
        if len(wires) == 1:
            q = wires[0]
            last = last_op[:, q]
            if spec.num_params > 0 or name in _PAULI_LIKE:
                ok[:, key] &= last != key
            elif name in _CANCEL:
                has = last >= 0
                last_name_is = np.zeros(B, dtype=bool)
                for k2, n2 in enumerate(name_of):
                    if n2 == _CANCEL[name]:
                        last_name_is |= last == k2
                ok[:, key] &= ~(has & last_name_is)
                if _DOUBLE[name] in single_names:
                    ok[:, key] &= last != key
            if name == "PlaceHolder":
                ok[:, key] &= touched[:, q]
        else:
            a, b = wires
            if spec.num_params > 0 or name in ("CNOT", "CZ", "CY"):
                ok[:, key] &= ~((last_op[:, a] == key) & (last_op[:, b] == key))
            ok[:, key] &= touched[:, a]
        if name in limits:
            ok[:, key] &= counts[name] < limits[name]
    return ok


def sample_structure_lists(op_pool, p, gate_limit_dict=None, B=1, seed=0, init_qubit_with_actions=()):
    """-> int32 [B, p] of legal structure lists (uniform over legal actions at every depth)."""
    rng = np.random.default_rng(seed)
    src = getattr(op_pool, "pool", op_pool)
    c = len(src)
    entries = []
    for key in range(c):
        (name, wires), = src[key].items()
        entries.append((name, list(wires)))
    n = op_pool.num_qubits if hasattr(op_pool, "num_qubits") else 1 + max(max(w) for _, w in entries)
    single_names = list(getattr(op_pool, "single_qubit_gates", []))
    limits = dict(gate_limit_dict or {})
    last_op = np.full((B, n), -1, dtype=np.int64)
    touched = np.zeros((B, n), dtype=bool)
    for q in init_qubit_with_actions:
        touched[:, q] = True
    counts = {name: np.zeros(B, dtype=np.int64) for name in limits}
    ks = np.zeros((B, p), dtype=np.int32)
    rows = np.arange(B)
    for depth in range(p):
        ok = legal_mask(entries, single_names, last_op, touched, counts, limits)
        nlegal = ok.sum(axis=1)
        if np.any(nlegal == 0):
            raise ValueError("no legal action at depth %d for some circuit" % depth)
        # uniform draw among the legal keys of each row
        pick = (rng.random(B) * nlegal).astype(np.int64)
        cum = np.cumsum(ok, axis=1)
        key = np.argmax(cum > pick[:, None], axis=1)
        ks[:, depth] = key
        for kk, (name, wires) in enumerate(entries):
            sel = key == kk
            if not sel.any():
                continue
            if name != "PlaceHolder":
                for q in wires:
                    last_op[sel, q] = kk
                touched[sel, wires[-1]] = True
            if name in limits:
                counts[name][sel] += 1
    return ks
