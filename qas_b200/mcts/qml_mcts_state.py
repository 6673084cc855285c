"""MCTS states for the QML gate pool (drop-in for qas/mcts/qml_mcts_state.py).

Same constructor, attributes and methods as upstream ``QMLStateBasicGates`` (:12-163) and
``QMLStateBasicGatesNoRestrictions`` (:165-226), same legality rules, but the
"last operator stacked on each qubit" table is carried incrementally from parent to child
instead of being rebuilt from the whole path for every candidate action (upstream is
O(c * depth) per expansion, :41-43,66).
"""
from typing import List, Optional, Set

from qas_b200.mcts.mcts_state import StateOfMCTS
from qas_b200.qml_models.qml_gate_ops import SUPPORTED_OPS_DICT, QMLPool

_PAULI_LIKE = ("PauliX", "PauliY", "PauliZ", "Hadamard")
_INVERSE = {"T": "Tdg", "Tdg": "T", "S": "Sdg", "Sdg": "S"}
_SQUARE = {"T": "S", "Tdg": "Sdg", "S": "PauliZ", "Sdg": "PauliZ"}


def _entry(pool_dict, key):
    (name, wires), = pool_dict[key].items()
    return name, wires


class QMLStateBasicGates(StateOfMCTS):
    name = "QMLStateBasicGates"

    def __init__(self, current_k: List[int] = [], op_pool: QMLPool = None, maxDepth=30,
                 qubit_with_actions: Set = None, gate_limit_dict: Optional[dict] = None,
                 _last_on_qubit=None, _gate_count=None):
        self.max_depth = maxDepth
        self.current_k = current_k
        self.current_depth = len(current_k)
        assert self.current_depth <= self.max_depth
        self.pool_obj = op_pool
        self.op_name_dict = op_pool.pool
        self.pool_keys = list(op_pool.pool.keys())
        self.single_qubit_gate_names = op_pool.single_qubit_gates
        self.state = None
        self.qubit_with_actions = qubit_with_actions if qubit_with_actions is not None else set()
        self.gate_limit_dict = gate_limit_dict if gate_limit_dict is not None else {}
        if _last_on_qubit is not None:
            self._last = _last_on_qubit
            self.gate_count = _gate_count
        else:
            self._last = [-1] * op_pool.num_qubits
            self.gate_count = {name: 0 for name in self.gate_limit_dict}
            for action in current_k:
                name, wires = _entry(self.op_name_dict, action)
                if name != "PlaceHolder":
                    for q in wires:
                        self._last[q] = action
                if name in self.gate_count:
                    self.gate_count[name] += 1

    def getLegalActions(self):
        if self.current_depth == self.max_depth:
            return []
        return [key for key in self.pool_keys if self.verifyDesirableAction(key)]

    def stackOpsOnQubit(self, k):
        stacked = [[] for _ in range(self.pool_obj.num_qubits)]
        for action in k:
            name, wires = _entry(self.op_name_dict, action)
            if name != "PlaceHolder":
                for q in wires:
                    stacked[q].append(action)
        return stacked

    def verifyDesirableAction(self, action: int):
        name, wires = _entry(self.op_name_dict, action)
        npar = SUPPORTED_OPS_DICT[name].num_params
        last = self._last
        if len(wires) == 1:
            prev = last[wires[0]]
            if npar > 0 or name in _PAULI_LIKE:
                if prev == action:
                    return False
            elif name in _INVERSE and prev >= 0:
                prev_name, _ = _entry(self.op_name_dict, prev)
                if prev_name == _INVERSE[name]:
                    return False
                if prev == action and _SQUARE[name] in self.single_qubit_gate_names:
                    return False
        elif len(wires) == 2:
            if last[wires[0]] == action and last[wires[1]] == action:
                if npar > 0 or name in ("CNOT", "CZ", "CY"):
                    return False
            if wires[0] not in self.qubit_with_actions:
                return False
        if name == "PlaceHolder" and wires[0] not in self.qubit_with_actions:
            return False
        if name in self.gate_limit_dict and self.gate_count[name] >= self.gate_limit_dict[name]:
            return False
        return True

    def takeAction(self, action: int):
        name, wires = _entry(self.op_name_dict, action)
        touched = set(self.qubit_with_actions)
        last = list(self._last)
        if name != "PlaceHolder":
            touched.add(wires[-1])
            for q in wires:
                last[q] = action
        count = dict(self.gate_count)
        if name in count:
            count[name] += 1
        child = QMLStateBasicGates(current_k=self.current_k + [action], op_pool=self.pool_obj,
                                   maxDepth=self.max_depth, qubit_with_actions=touched,
                                   gate_limit_dict=self.gate_limit_dict, _last_on_qubit=last,
                                   _gate_count=count)
        child.state = action
        return child

    def isTerminal(self):
        return self.current_depth >= self.max_depth

    def getReward(self):
        return

    def getCurrK(self):
        return self.current_k

    def __repr__(self):
        return "".join("OpAtDepth: {}\tOpKey: {}\tOpName: {}\n".format(i, key, self.op_name_dict[key])
                       for i, key in enumerate(self.current_k))


class QMLStateBasicGatesNoRestrictions(QMLStateBasicGates):
    """Every pool key is legal at every depth (upstream :165-226; note that upstream's
    takeAction returns a restricted QMLStateBasicGates child, so only the root is
    unrestricted there -- that behaviour is kept)."""
    name = "QMLStateBasicGatesNoRestrictions"

    def getLegalActions(self):
        if self.current_depth == self.max_depth:
            return []
        return list(self.pool_keys)

    def getReward(self):
        return 0.
