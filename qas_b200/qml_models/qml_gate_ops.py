"""PennyLane-free gate vocabulary and operator pool (drop-in for the reference's
``qas/qml_models/qml_gate_ops.py``).

Upstream, ``SUPPORTED_OPS_DICT`` maps gate names to PennyLane operator classes (:66-105) and
the rest of QAS only reads ``.num_params`` / ``.num_wires`` from them (``QMLGate.__init__``
:113-124, ``QMLPool.__init__`` :141-146, ``extractParamIndicesQML`` utils.py:21-22,
``verifyDesirableAction`` qml_mcts_state.py:64-65).  Here each name maps to a small
``GateSpec`` carrying the same two attributes plus the gate id of the C ABI; the arithmetic
lives in the CUDA kernels, not in Python objects.

Differences from upstream, all supersets: 'Sdg' and 'Tdg' are real one-wire gates here
(upstream they are ``qml.adjoint(S)`` closures without ``num_wires`` and cannot enter a pool);
names with three or more wires (CSWAP, Toffoli, MultiControlledX) and the variable-arity
MultiRZ are kept in the dictionary for ``name in SUPPORTED_OPS_NAME`` checks but, exactly as
upstream, can never pass ``QMLPool``'s one-/two-wire assertions.
"""
import json
from typing import AnyStr, List, Optional, Sequence

from qas_b200.ops import Pool, QuantumGate
from qas_b200 import _lib


class GateSpec:
    """Stand-in for a PennyLane operator class: (name, num_wires, num_params, gate id)."""
    __slots__ = ("name", "num_wires", "num_params", "gate_id")

    def __init__(self, name, num_wires, num_params):
        self.name = name
        self.num_wires = num_wires
        self.num_params = num_params
        self.gate_id = _lib.GATE_IDS.get(name, -1)

    def __call__(self, *params, wires=None):
        return GateOp(self, list(params), list(wires) if wires is not None else None)

    def __repr__(self):
        return "GateSpec(%s)" % self.name


class GateOp:
    """What ``QMLGate.getOp()`` returns: a plain record (spec, parameters, wires)."""
    __slots__ = ("spec", "parameters", "wires")

    def __init__(self, spec, parameters, wires):
        self.spec, self.parameters, self.wires = spec, parameters, wires

    @property
    def name(self):
        return self.spec.name

    def __repr__(self):
        return "%s(%s, wires=%s)" % (self.spec.name, ", ".join(repr(x) for x in self.parameters), self.wires)


_SPECS = [
    ("Hadamard", 1, 0), ("PauliX", 1, 0), ("PauliY", 1, 0), ("PauliZ", 1, 0), ("S", 1, 0),
    ("Sdg", 1, 0), ("T", 1, 0), ("Tdg", 1, 0), ("SX", 1, 0),
    ("CNOT", 2, 0), ("CZ", 2, 0), ("CY", 2, 0), ("SWAP", 2, 0), ("ISWAP", 2, 0), ("SISWAP", 2, 0),
    ("SQISW", 2, 0), ("CSWAP", 3, 0), ("Toffoli", 3, 0), ("MultiControlledX", -1, 0),
    ("Rot", 1, 3), ("RX", 1, 1), ("RY", 1, 1), ("RZ", 1, 1), ("MultiRZ", -1, 1),
    ("PhaseShift", 1, 1), ("ControlledPhaseShift", 2, 1), ("CPhase", 2, 1),
    ("CRX", 2, 1), ("CRY", 2, 1), ("CRZ", 2, 1), ("CRot", 2, 3),
    ("U1", 1, 1), ("U2", 1, 2), ("U3", 1, 3),
    ("IsingXX", 2, 1), ("IsingYY", 2, 1), ("IsingZZ", 2, 1),
    ("PlaceHolder", 1, 0),
]

SUPPORTED_OPS_DICT = {name: GateSpec(name, nw, npar) for name, nw, npar in _SPECS}
PlaceHolder = SUPPORTED_OPS_DICT["PlaceHolder"]
SUPPORTED_OPS_NAME = set(SUPPORTED_OPS_DICT.keys())


class QMLGate(QuantumGate):
    """(name, wires, params) with the reference's checks (qml_gate_ops.py:111-134)."""

    def __init__(self, name: str, pos: List[int], param: Optional[Sequence] = None):
        assert name in SUPPORTED_OPS_NAME
        spec = SUPPORTED_OPS_DICT[name]
        self.num_params = spec.num_params
        self.num_wires = spec.num_wires
        if param is not None:
            assert len(param) == self.num_params
        assert len(pos) == self.num_wires
        self.wires = pos
        self.name = name
        self.param = param
        self.op = spec(*param, wires=pos) if param is not None else spec(wires=pos)

    def getOp(self):
        return self.op

    def getQregPos(self):
        return list(self.wires)

    def __str__(self):
        return "{}, {}, {}".format(self.name, self.wires, list(self.param))


class QMLPool(Pool):
    """Operator pool with the reference's key layout (qml_gate_ops.py:136-185):
    for every qubit, one key per single-qubit gate name (in list order); then for every directed
    coupling, one key per two-qubit gate name."""

    def __init__(self, num_qubits: int, single_qubit_op_generator: List[AnyStr],
                 two_qubit_op_generator: List[AnyStr], complete_undirected_graph: bool = True,
                 two_qubit_gate_map: Optional[List[List[int]]] = None):
        super().__init__()
        assert all(SUPPORTED_OPS_DICT[g].num_wires == 1 for g in single_qubit_op_generator)
        assert all(SUPPORTED_OPS_DICT[g].num_wires == 2 for g in two_qubit_op_generator)
        if two_qubit_gate_map is None:
            self.coupling = [(a, b) for a in range(num_qubits) for b in range(num_qubits) if a != b]
        else:
            assert complete_undirected_graph is False
            for pair in two_qubit_gate_map:
                assert len(pair) == 2 and pair[0] != pair[1]
                assert 0 <= pair[0] < num_qubits and 0 <= pair[1] < num_qubits
            self.coupling = two_qubit_gate_map
        self.num_qubits = num_qubits
        self.single_qubit_gates = single_qubit_op_generator
        self.two_qubit_gates = two_qubit_op_generator
        entries = [{g: [q]} for q in range(num_qubits) for g in single_qubit_op_generator]
        entries += [{g: edge} for edge in self.coupling for g in two_qubit_op_generator]
        self.pool = dict(enumerate(entries))

    def __getitem__(self, key):
        return self.pool[key]

    def __str__(self):
        return json.dumps(self.pool)

    def __len__(self):
        return len(self.pool)


def pool_table(op_pool):
    """QMLPool / plain dict -> [(gate_id, wire0, wire1)] in key order, for the C ABI."""
    src = getattr(op_pool, "pool", op_pool)
    # the batched entries look their evaluator up by this table on every call: keep it with the pool object
    # (pools are built once per experiment and never edited -- qas/qml_models/qml_gate_ops.py:136-185)
    cached = getattr(op_pool, "_qas_pool_table", None) if isinstance(op_pool, QMLPool) else None
    if cached is not None and cached[0] == len(src):
        return cached[1]
    out = []
    for key in range(len(src)):
        entry = src[key] if key in src else src[str(key)]
        assert len(entry) == 1
        (name, wires), = entry.items()
        gid = SUPPORTED_OPS_DICT[name].gate_id
        if gid < 0:
            raise ValueError("gate %s is not supported by the evaluator" % name)
        w = list(wires)
        out.append((gid, int(w[0]), int(w[1]) if len(w) > 1 else -1))
    if isinstance(op_pool, QMLPool):
        op_pool._qas_pool_table = (len(src), out)
    return out
