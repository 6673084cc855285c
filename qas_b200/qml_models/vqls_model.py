"""Variational quantum linear solver cost (drop-in for ``VQLSDemo`` / ``VQLSDemo5Q``,
qas/qml_models/qml_models_legacy.py:1807-2062, :2064-2326) on the B200 evaluator.

Upstream one loss value is 2 * L^2 * (n + 1) Hadamard-test circuits on n + 1 qubits (L = number of
unitaries in A = sum_l c_l A_l; 160 circuits for the 4-qubit problem, :1880-1951):

    mu(l, l', j) = <x| A_l'^ U_b Z_j U_b^ A_l |x>          (Z_j -> 1 for j = -1; U_b = H^(x)n)
    C_L = 1/2 - 1/2 * |sum_{l,l',j} c_l c_l'* mu(l,l',j)| / (n * |sum_{l,l'} c_l c_l'* mu(l,l',-1)|)

With |x> = V(k, theta) H^(x)n |0>, Hermitian Pauli words A_l, real c_l and H Z_j H = X_j the two sums are
expectation values of two Hermitian Pauli sums in the SAME state,

    numerator   N = <x| sum_j A X_j A |x>,      denominator   D = <x| A A |x>,

so one loss is two evaluations of the ansatz on n qubits instead of 160 on n + 1, and a batch of
candidate circuits is two launches of the batched evaluator (uniform start state).  The gradient follows
from the quotient rule on the two adjoint gradients.  reward = exp(-10 loss) (:1969-1972).
"""
import ctypes
from typing import List, Sequence, Union

import numpy as np

from qas_b200 import _lib
from qas_b200.evaluator import BatchEvaluator
from qas_b200.models import ModelFromK
from qas_b200.qml_models.hamiltonian_model import default_device
from qas_b200.qml_models.qml_gate_ops import SUPPORTED_OPS_DICT, QMLPool, pool_table
from qas_b200.qml_models.utils import extractParamIndicesQML

_EVALUATORS = {}
# single-qubit Pauli products: (a, b) -> (phase, a*b)
_MUL = {("I", "I"): (1, "I"), ("I", "X"): (1, "X"), ("I", "Y"): (1, "Y"), ("I", "Z"): (1, "Z"),
        ("X", "I"): (1, "X"), ("X", "X"): (1, "I"), ("X", "Y"): (1j, "Z"), ("X", "Z"): (-1j, "Y"),
        ("Y", "I"): (1, "Y"), ("Y", "X"): (-1j, "Z"), ("Y", "Y"): (1, "I"), ("Y", "Z"): (1j, "X"),
        ("Z", "I"): (1, "Z"), ("Z", "X"): (1j, "Y"), ("Z", "Y"): (-1j, "X"), ("Z", "Z"): (1, "I")}


def pauli_product(a, b):
    """[(c, word)] * [(c, word)] -> merged [(c, word)] (complex coefficients)."""
    acc = {}
    for ca, wa in a:
        for cb, wb in b:
            ph, w = 1.0 + 0j, []
            for x, y in zip(wa, wb):
                f, z = _MUL[(x, y)]
                ph *= f
                w.append(z)
            w = "".join(w)
            acc[w] = acc.get(w, 0j) + ca * cb * ph
    return [(c, w) for w, c in acc.items() if abs(c) > 1e-15]


def _hermitian(terms):
    assert all(abs(complex(c).imag) < 1e-12 for c, _ in terms), "operator is not Hermitian"
    return [(float(complex(c).real), w) for c, w in terms]


class _VQLSModel(ModelFromK):
    name = "VQLS"
    n_qubits: int = None
    problem = None                       # A = sum_l c_l A_l as [(c_l, Pauli word)]

    def __init__(self, p: int, c: int, l: int, structure_list: List[int], op_pool: Union[QMLPool, dict]):
        self.k = structure_list
        self.pool = op_pool
        self.p, self.c, self.l = p, c, l
        self.param_indices = extractParamIndicesQML(self.k, self.pool)

    # ------------------------------------------------------------------ the two observables
    @classmethod
    def denominator_terms(cls):
        return _hermitian(pauli_product(cls.problem, cls.problem))

    @classmethod
    def numerator_terms(cls):
        n = cls.n_qubits
        acc = []
        for j in range(n):
            xj = [(1.0, "I" * j + "X" + "I" * (n - 1 - j))]
            acc += pauli_product(pauli_product(cls.problem, xj), cls.problem)
        merged = {}
        for c_, w in acc:
            merged[w] = merged.get(w, 0j) + c_
        return _hermitian([(c_, w) for w, c_ in merged.items() if abs(c_) > 1e-15])

    @classmethod
    def _evaluators(cls, op_pool, l=3, device=None):
        device = default_device() if device is None else device
        key = (cls, tuple(pool_table(op_pool)), l, device)
        ev = _EVALUATORS.get(key)
        if ev is None:
            ev = (BatchEvaluator(cls.n_qubits, cls.numerator_terms(), op_pool, ("plus", 0), l, device),
                  BatchEvaluator(cls.n_qubits, cls.denominator_terms(), op_pool, ("plus", 0), l, device))
            _EVALUATORS[key] = ev
        return ev

    @classmethod
    def loss_from(cls, N, D):
        return 0.5 - 0.5 * np.abs(N) / (cls.n_qubits * np.abs(D))

    @staticmethod
    def rewards_from_losses(losses):
        return np.exp(-10.0 * np.asarray(losses, dtype=np.float64))          # :1969-1972

    @classmethod
    def evaluate_batch(cls, ks, super_circ_params, op_pool, want_grad=True, avg=True, compact=False):
        """Losses [B] and the batch-mean (or sum) dense gradient [p, c, l] (mcts.py:340-347).  Two
        evaluator launches on device buffers; the quotient rule (qas_vqls_combine) and the batch mean
        (qas_batch_mean: the evaluator's deterministic two-stage reduction) are library kernels too, so only the
        losses and the dense gradient come back."""
        import torch
        ks = np.ascontiguousarray(np.asarray(ks, dtype=np.int32))
        if ks.ndim == 1:
            ks = ks[None, :]
        params = np.ascontiguousarray(np.asarray(super_circ_params, dtype=np.float64))
        p, c, l = params.shape
        B = len(ks)
        assert ks.shape[1] == p
        if ks.size and (ks.min() < 0 or ks.max() >= c):
            raise AssertionError("structure list entry out of range: need 0 <= k < c")      # utils.py:12-15
        ev_n, ev_d = cls._evaluators(op_pool, l)
        dev = torch.device("cuda", ev_n.device)
        stream = torch.cuda.current_stream(dev)
        with torch.cuda.device(dev):
            d_k = torch.from_numpy(ks).to(dev)
            d_p = torch.from_numpy(params).to(dev)
            N = torch.empty(B, dtype=torch.float64, device=dev)
            D = torch.empty(B, dtype=torch.float64, device=dev)
            gN = torch.empty(B, p, l, dtype=torch.float64, device=dev) if want_grad else None
            gD = torch.empty(B, p, l, dtype=torch.float64, device=dev) if want_grad else None
            ev_n.evaluate_device(d_k, d_p, N, gN, None, stream=stream)
            ev_d.evaluate_device(d_k, d_p, D, gD, None, stream=stream)
            # loss and quotient rule in one library kernel (gN becomes dloss), batch mean by the library's
            # deterministic two-stage reduction (nan_to_num -> mean | sum, mcts.py:345-347)
            lib = ev_n._lib
            sp = ctypes.c_void_p(stream.cuda_stream)
            loss = torch.empty(B, dtype=torch.float64, device=dev)
            rc = lib.qas_vqls_combine(N.data_ptr(), D.data_ptr(), gN.data_ptr() if want_grad else None,
                                      gD.data_ptr() if want_grad else None, B, p * l if want_grad else 0, cls.n_qubits,
                                      loss.data_ptr(), sp)
            if rc != 0:
                raise RuntimeError("qas_vqls_combine failed (%d)" % rc)
            out = {"energy": loss.cpu().numpy(), "grad_dense": None, "grad_compact": None}
            if want_grad:
                dense = torch.empty(p, c, l, dtype=torch.float64, device=dev)
                ev_n._check(lib.qas_batch_mean(ev_n._h, B, p, d_k.data_ptr(), gN.data_ptr(), dense.data_ptr(),
                                               0 if avg else _lib.QAS_F_GRAD_SUM, sp))
                out["grad_dense"] = dense.cpu().numpy()
                if compact:
                    out["grad_compact"] = gN.cpu().numpy()
        return out

    # ------------------------------------------------------------------ ModelFromK interface
    def _check(self, super_circ_params):
        assert super_circ_params.shape[0] == self.p
        assert super_circ_params.shape[1] == self.c
        assert super_circ_params.shape[2] == self.l

    def getLoss(self, super_circ_params: Union[np.ndarray, Sequence]):
        self._check(super_circ_params)
        return float(self.evaluate_batch([self.k], super_circ_params, self.pool, want_grad=False)["energy"][0])

    def getReward(self, super_circ_params: Union[np.ndarray, Sequence]):
        return float(np.exp(-10.0 * self.getLoss(super_circ_params)))

    def getGradient(self, super_circ_params: Union[np.ndarray, Sequence]):
        self._check(super_circ_params)
        gradients = np.zeros(np.shape(super_circ_params))
        if len(self.param_indices) == 0:
            return gradients
        gc = self.evaluate_batch([self.k], super_circ_params, self.pool, want_grad=True, compact=True)["grad_compact"][0]
        for (i, key, j) in self.param_indices:
            gradients[i, key, j] = gc[i, j]
        return gradients

    def toList(self, super_circ_params):
        gate_list = []
        for depth, key in enumerate(self.k):
            (name, wires), = self.pool[key].items()
            if name == "PlaceHolder":
                continue
            npar = SUPPORTED_OPS_DICT[name].num_params
            theta = [super_circ_params[depth, key, j] for j in range(npar)] if npar > 0 else None
            gate_list.append((name, wires, theta))
        return gate_list


class VQLSDemo(_VQLSModel):
    """A = I + 0.1 X_0 + 0.1 X_1 + 0.2 Z_2 Z_3, |b> = H^(x)4 |0>   (:1821-1831, :1838-1856)."""
    name = "VQLSDemo_4Q"
    n_qubits = 4
    problem = [(1.0, "IIII"), (0.1, "XIII"), (0.1, "IXII"), (0.2, "IIZZ")]


class VQLSDemo5Q(_VQLSModel):
    """A = I + 0.1 X_0 + 0.1 X_1 + 0.2 Z_2 Z_3 + 0.1 Z_3 Z_4, |b> = H^(x)5 |0>   (:2078-2092, :2099-2121)."""
    name = "VQLSDemo_5Q"
    n_qubits = 5
    problem = [(1.0, "IIIII"), (0.1, "XIIII"), (0.1, "IXIII"), (0.2, "IIZZI"), (0.1, "IIIZZ")]
