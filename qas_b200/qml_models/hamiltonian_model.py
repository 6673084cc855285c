"""Pauli-sum energy model on the B200 evaluator (drop-in for the reference's
``qas/qml_models/hamiltonian_model.py`` ``HamiltonianModel``, :15-116).

Upstream every ``getLoss`` / ``getReward`` / ``getGradient`` call builds a PennyLane device
and QNode and runs one circuit on the CPU.  Here a subclass only declares WHAT is evaluated
(qubit count, Hamiltonian table, initial state); the arithmetic runs in libqas_b200's CUDA
kernels, one evaluator context per (class, pool, device) shared by all instances, and the
class method ``evaluate_batch`` evaluates many structure lists in one launch -- the batched
entry that replaces the serial loops at qas/mcts/mcts.py:298-305 and :340-347.
"""
import os
from typing import List, Sequence, Union

import numpy as np

from qas_b200 import hamiltonians
from qas_b200.evaluator import BatchEvaluator
from qas_b200.models import ModelFromK
from qas_b200.qml_models.qml_gate_ops import SUPPORTED_OPS_DICT, QMLPool, pool_table
from qas_b200.qml_models.utils import extractParamIndicesQML

_EVALUATORS = {}


def default_device():
    return int(os.environ.get("QAS_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


class HamiltonianModel(ModelFromK):
    name: str = "HamiltonianModel"
    num_qubits: int = None
    hamiltonian: str = None            # table name under qas_b200/data/
    init_state = ("zero", 0)           # ("zero", 0) | ("basis", index) | ("plus", 0)
    reward_sign = -1.0                 # reward = -loss (hamiltonian_model.py:73)

    def __init__(self, p: int, c: int, l: int, structure_list: List[int], op_pool: Union[QMLPool, dict]):
        self.k = structure_list
        self.pool = op_pool
        self.p, self.c, self.l = p, c, l
        self.param_indices = extractParamIndicesQML(self.k, self.pool)

    # ------------------------------------------------------------------ evaluator plumbing
    @classmethod
    def terms(cls):
        return hamiltonians.load(cls.hamiltonian)["terms"]

    @classmethod
    def evaluator(cls, op_pool, l=3, device=None):
        device = default_device() if device is None else device
        key = (cls, tuple(pool_table(op_pool)), l, device)
        ev = _EVALUATORS.get(key)
        if ev is None:
            ev = BatchEvaluator(cls.num_qubits, cls.terms(), op_pool, cls.init_state, l, device)
            _EVALUATORS[key] = ev
        return ev

    @classmethod
    def evaluate_batch(cls, ks, super_circ_params, op_pool, want_grad=True, avg=True, compact=False, out=None):
        """Energies (losses) [B] and the batch-mean (or sum) dense gradient [p, c, l] of a batch
        of structure lists sharing `super_circ_params` (mcts.py:340-347 in one launch).
        out: optional preallocated result arrays (BatchEvaluator.evaluate)."""
        l = np.asarray(super_circ_params).shape[2]
        return cls.evaluator(op_pool, l).evaluate(ks, super_circ_params, want_grad=want_grad,
                                                  dense=True, compact=compact, avg=avg, out=out)

    @classmethod
    def evaluate_batch_sharded(cls, ks, super_circ_params, op_pool, want_grad=True, avg=True, group=None):
        """Multi-GPU form of evaluate_batch: every rank passes the same global batch, evaluates its
        own shard on its own GPU and the rewards / gradient are combined over NCCL
        (qas_b200/distributed.py).  Needs an initialised torch.distributed process group."""
        import torch
        from qas_b200.distributed import evaluate_sharded
        dev = default_device()

        def local(ks_local, params, avg=False):
            return cls.evaluator(op_pool, np.shape(params)[2], dev).evaluate(
                ks_local, params, want_grad=want_grad, dense=True, avg=avg)

        return evaluate_sharded(local, ks, super_circ_params, op_pool, want_grad=want_grad, avg=avg,
                                group=group, device=torch.device("cuda", dev))

    def _check(self, super_circ_params):
        assert super_circ_params.shape[0] == self.p
        assert super_circ_params.shape[1] == self.c
        assert super_circ_params.shape[2] == self.l

    # ------------------------------------------------------------------ ModelFromK interface
    def getLoss(self, super_circ_params: Union[np.ndarray, Sequence]):
        self._check(super_circ_params)
        out = self.evaluator(self.pool, self.l).evaluate([self.k], super_circ_params, want_grad=False)
        return float(out["energy"][0])

    def getReward(self, super_circ_params: Union[np.ndarray, Sequence]):
        return self.reward_sign * self.getLoss(super_circ_params)

    def getGradient(self, super_circ_params: Union[np.ndarray, Sequence]):
        self._check(super_circ_params)
        gradients = np.zeros(np.shape(super_circ_params))
        if len(self.param_indices) == 0:
            return gradients
        out = self.evaluator(self.pool, self.l).evaluate([self.k], super_circ_params, want_grad=True,
                                                         dense=False, compact=True)
        gc = out["grad_compact"][0]
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
