"""Abstract plugin surface shared by the model zoo, the pool and the search driver.

Three contracts, restating the reference's ``qas/models.py:3-21`` and ``qas/ops.py:16-38``:

``ModelFromK``   a candidate circuit built from a structure list ``k`` (p pool keys) that answers
                 loss / reward / gradient / op-list queries for a shared parameter tensor
                 ``params[p, c, l]``;
``QuantumGate``  a concrete gate record (name, wires, parameters) with ``getOp``;
``Pool``         the operator pool, a mapping pool key -> {gate name: wires}.
"""
import abc


class ModelFromK(abc.ABC):
    name: str

    @abc.abstractmethod
    def __init__(self, p: int, c: int, l: int, structure_list, op_pool, *args, **kargs):
        """p = depth, c = pool size, l = parameter slots per (depth, key)."""

    @abc.abstractmethod
    def getLoss(self, super_circ_params): ...

    @abc.abstractmethod
    def getReward(self, super_circ_params): ...

    @abc.abstractmethod
    def getGradient(self, super_circ_params):
        """Dense array shaped like the parameters, non-zero only at this circuit's (i, k_i, j)."""

    @abc.abstractmethod
    def toList(self, super_circ_params):
        """[(gate name, wires, parameters or None)] with PlaceHolders left out."""


class QuantumGate(abc.ABC):
    @abc.abstractmethod
    def __init__(self, *args): ...

    @abc.abstractmethod
    def getOp(self): ...

    @abc.abstractmethod
    def __str__(self): ...


class Pool(dict):
    def __init__(self, *args):
        dict.__init__(self)

    def __getitem__(self, key):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def __str__(self):
        raise NotImplementedError
