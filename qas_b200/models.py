"""The model plugin interface the search driver calls (reference qas/models.py:3-21).

A model is built per structure list ``k`` and answers loss / reward / gradient queries for
a shared parameter super-set ``params[p, c, l]``.
"""
from abc import ABC, abstractmethod


class ModelFromK(ABC):
    name: str

    @abstractmethod
    def __init__(self, p: int, c: int, l: int, structure_list, op_pool, *args, **kargs):
        raise NotImplementedError

    @abstractmethod
    def getLoss(self, super_circ_params):
        raise NotImplementedError

    @abstractmethod
    def getGradient(self, super_circ_params):
        raise NotImplementedError

    @abstractmethod
    def toList(self, super_circ_params):
        """[(op_name, op_qubits, op_params or None)] with PlaceHolders omitted."""
        raise NotImplementedError

    @abstractmethod
    def getReward(self, super_circ_params):
        raise NotImplementedError
