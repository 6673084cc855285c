"""Abstract gate / pool interfaces of the plugin surface (reference qas/ops.py:16-38)."""
from abc import ABC, abstractmethod


class QuantumGate(ABC):
    """A concrete gate: constructed from (name, wires, params); ``getOp`` returns the backend object."""

    @abstractmethod
    def __init__(self, *args):
        raise NotImplementedError

    @abstractmethod
    def __str__(self):
        raise NotImplementedError

    @abstractmethod
    def getOp(self):
        raise NotImplementedError


class Pool(dict):
    """Operator pool: integer key -> {gate name: wires}."""

    @abstractmethod
    def __init__(self, *args):
        super().__init__()

    @abstractmethod
    def __getitem__(self, key):
        raise NotImplementedError

    @abstractmethod
    def __str__(self):
        raise NotImplementedError

    @abstractmethod
    def __len__(self):
        raise NotImplementedError
