"""Abstract MCTS state: a partial structure list plus legality bookkeeping
(reference qas/mcts/mcts_state.py:11-54)."""
from abc import ABC, abstractmethod


class StateOfMCTS(ABC):
    name: str

    @abstractmethod
    def getLegalActions(self, *args):
        raise NotImplementedError

    @abstractmethod
    def takeAction(self, *args):
        raise NotImplementedError

    @abstractmethod
    def isTerminal(self):
        raise NotImplementedError

    @abstractmethod
    def getReward(self):
        raise NotImplementedError

    @abstractmethod
    def getCurrK(self):
        raise NotImplementedError
