"""Import-compatible face of the search driver: the reference exposes everything from one module
(``from qas.mcts.mcts import search, TreeNode, MCTSController, circuitModelTuning``); here the
tree lives in ``tree.py`` and the epoch loops in ``driver.py``."""
from qas_b200.mcts.driver import (SearchConfig, batchGradient, circuitModelTuning,  # noqa: F401
                                  circuitModelTuningDevice, search)
from qas_b200.mcts.tree import MCTSController, TreeNode  # noqa: F401


def getGradientFromModel(model, params):
    return model.getGradient(params)


def getLossFromModel(model, params):
    return model.getLoss(params)


def getRewardFromModel(model, params):
    return model.getReward(params)
