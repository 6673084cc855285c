"""``from qas_b200.mcts import search, TreeNode, circuitModelTuning, QMLStateBasicGates`` -- the
import line of the reference's experiment scripts (search-scripts-legacy/*.py:1,8)."""
from qas_b200.mcts.mcts import MCTSController, TreeNode, circuitModelTuning, search  # noqa: F401
from qas_b200.mcts.qml_mcts_state import QMLStateBasicGates, QMLStateBasicGatesNoRestrictions  # noqa: F401
