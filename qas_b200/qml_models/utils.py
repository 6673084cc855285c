"""Index helpers shared by the model classes (reference qas/qml_models/utils.py:11-25)."""
from typing import List, Union

from qas_b200.qml_models.qml_gate_ops import SUPPORTED_OPS_DICT, QMLPool


def extractParamIndicesQML(k: List[int], op_pool: Union[QMLPool, dict]) -> List:
    """[(depth i, key k_i, slot j)] for every parameter of every parameterised gate in k --
    the gather/scatter map between the shared tensor params[p, c, l] and one circuit."""
    assert min(k) >= 0
    assert max(k) < len(op_pool)
    indices = []
    for depth, key in enumerate(k):
        entry = op_pool[key]
        assert len(entry) == 1
        (name, _), = entry.items()
        indices.extend((depth, key, slot) for slot in range(SUPPORTED_OPS_DICT[name].num_params))
    return indices
