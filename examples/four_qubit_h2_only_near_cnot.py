"""The reference's experiment script search-scripts-legacy/four_qubit_h2_only_near_cnot.py with
only its imports switched to the B200 drop-in (and a shorter schedule by default so that it runs in
about a minute; pass --full for the paper's hyper-parameters)."""
import json
import sys
import time

import numpy as np

from qas_b200.mcts import QMLStateBasicGates, TreeNode, circuitModelTuning, search
from qas_b200.optim import AdamOptimizer
from qas_b200.qml_models.qml_gate_ops import QMLPool
from qas_b200.qml_models.qml_models_legacy import FourQubitH2


class NpEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


if __name__ == "__main__":
    full = "--full" in sys.argv
    model = FourQubitH2
    state_class = QMLStateBasicGates
    task = model.name + "_" + state_class.name
    pool = QMLPool(4, ["Rot", "PlaceHolder"], ["CNOT"], complete_undirected_graph=False,
                   two_qubit_gate_map=[[0, 1], [1, 0], [1, 2], [2, 1], [2, 3], [3, 2]])
    p, l, c = (40 if full else 12), 3, len(pool)
    ph_count_limit = p

    def penalty_func(r: float, node: TreeNode):
        ph_count = sum(1 for key in node.state.getCurrK() if "PlaceHolder" in pool[key])
        return r - (ph_count - ph_count_limit) if ph_count >= ph_count_limit else r

    init_params = np.random.randn(p, c, l)
    final_params, final_best_arc, final_best_node, final_best_reward, final_controller, reward_list = search(
        model=model, op_pool=pool, target_circuit_depth=p, init_qubit_with_controls=set(), init_params=init_params,
        num_iterations=200 if full else 12, num_warmup_iterations=20 if full else 3,
        super_circ_train_optimizer=AdamOptimizer, super_circ_train_gradient_noise_factor=0.0,
        early_stop_threshold=1.13, early_stop_lookback_count=5, super_circ_train_lr=1 if full else 0.2,
        penalty_function=penalty_func, gate_limit_dict={"CNOT": p},
        warmup_arc_batchsize=5000 if full else 500, search_arc_batchsize=100 if full else 20, alpha_max=2,
        alpha_decay_rate=0.99, prune_constant_max=0.99, prune_constant_min=0.80, max_visits_prune_threshold=20,
        min_num_children=c // 2 + 1, sampling_execute_rounds=10, exploit_execute_rounds=100 if full else 20,
        verbose=1, state_class=state_class, search_reset=True)
    final_params, loss_list = circuitModelTuning(
        initial_params=init_params, model=model, num_epochs=400 if full else 150, k=final_best_arc, op_pool=pool,
        opt_callable=AdamOptimizer, lr=0.1, grad_noise_factor=0, verbose=0, early_stop_threshold=-1.2)
    res = {"task": task, "pool": pool.pool, "params": final_params, "k": final_best_arc,
           "op_list": model(p, c, l, final_best_arc, pool).toList(final_params),
           "search_reward_list": reward_list, "fine_tune_loss": loss_list}
    name = time.strftime("%Y%m%d-%H%M%S") + "_" + task + ".json"
    with open(name, "w") as f:
        json.dump(res, f, indent=4, cls=NpEncoder)
    print("final loss", loss_list[-1], "(FCI -1.136189454088) ->", name)
