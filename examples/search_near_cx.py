"""The reference's experiment scripts search-scripts-legacy/search_LiH_near_cx.py and
search_H2O_near_cx.py (hyper-parameters at :64-136 / :66-138) with only their imports switched to the
B200 drop-in.  The penalty is the one the scripts define inline (PlaceholderPenalty), the optimiser
PennyLane's Adam rule.

    python examples/search_near_cx.py LiH        # 10 qubits, p = 20, 50 search epochs, 400 tuning steps
    python examples/search_near_cx.py H2O        #  8 qubits, p = 50, 50 search epochs, 100 tuning steps

Writes the result JSON in the reference's schema (task, pool, params, k, op_list, search_reward_list,
fine_tune_loss) and prints the wall time of the search and of the fine-tuning."""
import json
import sys
import time

import numpy as np

from qas_b200.mcts import QMLStateBasicGates, circuitModelTuning, search
from qas_b200.mcts.native import PlaceholderPenalty
from qas_b200.optim import AdamOptimizer
from qas_b200.qml_models.qml_gate_ops import QMLPool
from qas_b200.qml_models.qml_models_legacy import H2O, LiH

CONFIGS = {
    # model, qubits, p, search batch, tuning epochs, penalty scale, saved result of the reference, FCI quoted by the script
    "LiH": (LiH, 10, 20, 25, 400, 1.0, -7.952638648592746, -7.97241),
    "H2O": (H2O, 8, 50, 50, 100, 10.0, -75.42203598772306, -75.491788196432),
}


class NpEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def generate_near_cx_connection_list(num_qubits):
    out = []
    for i in range(num_qubits - 1):
        out.append([i, i + 1])
        out.append([i + 1, i])
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "LiH"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    model, n, p, batch, tune_epochs, pen_scale, ref_loss, e_fci = CONFIGS[which]
    np.random.seed(seed)
    state_class = QMLStateBasicGates
    task = model.name + "_" + state_class.name
    pool = QMLPool(n, ["Rot", "PlaceHolder"], ["CNOT"], complete_undirected_graph=False,
                   two_qubit_gate_map=generate_near_cx_connection_list(n))
    l, c = 3, len(pool)
    penalty_func = PlaceholderPenalty(pool, p, pen_scale)
    init_params = np.random.randn(p, c, l)
    t0 = time.time()
    final_params, final_best_arc, final_best_node, final_best_reward, final_controller, reward_list = search(
        model=model, op_pool=pool, target_circuit_depth=p, init_qubit_with_controls=set(), init_params=init_params,
        num_iterations=50, num_warmup_iterations=2, super_circ_train_optimizer=AdamOptimizer,
        super_circ_train_gradient_noise_factor=0.0, early_stop_threshold=999, early_stop_lookback_count=1,
        super_circ_train_lr=1, penalty_function=penalty_func, gate_limit_dict={"CNOT": p // 2},
        warmup_arc_batchsize=500, search_arc_batchsize=batch, alpha_max=2, alpha_decay_rate=0.99,
        prune_constant_max=0.9, prune_constant_min=0.8, max_visits_prune_threshold=10, min_num_children=5,
        sampling_execute_rounds=10, exploit_execute_rounds=20, cmab_sample_policy="local_optimal",
        cmab_exploit_policy="local_optimal", uct_sample_policy="local_optimal", verbose=0,
        state_class=state_class, search_reset=True)
    t1 = time.time()
    final_params, loss_list = circuitModelTuning(
        initial_params=final_params, model=model, num_epochs=tune_epochs, k=final_best_arc, op_pool=pool,
        opt_callable=AdamOptimizer, lr=0.1, grad_noise_factor=0, verbose=0, early_stop_threshold=-99999)
    t2 = time.time()
    res = {"task": task, "pool": pool.pool, "params": final_params, "k": final_best_arc,
           "op_list": model(p, c, l, final_best_arc, pool).toList(final_params),
           "search_reward_list": reward_list, "fine_tune_loss": loss_list}
    name = time.strftime("%Y%m%d-%H%M%S") + "_" + task + ".json"
    with open(name, "w") as f:
        json.dump(res, f, indent=4, cls=NpEncoder)
    print(json.dumps({"task": task, "seed": seed, "search_s": round(t1 - t0, 3), "fine_tune_s": round(t2 - t1, 3),
                      "final_loss": loss_list[-1], "reference_saved_result": ref_loss, "fci_quoted": e_fci,
                      "gates": len(res["op_list"]), "file": name}))
