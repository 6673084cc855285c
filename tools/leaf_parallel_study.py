"""Leaf-parallel rollouts against sequential rollouts ("virtual loss is not
semantics-preserving -> opt-in and reported separately").

Whole LiH / H2O searches with the reference scripts' hyper-parameters (search_LiH_near_cx.py:64-136,
search_H2O_near_cx.py:66-138: 2 warm-up + 48 search epochs, then fine-tuning of the best arc) for a number of
seeds, once with the reference's sequential selection rounds and once with `leaf_parallel` rounds per wave under a
virtual loss.  Reports, per mode, the search wall time and the distribution of the final fine-tuned energies.

    python tools/leaf_parallel_study.py LiH --seeds 20 --leaves 8
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))


def one_search(which, seed, leaves, tune_device):
    import random
    from search_near_cx import CONFIGS, generate_near_cx_connection_list
    from qas_b200.mcts import QMLStateBasicGates, circuitModelTuning, search
    from qas_b200.mcts.driver import circuitModelTuningDevice
    from qas_b200.mcts.native import PlaceholderPenalty
    from qas_b200.optim import AdamOptimizer
    from qas_b200.qml_models.qml_gate_ops import QMLPool
    model, n, p, batch, tune_epochs, pen_scale, ref_loss, e_fci = CONFIGS[which]
    np.random.seed(seed)
    random.seed(seed)
    pool = QMLPool(n, ["Rot", "PlaceHolder"], ["CNOT"], complete_undirected_graph=False,
                   two_qubit_gate_map=generate_near_cx_connection_list(n))
    c = len(pool)
    t0 = time.time()
    params, arc, node, reward, controller, history = search(
        model=model, op_pool=pool, target_circuit_depth=p, init_qubit_with_controls=set(), init_params=np.random.randn(p, c, 3),
        num_iterations=50, num_warmup_iterations=2, super_circ_train_optimizer=AdamOptimizer,
        super_circ_train_gradient_noise_factor=0.0, early_stop_threshold=999, early_stop_lookback_count=1,
        super_circ_train_lr=1, penalty_function=PlaceholderPenalty(pool, p, pen_scale), gate_limit_dict={"CNOT": p // 2},
        warmup_arc_batchsize=500, search_arc_batchsize=batch, alpha_max=2, alpha_decay_rate=0.99,
        prune_constant_max=0.9, prune_constant_min=0.8, max_visits_prune_threshold=10, min_num_children=5,
        sampling_execute_rounds=10, exploit_execute_rounds=20, verbose=0, state_class=QMLStateBasicGates, search_reset=True,
        leaf_parallel=leaves)
    t1 = time.time()
    if tune_device:
        _, losses = circuitModelTuningDevice(params, model, tune_epochs, arc, pool, lr=0.1, early_stop_threshold=-99999, check_every=50)
    else:
        _, losses = circuitModelTuning(initial_params=params, model=model, num_epochs=tune_epochs, k=arc, op_pool=pool,
                                       opt_callable=AdamOptimizer, lr=0.1, grad_noise_factor=0, verbose=0, early_stop_threshold=-99999)
    t2 = time.time()
    st = controller.stats() if hasattr(controller, "stats") else {}
    return {"seed": seed, "search_s": t1 - t0, "tune_s": t2 - t1, "final_loss": float(losses[-1]), "search_best_reward": float(reward),
            "reward_evaluations": int(st.get("reward_evaluations", 0)), "reward_batches": int(st.get("reward_batches", 0))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="LiH", choices=["LiH", "H2O"])
    ap.add_argument("--seeds", type=int, default=20)
    ap.add_argument("--leaves", type=int, default=8)
    ap.add_argument("--host-tuning", action="store_true", help="fine-tune with the host loop instead of the graph-captured device loop")
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from search_near_cx import CONFIGS
    out = {"task": args.which, "seeds": args.seeds, "reference_saved_result": CONFIGS[args.which][6], "fci_quoted": CONFIGS[args.which][7]}
    one_search(args.which, 10 ** 6, 1, not args.host_tuning)          # first-call costs
    for label, leaves in (("sequential", 1), ("leaf_parallel_%d" % args.leaves, args.leaves)):
        runs = [one_search(args.which, s, leaves, not args.host_tuning) for s in range(args.seeds)]
        fl = [r["final_loss"] for r in runs]
        out[label] = {"search_s_median": statistics.median(r["search_s"] for r in runs),
                      "search_s_total": sum(r["search_s"] for r in runs), "tune_s_median": statistics.median(r["tune_s"] for r in runs),
                      "final_loss_mean": statistics.mean(fl), "final_loss_median": statistics.median(fl),
                      "final_loss_best": min(fl), "final_loss_worst": max(fl), "final_loss_stdev": statistics.pstdev(fl),
                      "reward_evaluations_mean": statistics.mean(r["reward_evaluations"] for r in runs),
                      "reward_batches_mean": statistics.mean(r["reward_batches"] for r in runs), "runs": runs}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
