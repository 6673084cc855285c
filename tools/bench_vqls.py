"""VQLS cost as a second batched workload: losses + gradients of a batch of
candidate circuits through qas_b200.qml_models.vqls_model (two evaluator launches + the quotient rule on
the host), next to the reference's own formulation (2 L^2 (n + 1) Hadamard-test circuits per loss,
qml_models_legacy.py:1880-1951) restated in oracle/vqls.py and timed on one host core.

    python tools/bench_vqls.py [--model 4q|5q] [--batch 65536] [--steps 5]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="4q")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from oracle import vqls
    from qas_b200.qml_models.qml_gate_ops import QMLPool
    from qas_b200.qml_models.qml_models_legacy import VQLSDemo, VQLSDemo5Q
    from qas_b200.sampling import sample_structure_lists
    model = VQLSDemo if args.model == "4q" else VQLSDemo5Q
    n = model.n_qubits
    ring = [[i, (i + 1) % n] for i in range(n)] + [[(i + 1) % n, i] for i in range(n)]
    pool = QMLPool(n, ["Rot", "PlaceHolder"], ["CNOT"], False, ring)       # vqls_search.py: Rot / PlaceHolder / ring CNOTs
    p, c = 10, len(pool)
    params = np.random.default_rng(0).standard_normal((p, c, 3))
    ks = sample_structure_lists(pool, p, {"CNOT": p // 2}, args.batch, seed=0)
    for _ in range(2):
        model.evaluate_batch(ks, params, pool, want_grad=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = model.evaluate_batch(ks, params, pool, want_grad=True)
    dt = (time.perf_counter() - t0) / args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.evaluate_batch(ks, params, pool, want_grad=False)
    dt_r = (time.perf_counter() - t0) / args.steps
    # the reference formulation on one core: losses only (its gradient is reverse-mode through all circuits)
    m, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 10.0:
        ref = vqls.cost_loc(n, model.problem, [int(x) for x in ks[m]], pool.pool, params)
        assert abs(ref - out["energy"][m]) < 1e-9
        m += 1
    cpu = m / (time.perf_counter() - t0)
    print(json.dumps({"metric": "VQLS candidate-circuit evals/sec (loss + gradient, host API end to end)",
                      "value": args.batch / dt, "unit": "evals/s", "loss_only": args.batch / dt_r,
                      "config": {"workload": "vqls_" + args.model, "model": model.name, "n_qubits": n, "depth_p": p,
                                 "pool_c": c, "circuits_per_step": args.batch,
                                 "observables": {"numerator_terms": len(model.numerator_terms()),
                                                 "denominator_terms": len(model.denominator_terms())}},
                      "ms_per_step": 1e3 * dt,
                      "cpu_baseline": {"value": cpu, "unit": "losses/s", "cores": 1, "kind": "port",
                                       "sample": "%d circuits, loss only, %d Hadamard-test circuits on %d qubits each "
                                                 "(oracle/vqls.py), every one checked against the GPU loss to 1e-9"
                                                 % (m, 2 * len(model.problem) ** 2 * (n + 1), n + 1)}}))


if __name__ == "__main__":
    main()
