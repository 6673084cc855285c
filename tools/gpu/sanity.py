"""Tiny run of every kernel path for compute-sanitizer (memcheck / racecheck):
thread-per-circuit kernel (n=4), team kernel (n=8, n=10), global-memory path (n=13 with gradient),
large-state tile kernel (n=12).  Results are compared with the oracle so a silent corruption shows."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import __graft_entry__ as ge

ge.build()
import bench
from oracle import sim

for name, B in (("h2_4q", 96), ("h2o_8q", 24), ("lih_10q", 12)):
    model, pool, params, ks = bench.make_workload(name, B, 1)
    n = bench.WORKLOADS[name][1]
    out = model.evaluate_batch(ks, params, pool, want_grad=True, compact=True)
    terms = model.terms()
    for b in range(0, B, max(1, B // 4)):
        k = [int(x) for x in ks[b]]
        e = sim.energy(n, k, pool.pool, params, terms, model.init_state)
        g = sim.grad_adjoint(n, k, pool.pool, params, terms, model.init_state)
        assert abs(out["energy"][b] - e) < 1e-9, (name, b)
        for i, key in enumerate(k):
            assert np.max(np.abs(out["grad_compact"][b, i] - g[i, key])) < 1e-8, (name, b, i)
    print("[sanity] %s ok (%d circuits)" % (name, B), flush=True)

# global-memory instantiation (n = 13 with gradients)
from qas_b200.evaluator import BatchEvaluator
from qas_b200.qml_models.qml_gate_ops import QMLPool
from qas_b200.sampling import sample_structure_lists
n, p = 13, 8
pool = QMLPool(n, ["Rot", "PlaceHolder"], ["CNOT"], False, bench._line(n))
terms = [(0.5, "Z" + "I" * (n - 1)), (0.25, "XX" + "I" * (n - 2)), (-0.3, "I" * (n - 2) + "YY"), (0.2, "I" * 5 + "ZZ" + "I" * (n - 7))]
ev = BatchEvaluator(n, terms, pool)
params = np.random.default_rng(2).standard_normal((p, len(pool), 3))
ks = sample_structure_lists(pool, p, {"CNOT": 4}, 4, seed=2)
out = ev.evaluate(ks, params, want_grad=True, compact=True)
for b in range(2):
    k = [int(x) for x in ks[b]]
    assert abs(out["energy"][b] - sim.energy(n, k, pool.pool, params, terms, ("zero", 0))) < 1e-9
print("[sanity] global-memory path ok", flush=True)

# large-state tile kernel
from qas_b200.largestate import LargeStateSimulator, hea_gate_list, zz_chain_terms
s = LargeStateSimulator(12, device=0, tile_bits=8)
s.reset()
s.run(hea_gate_list(12, 2, seed=0))
print("[sanity] large-state energy", s.expval(zz_chain_terms(12)), flush=True)
print("[sanity] done", flush=True)
