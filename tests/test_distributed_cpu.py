"""CPU tier: the N>1 host path (shard plan + gather/reduce) on the gloo backend, world_size 2.
The per-shard evaluation is the oracle (test infrastructure); the product path plugs in the CUDA
evaluator instead (qas_b200.qml_models.hamiltonian_model.HamiltonianModel.evaluate_batch)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import helpers
from oracle import sim
from qas_b200.distributed import circuit_cost, evaluate_sharded, shard_plan
from qas_b200.sampling import sample_structure_lists


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_eval_fn(n, pool, terms, init):
    def fn(ks, params, avg=False):
        es, gs = [], np.zeros(np.shape(params))
        for k in ks:
            k = [int(x) for x in k]
            es.append(sim.energy(n, k, pool, params, terms, init))
            gs += np.nan_to_num(sim.grad_adjoint(n, k, pool, params, terms, init))
        return {"energy": np.array(es), "grad_dense": gs}
    return fn


def _worker(rank, world, port, ks, params, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qas_b200.qml_models.qml_models_legacy import FourQubitH2
    pool = helpers.near_cx_pool(4)
    fn = _oracle_eval_fn(4, pool.pool, FourQubitH2.terms(), FourQubitH2.init_state)
    out = evaluate_sharded(fn, ks, params, pool)
    if rank == 0:
        q.put((out["energy"], out["grad_dense"]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_plan_is_a_balanced_partition():
    pool = helpers.near_cx_pool(6)
    ks = sample_structure_lists(pool, 12, {"CNOT": 6}, 37, seed=0)
    for world in (1, 2, 4, 8):
        plan = shard_plan(ks, pool, world)
        allidx = np.sort(np.concatenate(plan))
        assert np.array_equal(allidx, np.arange(37))
        sizes = [len(ix) for ix in plan]
        assert max(sizes) - min(sizes) <= 1
        cost = circuit_cost(ks, pool)
        loads = [cost[ix].sum() for ix in plan]
        assert max(loads) - min(loads) <= cost.max() + 1


def test_two_rank_gloo_matches_single_process():
    from qas_b200.qml_models.qml_models_legacy import FourQubitH2
    pool = helpers.near_cx_pool(4)
    p, B = 8, 11                                  # odd batch: ragged shards
    ks = sample_structure_lists(pool, p, {"CNOT": 4}, B, seed=5)
    params = np.random.default_rng(5).standard_normal((p, len(pool), 3))
    fn = _oracle_eval_fn(4, pool.pool, FourQubitH2.terms(), FourQubitH2.init_state)
    ref = evaluate_sharded(fn, ks, params, pool)   # no process group: world 1
    ref_e = np.array([sim.energy(4, [int(x) for x in k], pool.pool, params, FourQubitH2.terms(),
                                 FourQubitH2.init_state) for k in ks])
    assert np.allclose(ref["energy"], ref_e, atol=1e-14)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ks, params, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    energy, grad = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert np.max(np.abs(energy - ref["energy"])) <= 1e-14
    assert np.max(np.abs(grad - ref["grad_dense"])) <= 1e-14
