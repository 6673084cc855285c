"""Candidate-batch sharding across GPUs (one process per GPU, torch.distributed).

The batch of candidate circuits an epoch of ``search`` evaluates (reference
qas/mcts/mcts.py:340-347: a serial loop over ``batch_models``) consists of independent units that
share one parameter super-set, so it shards by circuit with no data-path collective.  The only
exchange is the one the search loop itself needs afterwards:

  * ``all_gather`` of the B rewards (so that every rank can back-propagate them into its tree),
  * ``all_reduce(sum)`` of the dense ``(p, c, l)`` gradient, then ``/ B`` -- exactly
    ``np.mean(np.nan_to_num(np.stack(grads)), axis=0)`` (``avg_gradients=False`` keeps the sum).

Works with the ``nccl`` backend on CUDA tensors (production) and with ``gloo`` on CPU tensors
(the world_size-2 tests, where the per-shard evaluation is supplied by the test).
"""
import numpy as np


def circuit_cost(ks, op_pool):
    """Work proxy per circuit: number of non-PlaceHolder one- and two-wire gates that survive
    lowering as tape operations (CNOT/SWAP are folded away by the tape compiler)."""
    src = getattr(op_pool, "pool", op_pool)
    free = np.array([1 if next(iter(src[k])) in ("PlaceHolder", "CNOT", "SWAP") else 0 for k in range(len(src))])
    ks = np.asarray(ks)
    return (1 - free)[ks].sum(axis=1)


def shard_plan(ks, op_pool, world_size):
    """-> list of index arrays, one per rank: circuits sorted by cost and dealt round-robin, so
    every rank gets the same mix of long and short tapes (shards differ by at most one circuit)."""
    order = np.argsort(-circuit_cost(ks, op_pool), kind="stable")
    return [order[r::world_size] for r in range(world_size)]


def evaluate_sharded(evaluate_fn, ks, params, op_pool, want_grad=True, avg=True, group=None, device=None):
    """Evaluate the GLOBAL batch `ks` (identical on every rank) with each rank computing its shard.

    evaluate_fn(ks_local, params, avg=False) -> dict(energy[B_local], grad_dense[p,c,l] SUM or None)
    Returns dict(energy[B] in the original order, grad_dense[p,c,l] mean|sum over the global batch).
    """
    import torch
    import torch.distributed as dist

    ks = np.asarray(ks, dtype=np.int32)
    B = ks.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plan = shard_plan(ks, op_pool, world)
    mine = plan[rank]
    out = evaluate_fn(ks[mine], params, avg=False) if len(mine) else {"energy": np.zeros(0), "grad_dense": None}
    dev = device if device is not None else "cpu"
    # rewards: pad every shard to the same length, all_gather, undo the permutation
    width = max(len(ix) for ix in plan)
    e_local = torch.zeros(width, dtype=torch.float64, device=dev)
    e_local[:len(mine)] = torch.as_tensor(np.asarray(out["energy"]), dtype=torch.float64, device=dev)
    energy = np.empty(B, dtype=np.float64)
    if world > 1:
        gathered = torch.empty(world * width, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, e_local, group=group)
        g = gathered.cpu().numpy().reshape(world, width)
    else:
        g = e_local.cpu().numpy().reshape(1, width)
    for r, ix in enumerate(plan):
        energy[ix] = g[r, :len(ix)]
    grad = None
    if want_grad:
        p, c, l = np.shape(params)
        gsum = torch.zeros((p, c, l), dtype=torch.float64, device=dev)
        if out.get("grad_dense") is not None:
            gsum += torch.as_tensor(np.asarray(out["grad_dense"]), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(gsum, op=dist.ReduceOp.SUM, group=group)
        grad = gsum.cpu().numpy()
        if avg:
            grad = grad / B
    return {"energy": energy, "grad_dense": grad}
