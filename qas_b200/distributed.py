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


class PeerExchange:
    """The exchange step of a sharded batch over NVLink peer memory (csrc/exchange.cu; no NCCL on the data path).

    Every rank owns a receive buffer of two alternating sets of ``world`` rows plus ``world`` flags in torch
    symmetric memory, mapped into every peer.  ``push(row)`` stores the local packed row
    ``[energies | gradient sum]`` into this rank's slot on every peer (one kernel, plain stores over NVLink) and
    publishes the step number; ``wait_reduce`` waits on the local flags and reduces the gradient parts of the
    gathered rows in the fixed order of ``qas_reduce_rows``.  Contract (include/qas_b200.h): every rank calls
    push, wait_reduce, push, ... on ONE stream.  Construction is a collective over ``group``.
    """

    def __init__(self, row_doubles, device, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from qas_b200 import _lib
        self._lib = _lib.load()
        self._ct = ctypes
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.row = (int(row_doubles) + 1) & ~1            # 16-byte stores: an even number of doubles per row
        self.device = torch.device(device)
        nflag = 16
        self.buf = symm.empty(2 * self.world * self.row + nflag, dtype=torch.float64, device=self.device)
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        self.hdl = symm.rendezvous(self.buf, group)
        dist.barrier(group)                                # every rank's flags are zero before anybody pushes
        base = [int(x) for x in self.hdl.buffer_ptrs]
        vp = ctypes.c_void_p
        flag_off = 8 * 2 * self.world * self.row
        self._peer_rows = [(vp * self.world)(*[base[r] + 8 * (b * self.world + self.rank) * self.row for r in range(self.world)])
                           for b in (0, 1)]
        self._peer_flags = (vp * self.world)(*[base[r] + flag_off + 8 * self.rank for r in range(self.world)])
        self._local = base[self.rank]
        self._local_flags = base[self.rank] + flag_off
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.seq = 0

    def push(self, row, stream):
        """row: contiguous float64 CUDA tensor of exactly ``self.row`` doubles (the caller's packed row)."""
        assert row.is_cuda and row.is_contiguous() and row.numel() == self.row
        self.seq += 1
        rc = self._lib.qas_peer_push_row(row.data_ptr(), self.row, self._peer_rows[self.seq & 1], self._peer_flags, self.world,
                                         self.seq, self._ct.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError("qas_peer_push_row failed: %s" % self._lib.qas_last_error(None).decode())

    def wait_reduce(self, offset, total, scale, out, stream):
        """waits for every rank's row of the last push; out[x] = scale * sum_r rows[r][offset + x], x < total"""
        rows = self._local + 8 * (self.seq & 1) * self.world * self.row
        rc = self._lib.qas_peer_wait_reduce(rows, self.world, self.row, int(offset), int(total), float(scale),
                                            out.data_ptr() if out is not None else None, self._local_flags, self.seq,
                                            self.status.data_ptr(), self._ct.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError("qas_peer_wait_reduce failed: %s" % self._lib.qas_last_error(None).decode())

    def rows(self):
        """view [world, row] of the rows gathered by the last push / wait_reduce (valid once the stream got there)"""
        b = self.seq & 1
        return self.buf[b * self.world * self.row:(b + 1) * self.world * self.row].view(self.world, self.row)

    def timed_out(self):
        return bool(int(self.status.item()))
