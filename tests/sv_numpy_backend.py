"""NumPy stand-in for the slab kernels of the large-state mode (test infrastructure): lets the
CPU tier run qas_b200.largestate.LargeStateSimulator -- wire map, scheduler, rank-bit resolution,
all-to-all swaps -- over the gloo backend without a GPU."""
import numpy as np
import torch


def _parity(v):
    v = v.copy()
    out = np.zeros(v.shape, dtype=np.int64)
    while np.any(v):
        out ^= (v & 1)
        v >>= 1
    return out


class NumpySvBackend:
    def __init__(self):
        self.passes = 0

    def alloc(self, n_local):
        return torch.zeros((1 << n_local, 2), dtype=torch.float64)

    @staticmethod
    def _view(state):
        return state.numpy().view(np.complex128).reshape(-1)

    def init(self, state, n_local, index, amp):
        v = self._view(state)
        if index < 0:
            v[:] = amp
        else:
            v[:] = 0
            if index < v.size:
                v[index] = amp

    def apply(self, state, n_local, gates):
        v = self._view(state)
        idx = np.arange(v.size, dtype=np.int64)
        for kind, target, control, mask, m in gates:
            m = np.asarray(m, dtype=np.complex128).reshape(2, 2)
            if kind == 0:
                sel = (idx >> target) & 1 == 0
                if control >= 0:
                    sel &= (idx >> control) & 1 == 1
                i0 = idx[sel]
                i1 = i0 | (1 << target)
                a0, a1 = v[i0].copy(), v[i1].copy()
                v[i0] = m[0, 0] * a0 + m[0, 1] * a1
                v[i1] = m[1, 0] * a0 + m[1, 1] * a1
            else:
                v *= np.where(_parity(idx & mask) == 1, m[1, 1], m[0, 0])
        self.passes += 1

    def expval(self, state, n_local, n_total, rank_bits, masks):
        v = self._view(state)
        idx = np.arange(v.size, dtype=np.int64)
        e = 0.0
        for x, z, c in masks:
            ny = bin(x & z).count("1")
            j = idx ^ x
            full = j | (rank_bits << n_local)
            sign = np.where(_parity(full & z) == 1, -1.0, 1.0)
            e += c * np.real(np.vdot(v, (1j ** ny) * sign * v[j]))
        return float(e)

    def all_to_all(self, out, inp, group):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        ins = list(inp.view(-1).chunk(world))
        outs = list(out.view(-1).chunk(world))
        try:
            dist.all_to_all(outs, ins, group=group)
        except RuntimeError:        # gloo builds without alltoall: emulate with all_gather
            gathered = [torch.empty_like(inp.view(-1)) for _ in range(world)]
            dist.all_gather(gathered, inp.view(-1).contiguous(), group=group)
            rank = dist.get_rank(group)
            for r in range(world):
                outs[r].copy_(gathered[r].chunk(world)[rank])

    def all_reduce_sum(self, value, group):
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64)
        dist.all_reduce(t, group=group)
        return float(t.item())
