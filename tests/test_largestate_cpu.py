"""CPU tier: host logic of the large-state mode (gate lowering, wire map, dependency-aware
scheduling, rank-bit swaps) against the oracle, single process and 2/4 gloo ranks."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers
from oracle import sim
from qas_b200.largestate import LargeStateSimulator, hea_gate_list, lower_gate, zz_chain_terms
from sv_numpy_backend import NumpySvBackend


def _random_gate_list(n, ngates, seed):
    rng = np.random.default_rng(seed)
    pool = helpers.full_vocab_pool(n)
    gl = []
    for _ in range(ngates):
        (name, wires), = pool[int(rng.integers(0, len(pool)))].items()
        if name == "PlaceHolder":
            continue
        from oracle import gates as G
        npar = G.num_params(name)
        gl.append((name, list(wires), list(rng.standard_normal(npar)) if npar else None))
    return gl


def _terms(n, seed):
    rng = np.random.default_rng(seed)
    # X/Y only on the low wires' far end so that a sharded run can localise them with one swap
    out = zz_chain_terms(n)
    for _ in range(4):
        w = ["I"] * n
        for q in rng.choice(np.arange(n // 2, n), size=2, replace=False):
            w[q] = str(rng.choice(list("XYZ")))
        out.append((float(rng.standard_normal()), "".join(w)))
    return out


def test_lowering_matches_oracle_matrices():
    from oracle import gates as G
    rng = np.random.default_rng(0)
    for name, (nw, npar, fn) in G.GATES.items():
        if name == "PlaceHolder":
            continue
        th = list(rng.standard_normal(npar))
        wires = [0] if nw == 1 else [1, 0]
        n = 2
        st0 = rng.standard_normal(4) + 1j * rng.standard_normal(4)
        st0 /= np.linalg.norm(st0)
        ref = sim.apply_matrix(st0.reshape(2, 2), G.matrix(name, th), wires).reshape(-1)
        s = LargeStateSimulator(n, backend=NumpySvBackend())
        NumpySvBackend._view(s.state)[:] = st0
        s.run([(name, wires, th if npar else None)])
        assert np.allclose(NumpySvBackend._view(s.state), ref, atol=1e-13), name


def test_single_process_matches_oracle():
    n = 6
    gl = _random_gate_list(n, 40, seed=1)
    terms = _terms(n, 1)
    for init in (("zero", 0), ("plus", 0), ("basis", 37)):
        s = LargeStateSimulator(n, backend=NumpySvBackend()).reset(init)
        e = s.run(gl).expval(terms)
        ref = sim.expval(sim.run(n, gl, init), terms)
        assert abs(e - ref) < 1e-12


def _worker(rank, world, port, n, gl, terms, init, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = LargeStateSimulator(n, backend=NumpySvBackend()).reset(init)
    e = s.run(gl).expval(terms)
    if rank == 0:
        q.put((e, s.swaps))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_gloo_matches_oracle(world):
    n = 7
    gl = hea_gate_list(n, 2, seed=3) + _random_gate_list(n, 25, seed=4)
    terms = _terms(n, 2)
    init = ("basis", 0b1010011)
    ref = sim.expval(sim.run(n, gl, init), terms)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, gl, terms, init, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    e, swaps = q.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert abs(e - ref) < 1e-12
    assert swaps >= 1


def test_lower_gate_placeholder_and_unknown():
    assert lower_gate("PlaceHolder", [0], None) == []
    with pytest.raises(ValueError):
        lower_gate("Toffoli", [0, 1, 2], None)


def _plan(n_local, gates, tile_bits):
    import ctypes
    from qas_b200 import _lib
    lib = _lib.load()
    arr = (_lib.SvGate * len(gates))()
    for i, (kind, target, control, mask) in enumerate(gates):
        arr[i].kind, arr[i].target, arr[i].control, arr[i].mask = kind, target, control, mask
    blk = (ctypes.c_int * len(gates))()
    nb = ctypes.c_int(0)
    bits = (ctypes.c_uint64 * len(gates))()
    rc = lib.qas_sv_plan(n_local, arr, len(gates), tile_bits, blk, ctypes.byref(nb), bits, len(gates))
    assert rc == 0, lib.qas_sv_last_error()
    return list(blk), nb.value, list(bits)[:nb.value]


@pytest.mark.parametrize("n_local,tile_bits", [(6, 3), (6, 1), (10, 6), (20, 12), (27, 12), (12, 13)])
def test_pass_planner_covers_every_gate_and_respects_dependencies(n_local, tile_bits):
    rng = np.random.default_rng(n_local * 31 + tile_bits)
    gates = []
    for _ in range(200):
        if rng.random() < 0.15:
            m = int(rng.integers(1, 1 << min(n_local, 30)))
            gates.append((1, 0, -1, m))
        else:
            t = int(rng.integers(0, n_local))
            c = int(rng.integers(0, n_local)) if rng.random() < 0.4 else -1
            if c == t:
                c = -1
            gates.append((0, t, c, 0))
    blk, nb, bits = _plan(n_local, gates, tile_bits)
    kb = min(tile_bits, n_local, 13)
    assert nb >= 1 and all(0 <= b < nb for b in blk)
    for b in bits:
        assert bin(b).count("1") == kb
    for i, (kind, t, c, m) in enumerate(gates):
        if kind == 0:
            assert (bits[blk[i]] >> t) & 1          # the target is a tile bit of its pass
    # two gates that share a bit never run in the wrong order
    def touched(g):
        kind, t, c, m = g
        return ({t} | ({c} if c >= 0 else set())) if kind == 0 else {b for b in range(n_local) if (m >> b) & 1}
    for i in range(len(gates)):
        for j in range(i + 1, min(len(gates), i + 40)):
            if touched(gates[i]) & touched(gates[j]):
                assert blk[i] <= blk[j]


def test_pass_planner_blocks_hea_layer_into_few_passes():
    n = 27
    gates = [(0, n - 1 - w, -1, 0) for w in range(n)] + [(0, n - 2 - w, n - 1 - w, 0) for w in range(n - 1)]
    blk, nb, _ = _plan(n, gates * 4, 12)
    assert nb <= 4 * 9      # <= 9 HBM passes per layer of 53 gates
