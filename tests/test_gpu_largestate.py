"""GPU tier: large-state mode (HBM-resident statevector, tiled gate passes, Pauli-sum expectation,
NCCL rank-bit swaps) against the CPU oracle."""
import numpy as np
import pytest

import helpers
from oracle import sim
from qas_b200.largestate import CudaSvBackend, LargeStateSimulator, hea_gate_list, zz_chain_terms

pytestmark = pytest.mark.gpu


def _random_gate_list(n, ngates, seed):
    from oracle import gates as G
    rng = np.random.default_rng(seed)
    pool = helpers.full_vocab_pool(n)
    gl = []
    for _ in range(ngates):
        (name, wires), = pool[int(rng.integers(0, len(pool)))].items()
        if name == "PlaceHolder":
            continue
        npar = G.num_params(name)
        gl.append((name, list(wires), list(rng.standard_normal(npar)) if npar else None))
    return gl


def _state(s):
    return s.state.cpu().numpy().view(np.complex128).reshape(-1)


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("n,tile_bits", [(6, 3), (9, 6), (12, 8), (14, 12), (15, 13), (17, 9), (18, 12)])
def test_full_vocabulary_statevector_matches_oracle(n, tile_bits, tma, monkeypatch):
    """Tiles moved by TMA (cp.async.bulk.tensor + mbarrier, 128-byte swizzle; default wherever a 5-d tensor map can
    describe the tile) and by cp.async (QAS_B200_SV_TMA=0, and tiles below 128 amplitudes) give the oracle's state."""
    monkeypatch.setenv("QAS_B200_SV_TMA", tma)
    gl = _random_gate_list(n, 60, seed=n)
    rng = np.random.default_rng(n)
    terms = zz_chain_terms(n) + [(float(rng.standard_normal()), "".join(rng.choice(list("IXYZ"), size=n)))
                                 for _ in range(6)]
    for init in (("zero", 0), ("plus", 0), ("basis", (1 << n) - 3)):
        s = LargeStateSimulator(n, backend=CudaSvBackend(0, tile_bits)).reset(init)
        s.run(gl)
        ref = sim.run(n, gl, init).reshape(-1)
        assert np.max(np.abs(_state(s) - ref)) < 1e-12
        assert abs(s.expval(terms) - sim.expval(ref.reshape([2] * n), terms)) < 1e-10
        assert (s.backend.tma_passes > 0) == (tma == "1" and tile_bits >= 7)


def test_tma_and_cp_async_tiles_agree_on_hea_22q(monkeypatch):
    """Same HEA circuit through both tile-movement paths: identical energies, and the TMA path really ran
    (every pass of this circuit has a describable tile geometry)."""
    n = 22
    gl = hea_gate_list(n, 3, seed=1)
    out = {}
    for tma, pipe in (("1", "1"), ("1", "0"), ("0", "0")):       # double-buffered TMA pipeline, single-buffered TMA, cp.async
        monkeypatch.setenv("QAS_B200_SV_TMA", tma)
        monkeypatch.setenv("QAS_B200_SV_PIPE", pipe)
        s = LargeStateSimulator(n).run(gl)
        out[tma + pipe] = (s.expval(zz_chain_terms(n)), s.backend.tma_passes, s.backend.passes, _state(s)[:4096].copy())
    for key in ("10", "00"):
        assert abs(out["11"][0] - out[key][0]) < 1e-11
        assert np.max(np.abs(out["11"][3] - out[key][3])) < 1e-13
    assert out["11"][1] == out["11"][2] and out["10"][1] == out["10"][2] and out["00"][1] == 0


def test_hea_20q_energy_matches_oracle():
    n = 20
    gl = hea_gate_list(n, 2, seed=0)
    s = LargeStateSimulator(n).run(gl)
    ref = sim.run_flat(n, gl)
    assert abs(s.expval(zz_chain_terms(n)) - sim.expval_diag_zz_chain(ref, n)) < 1e-9
    assert np.max(np.abs(_state(s) - ref)) < 1e-12


def test_hea_24q_checked_against_cpu():
    """BASELINE config 5's check size: 24 qubits, L=4 hardware-efficient ansatz, sum Z_i Z_{i+1}."""
    n = 24
    gl = hea_gate_list(n, 4, seed=0)
    s = LargeStateSimulator(n)
    s.run(gl)
    e = s.expval(zz_chain_terms(n))
    ref = sim.run_flat(n, gl)
    assert abs(e - sim.expval_diag_zz_chain(ref, n)) < 1e-9
    # spot-check amplitudes without copying 256 MB back twice
    got = _state(s)
    probe = np.random.default_rng(0).integers(0, 1 << n, size=4096)
    assert np.max(np.abs(got[probe] - ref[probe])) < 1e-12
    # norm preserved, tiling does not change the result
    assert abs(s.expval([(1.0, "I" * n)]) - 1.0) < 1e-12
    s2 = LargeStateSimulator(n, backend=CudaSvBackend(0, 10)).run(gl)
    assert abs(s2.expval(zz_chain_terms(n)) - e) < 1e-10
    assert s.backend.passes < len(gl) / 4      # gates are blocked into far fewer HBM passes


def _sv_worker(rank, world, port, n, gl, terms, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s = LargeStateSimulator(n, device=rank)
    e = s.run(gl).expval(terms)
    if rank == 0:
        q.put((e, s.swaps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_state_matches_single_gpu():
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = 22
    gl = hea_gate_list(n, 3, seed=1)
    terms = zz_chain_terms(n)
    e1 = LargeStateSimulator(n).run(gl).expval(terms)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sv_worker, args=(r, 2, port, n, gl, terms, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    e2, swaps = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert abs(e1 - e2) < 1e-10 and swaps >= 1
