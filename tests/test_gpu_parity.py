"""GPU tier: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances are the north_star's: |dE| <= 1e-9 Ha, gradient components <= 1e-8 (floating point,
complex128 arithmetic on both sides).  Small cases compare against the oracle on the same seeded
inputs and against the committed golden fixtures; BASELINE-size batches are checked through
size-independent properties (spectral bounds, linearity in the Hamiltonian, the parameter-shift
identity between the energy path and the adjoint path, permutation invariance, determinism).
"""
import numpy as np
import pytest

import helpers
from oracle import sim
from qas_b200.evaluator import BatchEvaluator
from qas_b200.qml_models.qml_gate_ops import QMLPool
from qas_b200.sampling import sample_structure_lists

pytestmark = pytest.mark.gpu

E_TOL = 1e-9
G_TOL = 1e-8


def _oracle_batch(n, ks, pool, params, terms, init, grad=True):
    es, gs = [], []
    for k in ks:
        k = [int(x) for x in k]
        es.append(sim.energy(n, k, pool, params, terms, init))
        if grad:
            gs.append(sim.grad_adjoint(n, k, pool, params, terms, init))
    return np.array(es), gs


@pytest.mark.parametrize("cfg", list(helpers.CONFIGS))
def test_config_batch_matches_oracle(cfg):
    mname, n, p, mk, lim = helpers.CONFIGS[cfg]
    model = helpers.get_model(mname)
    pool = mk()
    c = len(pool)
    B = 24 if n <= 8 else 6
    scale = np.sqrt(2.0 / 2 ** n) if "qaoa" in cfg else 1.0      # qaoa_search_demo.py:55
    params = np.random.default_rng(1).standard_normal((p, c, 3)) * scale
    ks = sample_structure_lists(pool, p, lim, B, seed=2)
    out = model.evaluate_batch(ks, params, pool, want_grad=True, compact=True)
    e_ref, g_ref = _oracle_batch(n, ks, pool.pool, params, model.terms(), model.init_state)
    assert np.max(np.abs(out["energy"] - e_ref)) <= E_TOL
    dense = helpers.dense_from_compact(ks, out["grad_compact"], c)
    for b in range(B):
        assert np.max(np.abs(dense[b] - g_ref[b])) <= G_TOL
    assert np.max(np.abs(out["grad_dense"] - sim.batch_mean_gradient(g_ref))) <= G_TOL
    # energy-only launch (the reward path) gives the same energies
    out2 = model.evaluate_batch(ks, params, pool, want_grad=False)
    assert np.max(np.abs(out2["energy"] - e_ref)) <= E_TOL
    assert out2["grad_dense"] is None


@pytest.mark.parametrize("g", helpers.goldens(), ids=lambda g: g["fixture"])
def test_golden_through_model_interface(g):
    """getLoss / getReward / getGradient / toList of the drop-in classes on the reference's saved
    circuits: energy = oracle = the reference's recorded loss (to the one-step lag)."""
    ev = BatchEvaluator(g["n_qubits"], g["terms"], g["pool"], g["init"])
    out = ev.evaluate([g["k"]], g["params"], want_grad=True, compact=True)
    e = out["energy"][0]
    assert abs(e - sim.energy(g["n_qubits"], g["k"], g["pool"], g["params"], g["terms"], g["init"])) <= E_TOL
    if g["tolerance_vs_last_loss"] is not None:
        assert abs(e - g["fine_tune_loss_tail"][-1]) <= g["tolerance_vs_last_loss"]
    ga = sim.grad_adjoint(g["n_qubits"], g["k"], g["pool"], g["params"], g["terms"], g["init"])
    dense = helpers.dense_from_compact([g["k"]], out["grad_compact"], len(g["pool"]))[0]
    assert np.max(np.abs(dense - ga)) <= G_TOL
    ev.close()


def test_model_methods_mirror_reference_interface():
    from qas_b200.qml_models.qml_models_legacy import FourQubitH2, QAOAVQCDemo
    g = [x for x in helpers.goldens() if x["fixture"] == "h2_hf_a"][0]
    pool = helpers.near_cx_pool(4)
    assert pool.pool == g["pool"]
    p, c, l = g["params"].shape
    m = FourQubitH2(p, c, l, g["k"], pool)
    loss = m.getLoss(g["params"])
    assert isinstance(loss, float) and abs(loss - g["fine_tune_loss_tail"][-1]) < 5e-5
    assert m.getReward(g["params"]) == -loss
    grad = m.getGradient(g["params"])
    assert grad.shape == g["params"].shape and grad is not g["params"]
    ga = sim.grad_adjoint(4, g["k"], g["pool"], g["params"], m.terms(), m.init_state)
    assert np.max(np.abs(grad - ga)) <= G_TOL
    ops = m.toList(g["params"])
    assert [(a, list(b)) for a, b, _ in ops] == [(a, list(b)) for a, b, _ in g["op_list"]]
    with pytest.raises(AssertionError):
        m.getLoss(g["params"][:-1])
    # QAOA: reward = -loss = expected cut value
    q = [x for x in helpers.goldens() if x["fixture"] == "qaoa7q_a"][0]
    qpool = helpers.near_cx_pool(7, helpers.ring(7))
    assert qpool.pool == q["pool"]
    qm = QAOAVQCDemo(15, len(qpool), 3, q["k"], qpool)
    assert abs(qm.getReward(q["params"]) - 7.0) < 1e-8
    # a circuit with no trainable gate: zero gradient (hamiltonian_model.py:84-85)
    m0 = FourQubitH2(3, c, l, [0, 8, 1], pool)
    assert m0.getGradient(g["params"][:3]).shape == (3, c, l)


def test_full_vocabulary_random_circuits():
    rng = np.random.default_rng(3)
    for n in (3, 5, 6):
        pool = helpers.full_vocab_pool(n)
        c = len(pool)
        terms = [(float(rng.standard_normal()), "".join(rng.choice(list("IXYZ"), size=n))) for _ in range(12)]
        for init in (("zero", 0), ("basis", int(rng.integers(0, 2 ** n))), ("plus", 0)):
            ev = BatchEvaluator(n, terms, pool, init)
            p, B = 9, 10
            ks = rng.integers(0, c, size=(B, p)).astype(np.int32)
            params = rng.standard_normal((p, c, 3))
            out = ev.evaluate(ks, params, want_grad=True, compact=True)
            e_ref, g_ref = _oracle_batch(n, ks, pool.pool, params, terms, init)
            assert np.max(np.abs(out["energy"] - e_ref)) <= E_TOL
            dense = helpers.dense_from_compact(ks, out["grad_compact"], c)
            assert max(np.max(np.abs(dense[b] - g_ref[b])) for b in range(B)) <= G_TOL
            ev.close()


@pytest.mark.parametrize("compress", ["0", "1"])
@pytest.mark.parametrize("n,grad", [(9, True), (11, True), (12, True), (13, False), (13, True), (14, True), (15, True), (16, True)])
def test_qubit_count_sweep_including_global_memory_path(n, grad, compress, monkeypatch):
    """Classic kernels (QAS_B200_COMPRESS=0): 11-12 qubits: 256-thread teams; 13 energy-only still shared-memory
    resident; 13 with gradients and 14-16 fall to the global-memory path.  Default: these shallow circuits span
    fewer than n dimensions, so every one of them -- up to 16 qubits -- runs as a compressed register in shared
    memory (path 3)."""
    monkeypatch.setenv("QAS_B200_COMPRESS", compress)
    rng = np.random.default_rng(n)
    pool = helpers.near_cx_pool(n)
    c = len(pool)
    terms = [(float(rng.standard_normal()), "".join(rng.choice(list("IXYZ"), size=n, p=[.55, .15, .15, .15])))
             for _ in range(10)]
    p, B = 8, 3
    ev = BatchEvaluator(n, terms, pool, ("zero", 0))
    ks = sample_structure_lists(pool, p, {"CNOT": 4}, B, seed=n)
    params = rng.standard_normal((p, c, 3))
    out = ev.evaluate(ks, params, want_grad=grad, compact=grad)
    st = ev.stats()
    # 4 = tiled kernel (kernels_t.cuh: thread-block cluster per circuit, tile passes over the slab)
    assert st["path"] == (3 if compress == "1" else (0 if n <= (12 if grad else 13) else 4))
    e_ref, g_ref = _oracle_batch(n, ks, pool.pool, params, terms, ("zero", 0), grad)
    assert np.max(np.abs(out["energy"] - e_ref)) <= E_TOL
    if grad:
        dense = helpers.dense_from_compact(ks, out["grad_compact"], c)
        assert max(np.max(np.abs(dense[b] - g_ref[b])) for b in range(B)) <= G_TOL
    ev.close()


def test_edge_cases():
    from qas_b200.qml_models.qml_models_legacy import FourQubitH2
    pool = helpers.near_cx_pool(4)
    c = len(pool)
    params = np.random.default_rng(0).standard_normal((4, c, 3))
    terms = FourQubitH2.terms()
    # all-PlaceHolder-like circuit is impossible under legality, but the evaluator accepts any keys:
    # only PlaceHolders -> the initial state's energy, zero gradient
    out = FourQubitH2.evaluate_batch([[1, 3, 5, 7]], params, pool, compact=True)
    e0 = sim.energy(4, [1, 3, 5, 7], pool.pool, params, terms, FourQubitH2.init_state)
    assert abs(out["energy"][0] - e0) <= E_TOL and not out["grad_compact"].any()
    # only CNOTs: tape is empty after folding, state is a permuted basis state
    out = FourQubitH2.evaluate_batch([[8, 10, 12, 9]], params, pool, compact=True)
    e1 = sim.energy(4, [8, 10, 12, 9], pool.pool, params, terms, FourQubitH2.init_state)
    assert abs(out["energy"][0] - e1) <= E_TOL
    # out-of-range key -> AssertionError, like the reference's asserts (utils.py:12-15)
    with pytest.raises(AssertionError):
        FourQubitH2.evaluate_batch([[0, 14, 1, 2]], params, pool)
    # ragged batch sizes around the team/CTA granularity
    for B in (1, 7, 31, 33, 257):
        ks = sample_structure_lists(pool, 4, None, B, seed=B)
        out = FourQubitH2.evaluate_batch(ks, params, pool, want_grad=False)
        ref = np.array([sim.energy(4, [int(x) for x in k], pool.pool, params, terms, FourQubitH2.init_state)
                        for k in ks[:5]])
        assert np.max(np.abs(out["energy"][:5] - ref)) <= E_TOL
        assert np.all(np.isfinite(out["energy"]))
    # sum instead of mean (avg_gradients=False, mcts.py:347)
    ks = sample_structure_lists(pool, 4, None, 8, seed=4)
    a = FourQubitH2.evaluate_batch(ks, params, pool, avg=True)["grad_dense"]
    s = FourQubitH2.evaluate_batch(ks, params, pool, avg=False)["grad_dense"]
    assert np.max(np.abs(8 * a - s)) <= 1e-12


def test_device_pointer_entry_is_equivalent():
    import torch
    from qas_b200.qml_models.qml_models_legacy import LiH
    pool = helpers.near_cx_pool(10)
    c, p, B = len(pool), 20, 64
    params = np.random.default_rng(0).standard_normal((p, c, 3))
    ks = sample_structure_lists(pool, p, {"CNOT": 10}, B, seed=0)
    host = LiH.evaluate_batch(ks, params, pool, compact=True)
    ev = LiH.evaluator(pool)
    dk = torch.from_numpy(ks).cuda()
    dp = torch.from_numpy(params).cuda()
    de = torch.empty(B, dtype=torch.float64, device="cuda")
    dgc = torch.empty(B, p, 3, dtype=torch.float64, device="cuda")
    dgd = torch.empty(p, c, 3, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()
    ev.evaluate_device(dk, dp, de, dgc, dgd, stream=stream)
    torch.cuda.synchronize()
    assert np.array_equal(de.cpu().numpy(), host["energy"])
    assert np.array_equal(dgc.cpu().numpy(), host["grad_compact"])
    assert np.array_equal(dgd.cpu().numpy(), host["grad_dense"])


def test_dense_mean_without_zero_fill_reads_only_written_entries():
    """Only the dense batch mean requested (the search's gradient step, mcts.py:345-347): the compact gradients are not
    zero-filled and the reduction reads exactly the entries the kernels write.  Bit-identical to the zero-filled path;
    a circuit with an out-of-range key (device-pointer calls are not range-checked on the host: NaN energy, nothing
    written) contributes zeros although the scratch buffer still holds the previous call's values."""
    import torch
    from qas_b200.qml_models.qml_models_legacy import LiH
    pool = helpers.near_cx_pool(10)
    c, p, B = len(pool), 20, 96
    params = np.random.default_rng(3).standard_normal((p, c, 3))
    ks = sample_structure_lists(pool, p, {"CNOT": 10}, B, seed=21)
    ref = LiH.evaluate_batch(ks, params, pool, compact=True)              # compact requested: zero-filled path
    ev = LiH.evaluator(pool)
    dp = torch.from_numpy(params).cuda()
    de = torch.empty(B, dtype=torch.float64, device="cuda")
    dgd = torch.empty(p, c, 3, dtype=torch.float64, device="cuda")
    ev.evaluate_device(torch.from_numpy(ks).cuda(), dp, de, None, dgd)
    torch.cuda.synchronize()
    assert np.array_equal(de.cpu().numpy(), ref["energy"])
    assert np.array_equal(dgd.cpu().numpy(), ref["grad_dense"])
    bad = ks.copy()
    bad[5, 3] = c + 2
    ev.evaluate_device(torch.from_numpy(bad).cuda(), dp, de, None, dgd)
    torch.cuda.synchronize()
    e = de.cpu().numpy()
    assert np.isnan(e[5]) and np.all(np.isfinite(np.delete(e, 5)))
    dense = helpers.dense_from_compact(ks, ref["grad_compact"], c)
    dense[5] = 0.0
    assert np.max(np.abs(dgd.cpu().numpy() - dense.sum(axis=0) / B)) <= 1e-13


# --------------------------------------------------------------------------- BASELINE-size properties
def test_baseline_size_properties_lih():
    """LiH near_cx at the script's depth (p=20) and a saturating batch: properties that do not
    need the oracle at full size."""
    from qas_b200.qml_models.qml_models_legacy import LiH
    n, p, B = 10, 20, 8192
    pool = helpers.near_cx_pool(n)
    c = len(pool)
    params = np.random.default_rng(10).standard_normal((p, c, 3))
    ks = sample_structure_lists(pool, p, {"CNOT": p // 2}, B, seed=10)
    out = LiH.evaluate_batch(ks, params, pool, compact=True)
    e = out["energy"]
    # (1) spectral bounds of the 10-qubit operator
    ev = np.linalg.eigvalsh(sim.hamiltonian_matrix(n, LiH.terms()).real)
    assert e.min() >= ev[0] - 1e-9 and e.max() <= ev[-1] + 1e-9
    # (2) determinism + permutation invariance
    perm = np.random.default_rng(0).permutation(B)
    out2 = LiH.evaluate_batch(ks[perm], params, pool, compact=True)
    assert np.array_equal(out2["energy"], e[perm])
    assert np.array_equal(out2["grad_compact"], out["grad_compact"][perm])
    # (3) dense mean = mean over the batch of the scattered compact gradients (zeros included)
    dense = np.zeros((p, c, 3))
    for i in range(p):
        np.add.at(dense[i], ks[:, i], out["grad_compact"][:, i, :])
    assert np.max(np.abs(dense / B - out["grad_dense"])) <= 1e-12
    # (4) parameter-shift identity ties the adjoint kernel to the energy kernel at full size:
    #     dE/dtheta = (E(theta + pi/2) - E(theta - pi/2)) / 2 for every Rot angle
    rng = np.random.default_rng(1)
    for _ in range(6):
        i, j = int(rng.integers(0, p)), int(rng.integers(0, 3))
        key = 2 * int(rng.integers(0, n))             # a Rot key
        sel = np.nonzero(ks[:, i] == key)[0]
        if len(sel) == 0:
            continue
        pp, pm = params.copy(), params.copy()
        pp[i, key, j] += np.pi / 2
        pm[i, key, j] -= np.pi / 2
        ep = LiH.evaluate_batch(ks[sel], pp, pool, want_grad=False)["energy"]
        em = LiH.evaluate_batch(ks[sel], pm, pool, want_grad=False)["energy"]
        assert np.max(np.abs(0.5 * (ep - em) - out["grad_compact"][sel, i, j])) <= G_TOL
    # (5) oracle spot check on a few members of the big batch
    for b in (0, B // 2, B - 1):
        k = [int(x) for x in ks[b]]
        assert abs(e[b] - sim.energy(n, k, pool.pool, params, LiH.terms(), LiH.init_state)) <= E_TOL


@pytest.mark.parametrize("cfg", ["lih_10q", "h2o_8q"])
def test_support_restricted_h_application_equals_dense(cfg, monkeypatch):
    """The support-restricted H application (planner descriptor + per-coset class skipping) only
    leaves out exact zeros: energies and gradients equal the dense application's on the same batch
    (both against the oracle on a sample, and against each other on the whole batch)."""
    from qas_b200.evaluator import BatchEvaluator
    name, n, p, mkpool, lim = helpers.CONFIGS[cfg]
    model = helpers.get_model(name)
    pool = mkpool()
    # shallow circuits too, so that many different support dimensions occur
    outs = {}
    for hs in ("0", "1"):
        monkeypatch.setenv("QAS_B200_HSPARSE", hs)
        ev = BatchEvaluator(n, model.terms(), pool, init=model.init_state)
        res = []
        for depth, seed in ((p, 5), (max(4, p // 3), 6), (3, 7)):
            params = np.random.default_rng(seed).standard_normal((depth, len(pool), 3))
            ks = sample_structure_lists(pool, depth, {"CNOT": depth // 2}, 2048, seed=seed)
            res.append((ks, params, ev.evaluate(ks, params, want_grad=True, dense=True, compact=True)))
        outs[hs] = res
        ev.close()
    for (ks, params, a), (_, _, b) in zip(outs["0"], outs["1"]):
        assert np.max(np.abs(a["energy"] - b["energy"])) <= 1e-12
        assert np.max(np.abs(a["grad_compact"] - b["grad_compact"])) <= 1e-12
        assert np.max(np.abs(a["grad_dense"] - b["grad_dense"])) <= 1e-12
        for bi in range(0, 2048, 512):
            k = [int(x) for x in ks[bi]]
            e = sim.energy(n, k, pool.pool, params, model.terms(), model.init_state)
            assert abs(b["energy"][bi] - e) <= 1e-9


def test_support_restricted_path_with_the_full_gate_vocabulary(monkeypatch):
    """Controlled rotations, parity-diagonal (Ising) gates and swaps on 8 qubits, shallow enough that the
    support stays small: restricted and dense H application agree and both match the oracle."""
    from qas_b200.qml_models.qml_models_legacy import H2O
    n, p, B = 8, 5, 768
    pool = helpers.full_vocab_pool(n)
    c = len(pool)
    rng = np.random.default_rng(12)
    ks = rng.integers(0, c, size=(B, p)).astype(np.int32)
    params = rng.standard_normal((p, c, 3))
    res = {}
    for hs in ("0", "1"):
        monkeypatch.setenv("QAS_B200_HSPARSE", hs)
        ev = BatchEvaluator(n, H2O.terms(), pool, init=("basis", 0b10100110))
        res[hs] = ev.evaluate(ks, params, want_grad=True, dense=True, compact=True)
        ev.close()
    assert np.max(np.abs(res["0"]["energy"] - res["1"]["energy"])) <= 1e-12
    assert np.max(np.abs(res["0"]["grad_compact"] - res["1"]["grad_compact"])) <= 1e-12
    for b in range(0, B, 96):
        k = [int(x) for x in ks[b]]
        e = sim.energy(n, k, pool.pool, params, H2O.terms(), ("basis", 0b10100110))
        g = sim.grad_adjoint(n, k, pool.pool, params, H2O.terms(), ("basis", 0b10100110))
        assert abs(res["1"]["energy"][b] - e) <= E_TOL
        for i, key in enumerate(k):
            assert np.max(np.abs(res["1"]["grad_compact"][b, i] - g[i, key])) <= G_TOL


def _eval_both_paths(monkeypatch, n, terms, pool, init, cases, B):
    """The same batches through the support-compressed path (default) and the classic path (QAS_B200_COMPRESS=0)."""
    outs = {}
    monkeypatch.setenv("QAS_B200_COMPRESS_MIN_N", "7")       # by default registers below 9 qubits keep the classic path
    for flag in ("0", "1"):
        monkeypatch.setenv("QAS_B200_COMPRESS", flag)
        ev = BatchEvaluator(n, terms, pool, init=init)
        res = []
        for ks, params in cases:
            full = ev.evaluate(ks, params, want_grad=True, dense=True, compact=True)
            mean_only = ev.evaluate(ks, params, want_grad=True, dense=True, compact=False)      # fused finalize + reduce
            reward = ev.evaluate(ks, params, want_grad=False)
            tiny = ev.evaluate(ks[:5], params, want_grad=True, dense=True, compact=True)         # host-compiled records
            res.append((full, mean_only, reward, tiny))
        outs[flag] = (res, ev.stats()["path"])
        ev.close()
    return outs


@pytest.mark.parametrize("cfg", ["lih_10q", "h2o_8q"])
def test_support_compressed_path_equals_classic_and_oracle(cfg, monkeypatch):
    """Support-compressed evaluation (c-tapes, one launch per bin of the support dimension, sign tables read at
    the physical index) == classic support-restricted evaluation <= 1e-12 on whole batches of several depths,
    == oracle <= 1e-9 / 1e-8 on a sample; the fused finalize + batch-mean kernel and the host-compiled
    (<= 16 circuits) records give the same numbers."""
    name, n, p, mkpool, lim = helpers.CONFIGS[cfg]
    model = helpers.get_model(name)
    pool = mkpool()
    B = 2048
    cases = []
    for depth, seed in ((p, 5), (max(4, p // 3), 6), (3, 7)):
        params = np.random.default_rng(seed).standard_normal((depth, len(pool), 3))
        cases.append((sample_structure_lists(pool, depth, {"CNOT": depth // 2}, B, seed=seed), params))
    outs = _eval_both_paths(monkeypatch, n, model.terms(), pool, model.init_state, cases, B)
    assert outs["1"][1] == 3 and outs["0"][1] == 0           # the compressed kernel really ran (min-n knob for 8 qubits)
    for (ks, params), a, b in zip(cases, outs["0"][0], outs["1"][0]):
        assert np.max(np.abs(a[0]["energy"] - b[0]["energy"])) <= 1e-12
        assert np.max(np.abs(a[0]["grad_compact"] - b[0]["grad_compact"])) <= 1e-12
        assert np.max(np.abs(a[0]["grad_dense"] - b[0]["grad_dense"])) <= 1e-12
        for r in (a, b):
            assert np.max(np.abs(r[1]["grad_dense"] - r[0]["grad_dense"])) <= 1e-13
            assert np.array_equal(r[1]["energy"], r[0]["energy"])
            assert np.max(np.abs(r[2]["energy"] - r[0]["energy"])) <= 1e-12
            assert np.max(np.abs(r[3]["energy"] - r[0]["energy"][:5])) <= 1e-12
            assert np.max(np.abs(r[3]["grad_compact"] - r[0]["grad_compact"][:5])) <= 1e-12
        for bi in range(0, B, 256):
            k = [int(x) for x in ks[bi]]
            e = sim.energy(n, k, pool.pool, params, model.terms(), model.init_state)
            g = sim.grad_adjoint(n, k, pool.pool, params, model.terms(), model.init_state)
            assert abs(b[0]["energy"][bi] - e) <= E_TOL
            for i, key in enumerate(k):
                assert np.max(np.abs(b[0]["grad_compact"][bi, i] - g[i, key])) <= G_TOL


def test_support_compressed_path_with_the_full_gate_vocabulary(monkeypatch):
    """Every gate of the vocabulary from the all-zero state on 8 and 11 qubits (single- and multi-warp teams):
    controlled gates whose control never fires on the support, Ising phases, swaps."""
    from qas_b200.qml_models.qml_models_legacy import H2O
    rng = np.random.default_rng(21)
    for n, p, B in ((8, 6, 1024), (11, 9, 512)):
        pool = helpers.full_vocab_pool(n)
        c = len(pool)
        if n == 8:
            terms = H2O.terms()
        else:
            words = ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(40)] + ["I" * n, "Z" + "I" * (n - 1)]
            terms = [(float(rng.standard_normal()), w) for w in words]
        ks = rng.integers(0, c, size=(B, p)).astype(np.int32)
        params = rng.standard_normal((p, c, 3))
        outs = _eval_both_paths(monkeypatch, n, terms, pool, ("zero", 0), [(ks, params)], B)
        a, b = outs["0"][0][0], outs["1"][0][0]
        assert np.max(np.abs(a[0]["energy"] - b[0]["energy"])) <= 1e-12
        assert np.max(np.abs(a[0]["grad_compact"] - b[0]["grad_compact"])) <= 1e-12
        assert np.max(np.abs(b[1]["grad_dense"] - b[0]["grad_dense"])) <= 1e-13
        assert np.max(np.abs(b[3]["grad_compact"] - b[0]["grad_compact"][:5])) <= 1e-12
        for bi in range(0, B, 128):
            k = [int(x) for x in ks[bi]]
            e = sim.energy(n, k, pool.pool, params, terms, ("zero", 0))
            g = sim.grad_adjoint(n, k, pool.pool, params, terms, ("zero", 0))
            assert abs(b[0]["energy"][bi] - e) <= E_TOL
            for i, key in enumerate(k):
                assert np.max(np.abs(b[0]["grad_compact"][bi, i] - g[i, key])) <= G_TOL


def test_deep_tapes_fall_back_to_the_global_memory_kernel_with_a_consistent_layout():
    """the index layout (swizzled shared memory vs plain global slab) is decided once per call from the
    shared-memory fit, not from the qubit count: an n = 12 batch whose tapes are too long for the shared-memory
    kernel runs on the global-memory kernel and still matches the oracle."""
    n, p, B = 12, 900, 3
    pool = helpers.near_cx_pool(n)
    rng = np.random.default_rng(33)
    words = ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(12)]
    terms = [(float(rng.standard_normal()), w) for w in words]
    params = rng.standard_normal((p, len(pool), 3))
    ks = sample_structure_lists(pool, p, {"CNOT": p // 2}, B, seed=33)
    ev = BatchEvaluator(n, terms, pool, init=("basis", 0b101010101010))
    out = ev.evaluate(ks, params, want_grad=True, dense=False, compact=True)
    assert ev.stats()["path"] == 4
    ev.close()
    for b in range(B):
        k = [int(x) for x in ks[b]]
        assert abs(out["energy"][b] - sim.energy(n, k, pool.pool, params, terms, ("basis", 0b101010101010))) <= E_TOL
        g = sim.grad_adjoint(n, k, pool.pool, params, terms, ("basis", 0b101010101010))
        for i, key in enumerate(k):
            assert np.max(np.abs(out["grad_compact"][b, i] - g[i, key])) <= G_TOL


@pytest.mark.parametrize("n,tb,cs,init,vocab", [
    (11, 11, 0, ("zero", 0), "near_cx"), (12, 11, 2, ("basis", 0b100110101101), "full"), (12, 12, 1, ("plus", 0), "full"),
    (13, 11, 4, ("basis", 0b1000000000001), "near_cx"), (13, 12, 2, ("zero", 0), "full"), (14, 11, 1, ("zero", 0), "near_cx"),
    (14, 12, 4, ("plus", 0), "near_cx"), (15, 11, 4, ("zero", 0), "near_cx")])
def test_tiled_kernel_tile_sizes_cluster_sizes_start_states(n, tb, cs, init, vocab, monkeypatch):
    """eval_tiled_kernel (kernels_t.cuh) forced onto 11..15-qubit registers (QAS_B200_TILED_FORCE_N): both tile sizes,
    clusters of 1 / 2 / 4 CTAs, all three start states, the full gate vocabulary (controlled gates and parity
    phases whose masks fall on tile-number bits), energy-only and gradient launches -- against the oracle, and
    against the round-1 global-memory kernel (QAS_B200_TILED=0) to 1e-12."""
    monkeypatch.setenv("QAS_B200_COMPRESS", "0")
    monkeypatch.setenv("QAS_B200_TILED_FORCE_N", "11")
    monkeypatch.setenv("QAS_B200_TILED_TB", str(tb))
    monkeypatch.setenv("QAS_B200_TILED_CS", str(cs))
    rng = np.random.default_rng(1000 + n + tb + cs)
    pool = helpers.near_cx_pool(n) if vocab == "near_cx" else helpers.full_vocab_pool(n)
    c = len(pool)
    terms = [(float(rng.standard_normal()), "".join(rng.choice(list("IXYZ"), size=n, p=[.55, .15, .15, .15])))
             for _ in range(9)] + [(0.25, "I" * n)]
    p, B = (40 if vocab == "near_cx" else 18), 5
    if vocab == "near_cx":
        ks = sample_structure_lists(pool, p, {"CNOT": p // 2}, B, seed=n)
    else:
        ks = rng.integers(0, c, size=(B, p)).astype(np.int32)
    params = rng.standard_normal((p, c, 3))
    ev = BatchEvaluator(n, terms, pool, init)
    out = ev.evaluate(ks, params, want_grad=True, compact=True)
    assert ev.stats()["path"] == 4
    out_e = ev.evaluate(ks, params, want_grad=False)
    assert ev.stats()["path"] == 4
    ev.close()
    assert np.max(np.abs(out_e["energy"] - out["energy"])) <= 1e-12
    for b in range(2):
        k = [int(x) for x in ks[b]]
        assert abs(out["energy"][b] - sim.energy(n, k, pool.pool, params, terms, init)) <= E_TOL
        ga = sim.grad_adjoint(n, k, pool.pool, params, terms, init)
        dense = helpers.dense_from_compact([k], out["grad_compact"][b:b + 1], c)[0]
        assert np.max(np.abs(dense - ga)) <= G_TOL
    monkeypatch.setenv("QAS_B200_TILED", "0")
    monkeypatch.setenv("QAS_B200_TILED_FORCE_N", "99")
    ev = BatchEvaluator(n, terms, pool, init)
    ref = ev.evaluate(ks, params, want_grad=True, compact=True)
    assert ev.stats()["path"] in (0, 1)
    ev.close()
    assert np.max(np.abs(ref["energy"] - out["energy"])) <= 1e-12
    assert np.max(np.abs(ref["grad_compact"] - out["grad_compact"])) <= 1e-12


def test_linearity_in_hamiltonian():
    """E_{H1+H2} = E_{H1} + E_{H2} for the same circuits (X-mask grouping must not mix terms)."""
    from qas_b200.qml_models.qml_models_legacy import H2O
    n, p, B = 8, 50, 512
    pool = helpers.near_cx_pool(n)
    terms = H2O.terms()
    half = len(terms) // 2
    params = np.random.default_rng(3).standard_normal((p, len(pool), 3))
    ks = sample_structure_lists(pool, p, {"CNOT": 25}, B, seed=3)
    outs = []
    for tt in (terms, terms[:half], terms[half:]):
        ev = BatchEvaluator(n, tt, pool, ("zero", 0))
        outs.append(ev.evaluate(ks, params))
        ev.close()
    assert np.max(np.abs(outs[0]["energy"] - outs[1]["energy"] - outs[2]["energy"])) <= 1e-10
    assert np.max(np.abs(outs[0]["grad_dense"] - outs[1]["grad_dense"] - outs[2]["grad_dense"])) <= 1e-10


def test_kagome_16q_global_memory_path_small_depth():
    from qas_b200.qml_models.kagome_heisenberg_model import KagomeHeisenberg as K
    n, p, B = 16, 6, 2
    pool = QMLPool(n, ["Rot", "PlaceHolder"], ["CNOT"], False, helpers.line(n))
    params = np.random.default_rng(16).standard_normal((p, len(pool), 3))
    ks = sample_structure_lists(pool, p, {"CNOT": 3}, B, seed=16)
    out = K.evaluate_batch(ks, params, pool, compact=True)
    for b in range(B):
        k = [int(x) for x in ks[b]]
        assert abs(out["energy"][b] - sim.energy(n, k, pool.pool, params, K.terms(), K.init_state)) <= E_TOL
        ga = sim.grad_adjoint(n, k, pool.pool, params, K.terms(), K.init_state)
        dense = helpers.dense_from_compact([k], out["grad_compact"][b:b + 1], len(pool))[0]
        assert np.max(np.abs(dense - ga)) <= G_TOL


def _sharded_worker(rank, world, port, ks, params, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from qas_b200.qml_models.qml_models_legacy import LiH
    pool = helpers.near_cx_pool(10)
    out = LiH.evaluate_batch_sharded(ks, params, pool)
    if rank == 0:
        q.put((out["energy"], out["grad_dense"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_batch_matches_single_gpu():
    """Circuits sharded over 2 GPUs, rewards all-gathered and the gradient all-reduced over NCCL:
    identical (<=1e-12) to the single-GPU result.  Skipped on a 1-GPU box."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from qas_b200.qml_models.qml_models_legacy import LiH
    pool = helpers.near_cx_pool(10)
    p, B = 20, 301
    ks = sample_structure_lists(pool, p, {"CNOT": 10}, B, seed=21)
    params = np.random.default_rng(21).standard_normal((p, len(pool), 3))
    ref = LiH.evaluate_batch(ks, params, pool)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, ks, params, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    energy, grad = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert np.max(np.abs(energy - ref["energy"])) <= 1e-12
    assert np.max(np.abs(grad - ref["grad_dense"])) <= 1e-12


def _peer_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from qas_b200.distributed import PeerExchange
    B, npar = 1001, 333                      # odd sizes: the row is padded to an even number of doubles
    px = PeerExchange(B + npar, dev)
    stream = torch.cuda.Stream(device=dev)
    out = torch.empty(npar, dtype=torch.float64, device=dev)
    worst_rows, worst_mean = 0.0, 0.0
    for step in range(5):                    # both buffer parities, several times
        rows_ref = [np.random.default_rng(100 * step + r).standard_normal(px.row) for r in range(world)]
        mine = torch.from_numpy(rows_ref[rank]).to(dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            px.push(mine, stream)
            px.wait_reduce(B, npar, 1.0 / world, out, stream)
        stream.synchronize()
        got = px.rows().cpu().numpy()
        worst_rows = max(worst_rows, max(float(np.max(np.abs(got[r] - rows_ref[r]))) for r in range(world)))
        mean = sum(rows_ref[r][B:B + npar] for r in range(world)) * (1.0 / world)     # ascending rank order
        worst_mean = max(worst_mean, float(np.max(np.abs(out.cpu().numpy() - mean))))
    assert not px.timed_out()
    if rank == 0:
        q.put((worst_rows, worst_mean))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_peer_exchange_rows_and_fixed_order_mean():
    """qas_peer_push_row / qas_peer_wait_reduce (NVLink peer memory, no NCCL on the data path): every rank ends up
    with every rank's packed row bit for bit and with the fixed-order mean of the gradient parts.  Skipped on a
    1-GPU box; bench.py --gpus N runs the same path under its parity self-check."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    worst_rows, worst_mean = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    assert worst_rows == 0.0
    assert worst_mean <= 1e-15


def test_search_and_tuning_drop_in_on_gpu():
    """The reference's driver calls, through the CUDA-backed model: a short nested-MCTS search on the
    4-qubit H2 model followed by fine-tuning must reach chemical accuracy territory, and the golden
    circuit's tuning loop must reproduce the energy the reference recorded."""
    import random
    from qas_b200.mcts import QMLStateBasicGates, circuitModelTuning, search
    from qas_b200.qml_models.qml_models_legacy import FourQubitH2
    random.seed(0)
    np.random.seed(0)
    pool = helpers.near_cx_pool(4)
    p, c = 10, len(pool)
    init = np.random.default_rng(0).standard_normal((p, c, 3))
    params, arc, node, reward, ctl, hist = search(
        model=FourQubitH2, op_pool=pool, target_circuit_depth=p, init_qubit_with_controls=set(),
        init_params=init, num_iterations=6, num_warmup_iterations=2, super_circ_train_lr=0.2,
        early_stop_threshold=5.0, gate_limit_dict={"CNOT": p}, warmup_arc_batchsize=200, search_arc_batchsize=10,
        alpha_max=2, alpha_decay_rate=0.99, prune_constant_max=0.99, prune_constant_min=0.8,
        max_visits_prune_threshold=20, min_num_children=c // 2 + 1, sampling_execute_rounds=10,
        exploit_execute_rounds=10, verbose=0, state_class=QMLStateBasicGates)
    assert len(arc) == p and len(hist) == 6
    m = FourQubitH2(p, c, 3, arc, pool)
    assert abs(m.getReward(params) - reward) < 1e-12
    assert abs(reward + sim.energy(4, arc, pool.pool, params, FourQubitH2.terms(), FourQubitH2.init_state)) < 1e-9
    theta, losses = circuitModelTuning(initial_params=params, model=FourQubitH2, num_epochs=120, k=arc, op_pool=pool,
                                       lr=0.1, grad_noise_factor=0, verbose=0, early_stop_threshold=-2.0)
    assert losses[-1] <= losses[0] + 1e-12 and losses[-1] >= -1.136189454088 - 1e-9
    # golden circuit: continuing the reference's own fine-tuning does not move the energy it recorded
    g = [x for x in helpers.goldens() if x["fixture"] == "h2_hf_a"][0]
    theta, losses = circuitModelTuning(initial_params=g["params"], model=FourQubitH2, num_epochs=3, k=g["k"],
                                       op_pool=pool, lr=1e-4, grad_noise_factor=0, verbose=0, early_stop_threshold=-2.0)
    assert abs(losses[0] - g["fine_tune_loss_tail"][-1]) < 5e-5


def test_device_adam_step_matches_numpy_rule():
    """qas_adam_step = the qml.AdamOptimizer rule of qas_b200.optim.AdamOptimizer (mcts.py:351-352),
    including the nan_to_num / noise conditioning of the gradient (mcts.py:346-350)."""
    import torch
    from qas_b200.optim import AdamOptimizer, DeviceAdam
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal((7, 5, 3))
    host = AdamOptimizer(stepsize=0.05)
    xh = x0.copy()
    xd = torch.from_numpy(x0.copy()).cuda()
    dev = DeviceAdam(xd, stepsize=0.05)
    for step in range(6):
        g = rng.standard_normal(x0.shape)
        noise = rng.standard_normal(x0.shape)
        if step == 2:
            g[0, 0, 0] = np.nan
            g[1, 2, 1] = np.inf
        xh = host.apply_grad(np.nan_to_num(g) + 0.01 * noise, xh)
        dev.step(torch.from_numpy(g).cuda(), torch.from_numpy(noise).cuda(), 0.01)
        torch.cuda.synchronize()
        assert np.max(np.abs(xd.cpu().numpy() - xh)) <= 1e-13 * max(1.0, np.max(np.abs(xh)))


@pytest.mark.parametrize("use_graph", [True, False])
def test_device_tuning_loop_matches_host_loop(use_graph):
    """circuitModelTuningDevice (loop resident on the GPU; one epoch captured into a CUDA graph and replayed, or
    launched epoch by epoch) follows the host loop of circuitModelTuning (reference mcts.py:381-412) step for
    step when no noise is injected -- 4 qubits (thread-per-circuit kernel) and LiH-10q (binned compressed path,
    forked streams inside the captured graph)."""
    from qas_b200.mcts.mcts import circuitModelTuning, circuitModelTuningDevice
    from qas_b200.qml_models.qml_models_legacy import FourQubitH2, LiH
    for model, n, p, k, epochs in ((FourQubitH2, 4, 8, [0, 2, 8, 4, 6, 10, 3, 12], 12),
                                   (LiH, 10, 10, None, 40)):
        pool = helpers.near_cx_pool(n)
        c = len(pool)
        if k is None:
            k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": 5}, 1, seed=2)[0]]
        params = np.random.default_rng(11).standard_normal((p, c, 3))
        th_h, loss_h = circuitModelTuning(params, model, epochs, k, pool, lr=0.05, grad_noise_factor=0.0, verbose=0,
                                          early_stop_threshold=-1e9)
        th_d, loss_d = circuitModelTuningDevice(params, model, epochs, k, pool, lr=0.05, grad_noise_factor=0.0,
                                                early_stop_threshold=-1e9, use_graph=use_graph, check_every=7)
        assert len(loss_h) == len(loss_d) == epochs
        assert np.max(np.abs(np.array(loss_h) - np.array(loss_d))) <= 1e-9
        # parameters: Adam divides by sqrt(v) + eps, which amplifies last-bit differences of gradient components
        # that are zero up to rounding; the losses above are the meaningful comparison
        assert np.max(np.abs(th_h - th_d)) <= (1e-9 if n == 4 else 1e-6)
        assert loss_d[-1] < loss_d[0]
    # early stop: the graph loop reports the reference's stopping epoch
    thr = loss_h[5]
    _, l_stop = circuitModelTuningDevice(params, model, epochs, k, pool, lr=0.05, early_stop_threshold=thr,
                                         use_graph=use_graph, check_every=4)
    first = int(np.nonzero(np.array(loss_h) <= thr)[0][0]) + 1
    assert len(l_stop) == first and np.max(np.abs(np.array(l_stop) - np.array(loss_h[:first]))) <= 1e-9


def test_native_tree_with_resident_parameters_on_gpu():
    """Native tree policy bound to the CUDA evaluator: parameters uploaded once (qas_set_params),
    every rollout reward = one qas_eval_batch(QAS_F_PARAMS_RESIDENT); rewards equal the oracle's."""
    from qas_b200.mcts import QMLStateBasicGates
    from qas_b200.mcts.native import NativeMCTSController, PlaceholderPenalty
    from qas_b200.qml_models.qml_models_legacy import FourQubitH2
    pool = helpers.near_cx_pool(4)
    p, c = 10, len(pool)
    params = np.random.default_rng(5).standard_normal((p, c, 3))
    ctl = NativeMCTSController(FourQubitH2, pool, p, state_class=QMLStateBasicGates, gate_limit_dict={"CNOT": 5},
                               reward_penalty_function=PlaceholderPenalty(pool, p, 1.0), sampling_execute_rounds=10,
                               exploit_execute_rounds=20, seed=3)
    ctl.setParams(params)
    arcs, rewards = ctl.warmupBatch(64)
    terms = FourQubitH2.terms()
    for k, r in list(zip(arcs, rewards))[:16]:
        assert abs(r + sim.energy(4, k, pool.pool, params, terms, FourQubitH2.init_state)) <= E_TOL
    ctl._reset()
    k, leaf, r = ctl.sampleArcAndBackPropagate()
    assert abs(r + sim.energy(4, k, pool.pool, params, terms, FourQubitH2.init_state)) <= E_TOL
    # a plain batch call in between invalidates the resident copy; setParams restores it
    FourQubitH2.evaluate_batch(np.array(arcs[:4], dtype=np.int32), params * 0.5, pool, want_grad=False)
    with pytest.raises(RuntimeError):
        ctl.sampleArcAndBackPropagate()
    ctl.setParams(params)
    k2, leaf2 = ctl.exploitArcWithSuperCircParams()
    assert abs(leaf2.reward + sim.energy(4, k2, pool.pool, params, terms, FourQubitH2.init_state)) <= E_TOL
    assert ctl.stats()["reward_evaluations"] < ctl.stats()["reward_requests"]
    ctl.close()


# ------------------------------------------------------------------------------------------------
# property test (hypothesis): random registers, start states, depths and legal structure lists, kernel vs oracle
from hypothesis import HealthCheck, given, settings, strategies as st_


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(n=st_.integers(2, 11), p=st_.integers(1, 16), seed=st_.integers(0, 10 ** 6), init_kind=st_.sampled_from(["zero", "basis", "plus"]),
       B=st_.sampled_from([1, 7, 40]))
def test_property_random_legal_structure_lists_match_oracle(n, p, seed, init_kind, B):
    """Whatever the register size (thread-per-circuit, team, compressed kernels), start state, depth and batch size
    (host- and device-compiled tapes): energies and every derivative equal the oracle's."""
    rng = np.random.default_rng(seed)
    pool = helpers.near_cx_pool(n)
    c = len(pool)
    init = (init_kind, int(rng.integers(0, 1 << n)) if init_kind == "basis" else 0)
    words = ["".join(rng.choice(list("IXYZ"), size=n, p=[.4, .2, .2, .2])) for _ in range(6)] + ["I" * n]
    terms = [(float(rng.standard_normal()), w) for w in words]
    ks = sample_structure_lists(pool, p, {"CNOT": max(1, p // 2)}, B, seed=seed)
    params = rng.standard_normal((p, c, 3))
    ev = BatchEvaluator(n, terms, pool, init=init)
    out = ev.evaluate(ks, params, want_grad=True, dense=True, compact=True)
    ev.close()
    for b in sorted({0, B // 2, B - 1}):
        k = [int(x) for x in ks[b]]
        assert abs(out["energy"][b] - sim.energy(n, k, pool.pool, params, terms, init)) <= E_TOL
        g = sim.grad_adjoint(n, k, pool.pool, params, terms, init)
        for i, key in enumerate(k):
            assert np.max(np.abs(out["grad_compact"][b, i] - g[i, key])) <= G_TOL


@pytest.mark.parametrize("cfg,B", [("lih_10q", 5), ("lih_10q", 301), ("h2o_8q", 77), ("h2_4q_hf", 130)])
def test_uint8_structure_lists_equal_int32(cfg, B):
    """QAS_F_K_U8: keys uploaded as bytes and widened on the device give bit-identical results (host batches
    compiled on the host, large host batches through the copy stream, device-resident batches)."""
    import torch
    mname, n, p, mk, lim = helpers.CONFIGS[cfg]
    model = helpers.get_model(mname)
    pool = mk()
    c = len(pool)
    params = np.random.default_rng(5).standard_normal((p, c, 3))
    ks = sample_structure_lists(pool, p, lim, B, seed=11)
    a = model.evaluate_batch(ks, params, pool, want_grad=True, compact=True)
    b = model.evaluate_batch(ks.astype(np.uint8), params, pool, want_grad=True, compact=True)
    assert np.array_equal(a["energy"], b["energy"])
    assert np.array_equal(a["grad_compact"], b["grad_compact"])
    assert np.array_equal(a["grad_dense"], b["grad_dense"])
    ev = model.evaluator(pool, 3)
    dev = torch.device("cuda", ev.device)
    e8 = torch.empty(B, dtype=torch.float64, device=dev)
    g8 = torch.empty(p, c, 3, dtype=torch.float64, device=dev)
    ev.evaluate_device(torch.from_numpy(ks.astype(np.uint8)).to(dev), torch.from_numpy(params).to(dev), e8, None, g8)
    torch.cuda.synchronize()
    assert np.array_equal(e8.cpu().numpy(), a["energy"])
    if B > 16:
        assert np.array_equal(g8.cpu().numpy(), a["grad_dense"])
    else:
        # a handful of host circuits is compiled on the host and evaluated in ONE launch sized for the largest support
        # among them; the device-resident call bins them, and the small bins un-apply one gate per adjoint pass
        # (QAS_B200_C_AG1): another summation order of the cross matrices, not another result
        assert np.max(np.abs(g8.cpu().numpy() - a["grad_dense"])) <= 1e-13
    # a key outside the pool is still refused (utils.py:12-15)
    bad = ks.astype(np.uint8).copy()
    bad[0, 0] = c
    with pytest.raises(AssertionError):
        model.evaluate_batch(bad, params, pool, want_grad=False)
