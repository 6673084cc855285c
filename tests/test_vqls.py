"""VQLS cost: the two-observable form of qas_b200/qml_models/vqls_model.py against
the reference's own formulation -- one Hadamard-test circuit per (l, l', j, Re/Im) on n + 1 qubits,
qas/qml_models/qml_models_legacy.py:1880-1951 -- restated in oracle/vqls.py."""
import numpy as np
import pytest

import helpers
from oracle import sim, vqls
from qas_b200.qml_models.qml_models_legacy import VQLSDemo, VQLSDemo5Q
from qas_b200.sampling import sample_structure_lists


def _case(model, seed, p=8):
    n = model.n_qubits
    pool = helpers.near_cx_pool(n, helpers.ring(n))
    k = [int(x) for x in sample_structure_lists(pool, p, {"CNOT": p // 2}, 1, seed=seed)[0]]
    params = np.random.default_rng(seed).standard_normal((p, len(pool), 3))
    return n, pool, k, params


def _oracle_loss_and_grad(model, n, pool, k, params):
    tn, td = model.numerator_terms(), model.denominator_terms()
    init = ("plus", 0)
    N = sim.energy(n, k, pool.pool, params, tn, init)
    D = sim.energy(n, k, pool.pool, params, td, init)
    gN = sim.grad_adjoint(n, k, pool.pool, params, tn, init)
    gD = sim.grad_adjoint(n, k, pool.pool, params, td, init)
    loss = 0.5 - 0.5 * abs(N) / (n * abs(D))
    grad = (-0.5 / n) * (np.sign(N) * gN / abs(D) - abs(N) * np.sign(D) * gD / D ** 2)
    return loss, grad


@pytest.mark.parametrize("model", [VQLSDemo, VQLSDemo5Q], ids=lambda m: m.name)
def test_two_observable_form_equals_the_hadamard_test_cost(model):
    assert model.name in ("VQLSDemo_4Q", "VQLSDemo_5Q")
    for seed in (1, 2):
        n, pool, k, params = _case(model, seed)
        ref = vqls.cost_loc(n, model.problem, k, pool.pool, params)       # 2 L^2 (n + 1) circuits on n + 1 qubits
        loss, grad = _oracle_loss_and_grad(model, n, pool, k, params)
        assert abs(loss - ref) < 1e-12
        # quotient-rule gradient against central differences of the reference formulation
        i = next(d for d, key in enumerate(k) if "Rot" in pool[key])
        for j in range(3):
            pp, pm = params.copy(), params.copy()
            pp[i, k[i], j] += 1e-5
            pm[i, k[i], j] -= 1e-5
            fd = (vqls.cost_loc(n, model.problem, k, pool.pool, pp) - vqls.cost_loc(n, model.problem, k, pool.pool, pm)) / 2e-5
            assert abs(grad[i, k[i], j] - fd) < 1e-8


def test_observables_are_hermitian_pauli_sums():
    for model in (VQLSDemo, VQLSDemo5Q):
        for terms in (model.numerator_terms(), model.denominator_terms()):
            assert all(isinstance(c, float) and len(w) == model.n_qubits for c, w in terms)
        # A A >= 0 and its identity coefficient is sum c_l^2
        ident = dict((w, c) for c, w in model.denominator_terms())["I" * model.n_qubits]
        assert abs(ident - sum(c * c for c, _ in model.problem)) < 1e-14
    assert abs(VQLSDemo.rewards_from_losses([0.0, 0.1])[1] - np.exp(-1.0)) < 1e-15      # :1969-1972


@pytest.mark.gpu
@pytest.mark.parametrize("model", [VQLSDemo, VQLSDemo5Q], ids=lambda m: m.name)
def test_vqls_models_on_gpu_match_oracle(model):
    n = model.n_qubits
    pool = helpers.near_cx_pool(n, helpers.ring(n))
    p, B = 10, 48
    c = len(pool)
    params = np.random.default_rng(4).standard_normal((p, c, 3))
    ks = sample_structure_lists(pool, p, {"CNOT": p // 2}, B, seed=9)
    out = model.evaluate_batch(ks, params, pool, want_grad=True, compact=True)
    dense = np.zeros((p, c, 3))
    for b in range(B):
        k = [int(x) for x in ks[b]]
        loss, grad = _oracle_loss_and_grad(model, n, pool, k, params)
        assert abs(out["energy"][b] - loss) <= 1e-9
        for i, key in enumerate(k):
            assert np.max(np.abs(out["grad_compact"][b, i] - grad[i, key])) <= 1e-8
        dense += grad
    assert np.max(np.abs(out["grad_dense"] - dense / B)) <= 1e-8
    # the reference's own formulation on one circuit, and the five interface methods
    k = [int(x) for x in ks[0]]
    m = model(p, c, 3, k, pool)
    ref = vqls.cost_loc(n, model.problem, k, pool.pool, params)
    assert abs(m.getLoss(params) - ref) <= 1e-9
    assert abs(m.getReward(params) - np.exp(-10 * ref)) <= 1e-8
    g = m.getGradient(params)
    assert g.shape == params.shape and np.count_nonzero(g) <= 3 * p
    assert m.toList(params) == sim.to_list(k, pool.pool, params)
