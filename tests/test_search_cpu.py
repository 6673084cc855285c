"""CPU tier: the search driver (tree policy, epoch loop, batched gradient / warm-up rewards, reward
cache, Adam rule) with an oracle-backed model standing in for the CUDA evaluator."""
import random

import numpy as np
import pytest

import helpers
from oracle import sim
from qas_b200.mcts import QMLStateBasicGates, TreeNode, circuitModelTuning, search
from qas_b200.mcts.driver import SearchConfig
from qas_b200.mcts.tree import MCTSController
from qas_b200.models import ModelFromK
from qas_b200.optim import AdamOptimizer
from qas_b200.qml_models.utils import extractParamIndicesQML

TERMS = [(-0.3324042513238793, "II"), (0.39793742484318045, "IZ"), (-0.39793742484318045, "ZI"),
         (-0.01128010425623538, "ZZ"), (0.18093119978423156, "XX")]


class OracleModel(ModelFromK):
    """2-qubit H2 stand-in evaluated by the oracle, one circuit per call (the reference's shape)."""
    name = "OracleTwoQubitH2"
    calls = 0

    def __init__(self, p, c, l, structure_list, op_pool):
        self.p, self.c, self.l, self.k, self.pool = p, c, l, structure_list, op_pool
        self.param_indices = extractParamIndicesQML(self.k, self.pool)

    def getLoss(self, params):
        type(self).calls += 1
        return sim.energy(2, self.k, self.pool.pool, params, TERMS)

    def getReward(self, params):
        return -self.getLoss(params)

    def getGradient(self, params):
        return sim.grad_adjoint(2, self.k, self.pool.pool, params, TERMS)

    def toList(self, params):
        return sim.to_list(self.k, self.pool.pool, params)


class OracleBatchModel(OracleModel):
    """Same arithmetic behind the batched entry the CUDA-backed classes expose."""
    name = "OracleTwoQubitH2Batched"
    reward_sign = -1.0
    batch_calls = 0

    @classmethod
    def evaluate_batch(cls, ks, params, op_pool, want_grad=True, avg=True, compact=False):
        cls.batch_calls += 1
        ks = [[int(x) for x in k] for k in np.asarray(ks)]
        e = np.array([sim.energy(2, k, op_pool.pool, params, TERMS) for k in ks])
        g = None
        if want_grad:
            g = sim.batch_mean_gradient([sim.grad_adjoint(2, k, op_pool.pool, params, TERMS) for k in ks], avg)
        return {"energy": e, "grad_dense": g, "grad_compact": None}


def _run(model, seed, **over):
    random.seed(seed)
    np.random.seed(seed)
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    p = 3
    kw = dict(model=model, op_pool=pool, target_circuit_depth=p, init_qubit_with_controls=set(),
              init_params=np.random.default_rng(seed).standard_normal((p, len(pool), 3)),
              num_iterations=6, num_warmup_iterations=2, super_circ_train_lr=0.1, early_stop_threshold=10.0,
              gate_limit_dict={"CNOT": 2}, warmup_arc_batchsize=6, search_arc_batchsize=4, alpha_max=2,
              prune_constant_max=0.99, prune_constant_min=0.8, max_visits_prune_threshold=3, min_num_children=2,
              sampling_execute_rounds=5, exploit_execute_rounds=5, verbose=0, state_class=QMLStateBasicGates)
    kw.update(over)
    return search(**kw)


def test_batched_driver_equals_per_circuit_driver():
    """Same seeds -> identical parameters, arcs and rewards whether the gradient batch / warm-up
    rewards go through evaluate_batch or through one model per circuit (the reference's loop)."""
    a = _run(OracleModel, 3)
    b = _run(OracleBatchModel, 3)
    assert np.max(np.abs(a[0] - b[0])) < 1e-13
    assert a[1] == b[1] and abs(a[3] - b[3]) < 1e-13
    assert [x[1] for x in a[5]] == [x[1] for x in b[5]]
    assert OracleBatchModel.batch_calls >= 6 + 2          # one gradient launch per epoch + warm-up rewards


def test_search_return_shape_and_reward_cache():
    params, arc, node, reward, controller, history = _run(OracleBatchModel, 1)
    assert params.shape == (3, 6, 3) and len(arc) == 3 and isinstance(node, TreeNode)
    assert len(history) == 6 and history[-1][1] == arc
    assert node.state.getCurrK() == arc
    # depth-3 trees over a 6-key pool revisit terminals constantly: the cache must absorb most requests
    assert controller.num_reward_evaluations < controller.num_reward_requests


def test_search_argument_handling():
    with pytest.raises(TypeError):
        search(model=OracleModel, not_an_argument=1)
    cfg = SearchConfig.from_call((), dict(model=OracleModel))
    assert cfg.num_iterations == 500 and cfg.warmup_arc_batchsize == 200 and cfg.alpha_decay_rate == 0.95
    assert cfg.super_circ_train_optimizer is AdamOptimizer


def test_penalty_function_is_applied_by_the_driver():
    seen = []

    def penalty(r, node):
        seen.append(node.state.getCurrK())
        return r - 1.0
    out = _run(OracleBatchModel, 2, penalty_function=penalty, num_iterations=3)
    assert seen and all(h[3] == h[2] - 1.0 for h in out[5])


def test_uct_and_prune_rules():
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    ctl = MCTSController(OracleModel, pool, 3, state_class=QMLStateBasicGates, alpha=0.0,
                         prune_reward_ratio=0.9, max_visits_prune_threshold=1, min_num_children=1)
    ctl.setRoot()
    root = ctl.root
    while not root.isFullyExpanded:
        ctl.expand(root)
    kids = list(root.children.values())
    assert len(kids) == 2                      # first move: Rot on qubit 0 or 1
    for i, ch in enumerate(kids):
        ch.numVisits, ch.totalReward = 4, 4.0 * (i + 1)
    root.numVisits, root.totalReward = 8, 12.0
    assert ctl.getBestChild(root, 0.0) is kids[1]
    assert ctl.getBestChild(root, 0.0, policy="local_sub_optimal") is kids[0]
    with pytest.raises(ValueError):
        ctl.getBestChild(root, 0.0, policy="nope")
    # mean(kid0) = 1 < 0.9 * 1.5 and visited > threshold -> pruned (min_num_children=1 keeps the rest)
    assert ctl.selectNode(root) is kids[1]
    assert len(root.children) == 1 and ctl.prune_counter == 1


def test_adam_rule():
    opt = AdamOptimizer(0.1)
    x = np.array([1.0, -2.0])
    g = np.array([0.5, -0.25])
    x1 = opt.apply_grad(g, x)
    # first step: m = 0.1 g, v = 0.01 g^2, bias-corrected step = lr * g / (|g| + eps*sqrt(..)) ~ lr * sign(g)
    assert np.allclose(x1, x - 0.1 * np.sign(g), atol=1e-6)
    x2 = opt.apply_grad(g, x1)
    m = 0.9 * 0.1 * g + 0.1 * g
    v = 0.99 * 0.01 * g * g + 0.01 * g * g
    step = 0.1 * np.sqrt(1 - 0.99 ** 2) / (1 - 0.9 ** 2)
    assert np.allclose(x2, x1 - step * m / (np.sqrt(v) + 1e-8), atol=1e-15)


def test_circuit_model_tuning_converges_on_the_oracle_model():
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    k = [0, 4, 2]                                   # Rot(0), CNOT(0,1), Rot(1)
    params = np.random.default_rng(0).standard_normal((3, len(pool), 3))
    for model in (OracleModel, OracleBatchModel):
        np.random.seed(0)
        theta, losses = circuitModelTuning(initial_params=params, model=model, num_epochs=150, k=k, op_pool=pool,
                                           lr=0.1, grad_noise_factor=0, verbose=0, early_stop_threshold=-2.0)
        assert len(losses) == 150 and losses[-1] < losses[0]
        assert losses[-1] < -1.11          # ground energy of the stand-in operator is -1.1373


# ---------------------------------------------------------------------------------------------
# native tree policy (libqas_b200: qas_b200_mcts.h) against the Python controller it restates
def _walk_pools():
    from qas_b200.qml_models.qml_gate_ops import QMLPool
    yield helpers.near_cx_pool(4), {"CNOT": 3}
    yield helpers.near_cx_pool(3, helpers.ring(3)), {}
    yield QMLPool(3, ["T", "Tdg", "S", "Sdg", "PauliZ", "Hadamard", "RX", "PlaceHolder"], ["CZ", "CRZ"], True), {"CZ": 2, "T": 3}
    yield QMLPool(2, ["T", "S", "Rot"], ["CNOT", "IsingXX"], False, [[0, 1], [1, 0]]), {}


def test_native_legality_rules_match_python_state_class():
    """qas_mcts_legal_actions == QMLStateBasicGates.getLegalActions on random walks (the rules of
    qas/mcts/qml_mcts_state.py:58-148), including gate limits, T/S cancellation and the
    unrestricted root of QMLStateBasicGatesNoRestrictions."""
    from qas_b200.mcts import QMLStateBasicGatesNoRestrictions
    from qas_b200.mcts.native import NativeMCTSController
    rng = random.Random(4)
    for pool, limits in _walk_pools():
        for state_class in (QMLStateBasicGates, QMLStateBasicGatesNoRestrictions):
            depth = 9
            ctl = NativeMCTSController(OracleModel, pool, depth, state_class=state_class, gate_limit_dict=limits,
                                       initial_legal_control_qubit_choice=set(), use_gpu=False, seed=1)
            for walk in range(30):
                state = state_class(op_pool=pool, maxDepth=depth, qubit_with_actions=set(), gate_limit_dict=limits)
                k = []
                while True:
                    want = sorted(state.getLegalActions())
                    assert ctl.legalActions(k) == want, (pool.pool, k)
                    if not want:
                        break
                    a = rng.choice(want)
                    k.append(a)
                    state = state.takeAction(a)
            ctl.close()


def test_native_tree_warmup_sampling_and_cache():
    from qas_b200.mcts.native import NativeMCTSController, PlaceholderPenalty
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    p = 3
    params = np.random.default_rng(0).standard_normal((p, len(pool), 3))
    before = OracleBatchModel.batch_calls
    ctl = NativeMCTSController(OracleBatchModel, pool, p, state_class=QMLStateBasicGates, gate_limit_dict={"CNOT": 2},
                               reward_penalty_function=PlaceholderPenalty(pool, 2, 1.0), sampling_execute_rounds=5,
                               exploit_execute_rounds=5, max_visits_prune_threshold=3, min_num_children=2,
                               use_gpu=False, seed=5)
    ctl.setParams(params)
    arcs, rewards = ctl.warmupBatch(40)
    assert OracleBatchModel.batch_calls == before + 1           # ONE batched evaluation
    st = ctl.stats()
    assert st["root_visits"] == 40 and st["reward_requests"] == 40
    assert st["reward_evaluations"] == len({tuple(k) for k in arcs}) < 40
    for k, r in zip(arcs, rewards):
        assert abs(r + sim.energy(2, k, pool.pool, params, TERMS)) < 1e-12      # reward = -E
        state = QMLStateBasicGates(op_pool=pool, maxDepth=p, qubit_with_actions=set(), gate_limit_dict={"CNOT": 2})
        for a in k:                                               # every arc is a legal walk
            assert a in state.getLegalActions()
            state = state.takeAction(a)
    # rollouts: arcs are legal, rewards come from the cache where possible, exploitation returns the
    # best arc's raw and penalised reward
    k, leaf, r = ctl.sampleArcAndBackPropagate()
    assert leaf.state.getCurrK() == k and abs(r + sim.energy(2, k, pool.pool, params, TERMS)) < 1e-12
    k2, leaf2 = ctl.exploitArcWithSuperCircParams()
    ph = sum(1 for a in k2 if "PlaceHolder" in pool[a])
    assert abs(leaf2.penalised_reward - (leaf2.reward - (ph - 2 if ph >= 2 else 0))) < 1e-12
    evals = ctl.stats()["reward_evaluations"]
    ctl.setParams(params)                                         # new parameters -> cache dropped
    ctl._reset()
    ctl.sampleArcAndBackPropagate()
    assert ctl.stats()["reward_evaluations"] > evals
    ctl.close()


def test_native_uct_and_prune_follow_the_reference_rules():
    """alpha = 0, local_optimal: after enough rounds the root's best child by mean reward is the one
    the greedy descent takes; pruning removes under-performing, often-visited children only."""
    from qas_b200.mcts.native import NativeMCTSController
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    p = 3
    params = np.random.default_rng(2).standard_normal((p, len(pool), 3))
    ctl = NativeMCTSController(OracleBatchModel, pool, p, state_class=QMLStateBasicGates, alpha=0.5,
                               prune_reward_ratio=0.5, max_visits_prune_threshold=2, min_num_children=1,
                               sampling_execute_rounds=200, exploit_execute_rounds=50, use_gpu=False, seed=9)
    ctl.setParams(params)
    k, leaf, r = ctl.sampleArcAndBackPropagate()
    # brute force: the best depth-3 legal arc under these parameters
    best = None
    def rec(state, kk):
        nonlocal best
        if len(kk) == p:
            rr = -sim.energy(2, kk, pool.pool, params, TERMS)
            if best is None or rr > best[0]:
                best = (rr, list(kk))
            return
        for a in state.getLegalActions():
            rec(state.takeAction(a), kk + [a])
    rec(QMLStateBasicGates(op_pool=pool, maxDepth=p, qubit_with_actions=set(), gate_limit_dict={}), [])
    k2, leaf2 = ctl.exploitArcWithSuperCircParams()
    assert leaf2.reward >= best[0] - 0.05                # 200 + 150 rounds over a tree of < 60 leaves find (near) the optimum
    assert ctl.stats()["root_visits"] >= 200
    ctl.close()


def test_search_with_native_tree_matches_python_tree_statistically():
    """The driver with the native tree (forced) reaches the same quality as with the Python tree on
    the 2-qubit stand-in and returns the same shapes; trajectories differ (separate generators)."""
    from qas_b200.mcts import driver
    from qas_b200.mcts.native import NativeMCTSController
    py = _run(OracleBatchModel, 7, num_iterations=8)
    old = driver.USE_NATIVE_TREE
    driver.USE_NATIVE_TREE = True
    try:
        nat = _run(OracleBatchModel, 7, num_iterations=8)
    finally:
        driver.USE_NATIVE_TREE = old
    assert isinstance(nat[4], NativeMCTSController)
    assert nat[0].shape == py[0].shape and len(nat[1]) == 3 and len(nat[5]) == 8
    assert nat[2].state.getCurrK() == nat[1]
    assert nat[4].num_reward_evaluations < nat[4].num_reward_requests
    assert nat[3] > py[3] - 0.2                            # best reward in the same ballpark
    nat[4].close()


def test_reward_cache_is_tied_to_the_parameter_values():
    """rewards cached for one parameter super-set are never served for another one -- not for a new
    array, and not for the same array modified in place (both change the fingerprint)."""
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    ctl = MCTSController(OracleBatchModel, pool, 3, state_class=QMLStateBasicGates)
    ctl.setRoot()
    k = [0, 4, 2]
    a = np.random.default_rng(0).standard_normal((3, len(pool), 3))
    r1 = ctl.simulationWithSuperCircuitParamsAndK(k, a)
    assert ctl.simulationWithSuperCircuitParamsAndK(k, a) == r1 and ctl.num_reward_evaluations == 1
    b = a + 0.25
    r2 = ctl.simulationWithSuperCircuitParamsAndK(k, b)
    assert ctl.num_reward_evaluations == 2 and abs(r2 + sim.energy(2, k, pool.pool, b, TERMS)) < 1e-14
    b[0, 0, 1] += 0.5                                   # in place
    r3 = ctl.batchRewards([k, k], b)
    assert ctl.num_reward_evaluations == 3 and abs(r3[0] + sim.energy(2, k, pool.pool, b, TERMS)) < 1e-14 and r3[0] == r3[1]


def test_leaf_parallel_rounds_book_and_replace_virtual_losses():
    """Opt-in leaf-parallel rollouts: every wave books its leaves under a virtual loss, rewards them with ONE
    batched evaluation and replaces the virtual loss by the reward -- afterwards the statistics are those of
    `rounds` ordinary back-propagations (visits) with the true rewards (totals)."""
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    params = np.random.default_rng(4).standard_normal((3, len(pool), 3))
    random.seed(5)
    np.random.seed(5)
    ctl = MCTSController(OracleBatchModel, pool, 3, state_class=QMLStateBasicGates, gate_limit_dict={"CNOT": 2},
                         sampling_execute_rounds=24, leaf_parallel=8, virtual_loss=-3.0, prune=False)
    ctl.setRoot()
    calls0 = OracleBatchModel.batch_calls
    arc, leaf = ctl.sampleArcWithSuperCircParams(params)
    assert OracleBatchModel.batch_calls - calls0 <= 3           # 24 rounds in 3 waves
    assert ctl.root.numVisits == 24

    def check(node):
        if node.children:
            assert node.numVisits >= sum(ch.numVisits for ch in node.children.values())
            for ch in node.children.values():
                check(ch)
        if node.state.isTerminal() and node.numVisits:
            r = -sim.energy(2, node.state.getCurrK(), pool.pool, params, TERMS)
            assert abs(node.totalReward - node.numVisits * r) < 1e-10      # no virtual loss left behind
    check(ctl.root)
    assert abs(ctl.root.totalReward - sum(ch.totalReward for ch in ctl.root.children.values())) < 1e-10
    assert len(arc) == 3 and leaf.state.isTerminal()
    # the driver passes the option through
    out = _run(OracleBatchModel, 6, leaf_parallel=4, num_iterations=3)
    assert out[4].leaf_parallel == 4 and len(out[5]) == 3


def test_native_tree_is_refused_for_non_linear_rewards():
    """the native tree computes reward = reward_sign * loss; VQLS' exp(-10 loss) must keep the Python tree."""
    from qas_b200.mcts.driver import _native_possible
    from qas_b200.mcts.native import PlaceholderPenalty
    from qas_b200.qml_models.qml_models_legacy import LiH
    from qas_b200.qml_models.vqls_model import VQLSDemo
    pool = helpers.near_cx_pool(4)
    base = dict(op_pool=pool, state_class=QMLStateBasicGates, penalty_function=PlaceholderPenalty(pool, 5, 1.0))
    assert _native_possible(SearchConfig.from_call((), dict(model=LiH, **base)))
    assert not _native_possible(SearchConfig.from_call((), dict(model=VQLSDemo, **base)))

    class Custom(LiH):
        def getReward(self, params):
            return float(np.exp(-self.getLoss(params)))
    assert not _native_possible(SearchConfig.from_call((), dict(model=Custom, **base)))


def test_native_tree_leaf_parallel_rounds():
    """qas_mcts_set_leaf_parallel: waves of selections under a virtual loss, one batched reward call per wave; the
    tree ends with `rounds` visits at the root and no virtual loss left behind (every terminal's total = visits x
    its penalised reward)."""
    from qas_b200.mcts.native import NativeMCTSController
    pool = helpers.near_cx_pool(2, [[0, 1], [1, 0]])
    p = 3
    params = np.random.default_rng(8).standard_normal((p, len(pool), 3))
    seq = NativeMCTSController(OracleBatchModel, pool, p, state_class=QMLStateBasicGates, gate_limit_dict={"CNOT": 2},
                               sampling_execute_rounds=40, seed=11, use_gpu=False, prune=False)
    par = NativeMCTSController(OracleBatchModel, pool, p, state_class=QMLStateBasicGates, gate_limit_dict={"CNOT": 2},
                               sampling_execute_rounds=40, seed=11, use_gpu=False, prune=False, leaf_parallel=8,
                               virtual_loss=-2.0)
    for ctl in (seq, par):
        ctl.setRoot()
        ctl.setParams(params)
    k_seq = seq.sampleArcAndBackPropagate()[0]
    k_par = par.sampleArcAndBackPropagate()[0]
    s_seq, s_par = seq.stats(), par.stats()
    assert s_seq["root_visits"] == s_par["root_visits"] == 41          # 40 rounds + the arc's own reward
    assert s_par["reward_batches"] <= 5 and s_seq["reward_batches"] == 0
    assert len(k_seq) == len(k_par) == p
    # both trees hold real rewards only: the root's mean is a mean of genuine rewards of legal arcs
    lo = min(-sim.energy(2, [0, 4, 2], pool.pool, params, TERMS), -1.5)
    assert s_par["reward_requests"] >= 41 and lo < 10
