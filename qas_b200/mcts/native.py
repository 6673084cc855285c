"""Native tree policy (libqas_b200: include/qas_b200_mcts.h) behind the controller surface the
search driver uses.

With rewards coming from the GPU a search epoch of the reference is bound by its Python tree
policy (qas/mcts/mcts.py:83-201 and the legality rules of qas/mcts/qml_mcts_state.py:58-148: about
10^5 expansions per epoch at the four-qubit H2 configuration).  ``NativeMCTSController`` keeps the
tree, the legality bookkeeping, UCT / pruning, the rollouts and the per-epoch reward cache in C++;
single rewards go straight to the CUDA evaluator with the parameter super-set resident on the
device.  It follows the same policy as ``MCTSController`` but draws from its own random generator.

Penalties must be expressible natively: ``None`` or a ``PlaceholderPenalty`` (the penalty every
reference script defines by hand, h2_sto-3g.py:55-64).  Anything else keeps the Python controller.
"""
import ctypes

import numpy as np

from qas_b200 import _lib
from qas_b200.qml_models.qml_gate_ops import pool_table

_POLICY = {"local_optimal": 0, "local_sub_optimal": 1, "random": 2}
_NGATES = max(_lib.GATE_IDS.values()) + 1


class MctsConfig(ctypes.Structure):
    _fields_ = [("max_depth", ctypes.c_int32), ("num_qubits", ctypes.c_int32), ("alpha", ctypes.c_double),
                ("prune_reward_ratio", ctypes.c_double), ("max_visits_prune_threshold", ctypes.c_int32),
                ("min_num_children", ctypes.c_int32), ("sampling_execute_rounds", ctypes.c_int32),
                ("exploit_execute_rounds", ctypes.c_int32), ("sample_policy", ctypes.c_int32),
                ("exploit_policy", ctypes.c_int32), ("prune", ctypes.c_int32), ("restricted_root", ctypes.c_int32),
                ("init_touched", ctypes.c_uint32), ("penalty_kind", ctypes.c_int32), ("penalty_limit", ctypes.c_int32),
                ("penalty_scale", ctypes.c_double), ("gate_limit", ctypes.c_int32 * _NGATES), ("seed", ctypes.c_uint64)]


class MctsStats(ctypes.Structure):
    _fields_ = [("nodes", ctypes.c_uint64), ("reward_requests", ctypes.c_uint64), ("reward_evaluations", ctypes.c_uint64),
                ("expansions", ctypes.c_uint64), ("prune_counter", ctypes.c_int32), ("root_visits", ctypes.c_int32),
                ("reward_batches", ctypes.c_uint64)]


ENERGY_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int,
                             ctypes.POINTER(ctypes.c_double))
MCTS_EXPORTS = ["qas_mcts_create", "qas_mcts_destroy", "qas_mcts_last_error", "qas_mcts_bind_evaluator",
                "qas_mcts_set_params", "qas_mcts_reset", "qas_mcts_set_alpha", "qas_mcts_set_prune_ratio",
                "qas_mcts_warmup_batch", "qas_mcts_sample_arc", "qas_mcts_exploit_arc", "qas_mcts_legal_actions",
                "qas_mcts_get_stats", "qas_mcts_set_leaf_parallel"]


def _bind(lib):
    if getattr(lib, "_mcts_bound", False):
        return lib
    vp = ctypes.c_void_p
    lib.qas_mcts_create.restype = ctypes.c_int
    lib.qas_mcts_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(MctsConfig), ctypes.POINTER(_lib.PoolEntry),
                                    ctypes.c_int, ENERGY_FN, vp, ctypes.c_double]
    lib.qas_mcts_destroy.argtypes = [vp]
    lib.qas_mcts_last_error.restype = ctypes.c_char_p
    lib.qas_mcts_last_error.argtypes = [vp]
    lib.qas_mcts_bind_evaluator.argtypes = [vp, vp]
    lib.qas_mcts_set_params.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.qas_mcts_reset.argtypes = [vp]
    lib.qas_mcts_set_alpha.argtypes = [vp, ctypes.c_double]
    lib.qas_mcts_set_prune_ratio.argtypes = [vp, ctypes.c_double]
    lib.qas_mcts_warmup_batch.argtypes = [vp, ctypes.c_int, vp, vp]
    lib.qas_mcts_sample_arc.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_double)]
    lib.qas_mcts_exploit_arc.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.qas_mcts_legal_actions.argtypes = [vp, vp, ctypes.c_int, vp]
    lib.qas_mcts_get_stats.argtypes = [vp, ctypes.POINTER(MctsStats)]
    lib.qas_mcts_set_leaf_parallel.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    lib._mcts_bound = True
    return lib


class PlaceholderPenalty:
    """``r - scale * (ph_count - limit)`` once an arc holds ``limit`` or more PlaceHolders -- the
    penalty function the reference's scripts define inline (h2_sto-3g.py:55-64; scale 10 in
    search_H2O_near_cx.py:80-89).  Callable like theirs, and understood by the native tree."""

    def __init__(self, pool, limit, scale=1.0):
        self.pool, self.limit, self.scale = pool, int(limit), float(scale)

    def __call__(self, r, node):
        ph = sum(1 for key in node.state.getCurrK() if "PlaceHolder" in self.pool[key])
        return r - self.scale * (ph - self.limit) if ph >= self.limit else r


class _Leaf:
    """What the driver and user code read from a returned node: ``node.state.getCurrK()``."""

    class _State:
        def __init__(self, k, pool):
            self._k, self._pool = k, pool

        def getCurrK(self):
            return self._k

        def __repr__(self):
            return "".join("OpAtDepth: {}\tOpKey: {}\tOpName: {}\n".format(i, key, self._pool.pool[key])
                           for i, key in enumerate(self._k))

    def __init__(self, k, pool):
        self.state = self._State(k, pool)


def native_penalty(penalty_function):
    """(kind, limit, scale) if the penalty can run natively, else None."""
    if penalty_function is None:
        return 0, 0, 0.0
    if isinstance(penalty_function, PlaceholderPenalty):
        return 1, penalty_function.limit, penalty_function.scale
    return None


class NativeMCTSController:
    def __init__(self, model, op_pool, target_circuit_depth, state_class=None, reward_penalty_function=None,
                 initial_legal_control_qubit_choice=None, alpha=1 / np.sqrt(2), prune_reward_ratio=0.9,
                 max_visits_prune_threshold=50, min_num_children=5, sampling_execute_rounds=20,
                 exploit_execute_rounds=200, sample_policy="local_optimal", exploit_policy="local_optimal",
                 gate_limit_dict=None, prune=True, seed=None, use_gpu=None, leaf_parallel=1, virtual_loss=None,
                 param_slots=3):
        self._lib = _bind(_lib.load())
        self.model, self.pool, self.max_depth = model, op_pool, int(target_circuit_depth)
        pen = native_penalty(reward_penalty_function)
        if pen is None:
            raise ValueError("penalty function cannot run natively; use PlaceholderPenalty or the Python controller")
        cfg = MctsConfig()
        cfg.max_depth, cfg.num_qubits = self.max_depth, int(op_pool.num_qubits)
        cfg.alpha, cfg.prune_reward_ratio = float(alpha), float(prune_reward_ratio)
        cfg.max_visits_prune_threshold, cfg.min_num_children = int(max_visits_prune_threshold), int(min_num_children)
        cfg.sampling_execute_rounds, cfg.exploit_execute_rounds = int(sampling_execute_rounds), int(exploit_execute_rounds)
        cfg.sample_policy, cfg.exploit_policy = _POLICY[sample_policy], _POLICY[exploit_policy]
        cfg.prune = 1 if prune else 0
        cfg.restricted_root = 0 if getattr(state_class, "name", "") == "QMLStateBasicGatesNoRestrictions" else 1
        cfg.init_touched = sum(1 << int(q) for q in (initial_legal_control_qubit_choice or ()))
        cfg.penalty_kind, cfg.penalty_limit, cfg.penalty_scale = pen
        for g in range(_NGATES):
            cfg.gate_limit[g] = -1
        for name, lim in (gate_limit_dict or {}).items():
            cfg.gate_limit[_lib.GATE_IDS[name]] = int(lim)
        cfg.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        table = pool_table(op_pool)
        self._pool_arr = (_lib.PoolEntry * len(table))(*[_lib.PoolEntry(*e) for e in table])
        self.c = len(table)
        self._params = None
        self._alpha, self._ratio = float(alpha), float(prune_reward_ratio)
        self._cb = ENERGY_FN(self._energy_callback)            # keep alive
        h = ctypes.c_void_p()
        rc = self._lib.qas_mcts_create(ctypes.byref(h), ctypes.byref(cfg), self._pool_arr, self.c, self._cb, None,
                                       float(getattr(model, "reward_sign", -1.0)))
        if rc != 0:
            raise RuntimeError("qas_mcts_create failed (%d): %s" % (rc, self._lib.qas_mcts_last_error(None).decode()))
        self._h = h
        self.leaf_parallel = max(1, int(leaf_parallel))
        if self.leaf_parallel > 1:      # opt-in: waves of selections under a virtual loss, one batched reward call per wave
            self._check(self._lib.qas_mcts_set_leaf_parallel(self._h, self.leaf_parallel, 0 if virtual_loss is None else 1,
                                                             0.0 if virtual_loss is None else float(virtual_loss)))
        # GPU-backed model classes: rewards straight from their evaluator, parameters resident
        self._evaluator = None
        gpu = hasattr(model, "evaluator") if use_gpu is None else use_gpu
        if gpu:
            self._evaluator = model.evaluator(op_pool, int(param_slots))
            self._check(self._lib.qas_mcts_bind_evaluator(self._h, self._evaluator._h))

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.qas_mcts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("native MCTS error %d: %s" % (rc, self._lib.qas_mcts_last_error(self._h).decode()))

    def _energy_callback(self, user, ks_ptr, B, p, out_ptr):
        try:
            ks = np.ctypeslib.as_array(ks_ptr, shape=(B, p)).copy()
            if hasattr(self.model, "evaluate_batch"):
                e = self.model.evaluate_batch(ks, self._params, self.pool, want_grad=False)["energy"]
            else:
                sign = float(getattr(self.model, "reward_sign", -1.0))
                pp, c, l = self._params.shape
                e = [sign * self.model(pp, c, l, [int(x) for x in k], self.pool).getReward(self._params) for k in ks]
            for b in range(B):
                out_ptr[b] = float(e[b])
            return 0
        except Exception:                       # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return 1

    # ------------------------------------------------------------------ controller surface
    def setParams(self, params):
        """New parameter super-set: uploads it (GPU evaluator) and drops every cached reward."""
        self._params = np.ascontiguousarray(params, dtype=np.float64)
        p, c, l = self._params.shape
        if self._evaluator is not None and l != self._evaluator.l:
            raise ValueError("parameter super-set has %d slots per (depth, key) but the bound evaluator was created for %d "
                             "(pass param_slots=%d to NativeMCTSController)" % (l, self._evaluator.l, l))
        self._check(self._lib.qas_mcts_set_params(self._h, self._params.ctypes.data, p, c, l))

    def clearRewardCache(self):
        if self._params is not None:
            self.setParams(self._params)

    def _reset(self):
        self._check(self._lib.qas_mcts_reset(self._h))

    setRoot = _reset

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, value):
        self._alpha = float(value)
        self._check(self._lib.qas_mcts_set_alpha(self._h, self._alpha))

    @property
    def prune_reward_ratio(self):
        return self._ratio

    @prune_reward_ratio.setter
    def prune_reward_ratio(self, value):
        self._ratio = float(value)
        self._check(self._lib.qas_mcts_set_prune_ratio(self._h, self._ratio))

    def stats(self):
        s = MctsStats()
        self._check(self._lib.qas_mcts_get_stats(self._h, ctypes.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    @property
    def prune_counter(self):
        return self.stats()["prune_counter"]

    @property
    def num_reward_requests(self):
        return self.stats()["reward_requests"]

    @property
    def num_reward_evaluations(self):
        return self.stats()["reward_evaluations"]

    def warmupBatch(self, B):
        """B arcs by randomSample, one batched evaluation, penalised back-propagation
        (mcts.py:298-305).  Returns (arcs as lists, raw rewards)."""
        ks = np.empty((B, self.max_depth), dtype=np.int32)
        r = np.empty(B, dtype=np.float64)
        self._check(self._lib.qas_mcts_warmup_batch(self._h, B, ks.ctypes.data, r.ctypes.data))
        return [[int(x) for x in k] for k in ks], r

    def sampleArcAndBackPropagate(self):
        """sampleArcWithSuperCircParams + the arc's own reward back-propagated (mcts.py:332-338)."""
        k = np.empty(self.max_depth, dtype=np.int32)
        r = ctypes.c_double(0.0)
        self._check(self._lib.qas_mcts_sample_arc(self._h, k.ctypes.data, ctypes.byref(r)))
        k = [int(x) for x in k]
        return k, _Leaf(k, self.pool), r.value

    def exploitArcWithSuperCircParams(self, params=None):
        k = np.empty(self.max_depth, dtype=np.int32)
        r, pr = ctypes.c_double(0.0), ctypes.c_double(0.0)
        self._check(self._lib.qas_mcts_exploit_arc(self._h, k.ctypes.data, ctypes.byref(r), ctypes.byref(pr)))
        k = [int(x) for x in k]
        leaf = _Leaf(k, self.pool)
        leaf.reward, leaf.penalised_reward = r.value, pr.value
        return k, leaf

    def legalActions(self, k):
        k = np.ascontiguousarray(np.asarray(k, dtype=np.int32))
        out = np.zeros(self.c, dtype=np.uint8)
        self._check(self._lib.qas_mcts_legal_actions(self._h, k.ctypes.data if len(k) else None, len(k), out.ctypes.data))
        return [i for i in range(self.c) if out[i]]
