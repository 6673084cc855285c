"""Epoch loop of the circuit search and the fine-tuning loop, on the batched evaluator.

``search`` and ``circuitModelTuning`` accept exactly the keyword arguments of the reference
functions (qas/mcts/mcts.py:216-250 and :381-392; every experiment script calls them by keyword)
and return the same tuples.  Epoch structure (reference :284-376):

  warm-up epoch   sample ``warmup_arc_batchsize`` uniformly random legal arcs, reward each, discard
                  the tree if ``warm_up_reset``;
  search epoch    reset / decay alpha / ramp the prune ratio, then ``search_arc_batchsize`` times:
                  grow the tree by ``sampling_execute_rounds`` rollouts, descend greedily, reward;
  both            gradient of every sampled arc w.r.t. the shared parameters -> nan_to_num ->
                  mean (or sum) -> + Gaussian noise -> one optimiser step -> exploit the tree for
                  the current best arc -> early stopping on the mean of the last penalised rewards.

Differences, all result-preserving: the gradient batch is one
``evaluate_batch`` launch instead of a Python loop over models (:340-347); warm-up rewards are
computed in one launch after all arcs are drawn (``randomSample`` never looks at rewards, :107-114);
rewards are cached per structure list while the parameters are unchanged.
"""
import time
from dataclasses import dataclass, fields
from typing import Any, Callable, Optional, Set

import numpy as np

from qas_b200.mcts.tree import MCTSController
from qas_b200.optim import AdamOptimizer


@dataclass
class SearchConfig:
    """Keyword arguments of ``search`` with the reference's defaults."""
    model: Any = None
    op_pool: Any = None
    target_circuit_depth: int = None
    init_qubit_with_controls: Optional[Set] = None
    init_params: Any = None
    state_class: Any = None
    num_iterations: int = 500
    num_warmup_iterations: int = 20
    warm_up_reset: bool = True
    super_circ_train_optimizer: Any = AdamOptimizer
    early_stop_threshold: float = 0.95
    early_stop_lookback_count: int = 5
    super_circ_train_gradient_noise_factor: float = 0.
    super_circ_train_lr: float = 0.01
    penalty_function: Optional[Callable] = None
    gate_limit_dict: Optional[dict] = None
    warmup_arc_batchsize: int = 200
    search_arc_batchsize: int = 20
    alpha_max: float = 2
    alpha_decay_rate: float = 0.95
    prune_constant_max: float = 0.9
    prune_constant_min: float = 0.2
    max_visits_prune_threshold: int = 100
    min_num_children: int = 5
    sampling_execute_rounds: int = 200
    exploit_execute_rounds: int = 200
    cmab_sample_policy: str = 'local_optimal'
    cmab_exploit_policy: str = 'local_optimal'
    uct_sample_policy: str = 'local_optimal'
    verbose: int = 1
    search_reset: bool = True
    prune: bool = True
    avg_gradients: bool = True
    # not upstream (opt-in): selection rounds issued `leaf_parallel` at a time under a
    # virtual loss and rewarded by one batched launch; 1 = the reference's sequential rounds
    leaf_parallel: int = 1
    virtual_loss: Optional[float] = None

    @classmethod
    def from_call(cls, args, kwargs):
        names = [f.name for f in fields(cls)]
        if len(args) > len(names):
            raise TypeError("search() takes at most %d positional arguments" % len(names))
        merged = dict(zip(names, args))
        for key, value in kwargs.items():
            if key not in names:
                raise TypeError("search() got an unexpected keyword argument %r" % key)
            if key in merged:
                raise TypeError("search() got multiple values for argument %r" % key)
            merged[key] = value
        return cls(**merged)


def batchGradient(model, arcs, params, op_pool, avg_gradients=True):
    """nan_to_num'ed mean (or sum) of the arcs' gradients: one launch for evaluator-backed models,
    the reference's per-model loop otherwise."""
    if hasattr(model, "evaluate_batch"):
        ks = np.asarray(arcs, dtype=np.int32)
        return model.evaluate_batch(ks, params, op_pool, want_grad=True, avg=avg_gradients)["grad_dense"]
    p, c, l = params.shape
    stack = np.nan_to_num(np.stack([model(p, c, l, k, op_pool).getGradient(params) for k in arcs]))
    return stack.mean(axis=0) if avg_gradients else stack.sum(axis=0)


# "auto": use the native tree policy when the model is backed by the CUDA evaluator (that is when the
# Python tree is the bottleneck) and the configuration can run natively (reference state classes,
# penalty None or a PlaceholderPenalty); True: whenever it can run natively; False: never.
USE_NATIVE_TREE = "auto"


def _native_possible(cfg):
    """The native tree computes reward = reward_sign * loss itself: it can only stand in for models whose
    reward is that linear map (not VQLS' exp(-10 loss), not a class that overrides getReward)."""
    from qas_b200.mcts.native import native_penalty
    from qas_b200.qml_models.hamiltonian_model import HamiltonianModel
    name = getattr(cfg.state_class, "name", "")
    model = cfg.model
    linear_reward = not hasattr(model, "rewards_from_losses") and (
        not isinstance(model, type) or not issubclass(model, HamiltonianModel) or model.getReward is HamiltonianModel.getReward)
    return (name in ("QMLStateBasicGates", "QMLStateBasicGatesNoRestrictions") and linear_reward
            and native_penalty(cfg.penalty_function) is not None)


def _penalised(cfg, reward, node):
    return reward if cfg.penalty_function is None else cfg.penalty_function(reward, node)


def _warmup_epoch(cfg, controller, params):
    if hasattr(controller, "warmupBatch"):                 # native tree: sampling, batch reward, back-propagation
        arcs, _ = controller.warmupBatch(cfg.warmup_arc_batchsize)
        if cfg.warm_up_reset:
            controller._reset()
        return arcs
    drawn = [controller.randomSample() for _ in range(cfg.warmup_arc_batchsize)]
    arcs = [k for k, _ in drawn]
    for (k, node), reward in zip(drawn, controller.batchRewards(arcs, params)):
        controller.backPropagate(node, _penalised(cfg, reward, node))
    if cfg.warm_up_reset:
        controller._reset()
    return arcs


def _search_epoch(cfg, controller, params, epoch):
    if cfg.search_reset:
        controller._reset()
    controller.alpha *= cfg.alpha_decay_rate
    span = cfg.num_iterations - cfg.num_warmup_iterations
    done = epoch + 1 - cfg.num_warmup_iterations
    controller.prune_reward_ratio = cfg.prune_constant_min + (cfg.prune_constant_max - cfg.prune_constant_min) / span * done
    arcs = []
    if hasattr(controller, "sampleArcAndBackPropagate"):   # native tree
        for _ in range(cfg.search_arc_batchsize):
            arcs.append(controller.sampleArcAndBackPropagate()[0])
        return arcs
    for _ in range(cfg.search_arc_batchsize):
        k, node = controller.sampleArcWithSuperCircParams(params)
        arcs.append(k)
        reward = controller.simulationWithSuperCircuitParamsAndK(k, params)
        controller.backPropagate(node, _penalised(cfg, reward, node))
    return arcs


def search(*args, **kwargs):
    cfg = SearchConfig.from_call(args, kwargs)
    model, pool = cfg.model, cfg.op_pool
    p, c, l = cfg.target_circuit_depth, len(pool), cfg.init_params.shape[2]
    assert p == cfg.init_params.shape[0]
    assert c == cfg.init_params.shape[1]
    assert cfg.prune_constant_min <= cfg.prune_constant_max
    native = _native_possible(cfg) and (USE_NATIVE_TREE is True or (USE_NATIVE_TREE == "auto" and hasattr(model, "evaluator")))
    if native:
        from qas_b200.mcts.native import NativeMCTSController as Controller
    else:
        Controller = MCTSController
    controller = Controller(
        model=model, op_pool=pool, target_circuit_depth=p, state_class=cfg.state_class,
        reward_penalty_function=cfg.penalty_function,
        initial_legal_control_qubit_choice=cfg.init_qubit_with_controls, alpha=cfg.alpha_max,
        prune_reward_ratio=cfg.prune_constant_min, max_visits_prune_threshold=cfg.max_visits_prune_threshold,
        min_num_children=cfg.min_num_children, sampling_execute_rounds=cfg.sampling_execute_rounds,
        exploit_execute_rounds=cfg.exploit_execute_rounds, sample_policy=cfg.cmab_sample_policy,
        exploit_policy=cfg.cmab_exploit_policy, gate_limit_dict=cfg.gate_limit_dict, prune=cfg.prune,
        leaf_parallel=cfg.leaf_parallel, virtual_loss=cfg.virtual_loss)
    controller.setRoot()
    params = np.array(cfg.init_params, dtype=np.float64)
    optimizer = cfg.super_circ_train_optimizer(cfg.super_circ_train_lr)
    log = print if cfg.verbose >= 1 else (lambda *a, **k: None)
    history, recent = [], []
    best_arc = best_node = best_reward = None
    for epoch in range(cfg.num_iterations):
        t0 = time.time()
        if native:
            controller.setParams(params)                   # upload once per epoch, drops cached rewards
        else:
            controller.clearRewardCache()
        warm = epoch < cfg.num_warmup_iterations
        log("[%s] epoch %d/%d (%s): pool %d, batch %d" % (
            model.name, epoch + 1, cfg.num_iterations, "warm-up" if warm else "search", c,
            cfg.warmup_arc_batchsize if warm else cfg.search_arc_batchsize))
        arcs = _warmup_epoch(cfg, controller, params) if warm else _search_epoch(cfg, controller, params, epoch)
        grad = batchGradient(model, arcs, params, pool, cfg.avg_gradients)
        grad = grad + np.random.randn(p, c, l) * cfg.super_circ_train_gradient_noise_factor
        params = np.array(optimizer.apply_grad(grad=grad, args=params))
        if native:
            controller.setParams(params)                   # rewards of the old parameters are stale
        else:
            controller.clearRewardCache()
        best_arc, best_node = controller.exploitArcWithSuperCircParams(params)
        best_model = model(p, c, l, best_arc, pool)
        best_loss = best_model.getLoss(params)
        best_reward = best_model.getReward(params)
        penalised = _penalised(cfg, best_reward, best_node)
        log("  pruned %d | best reward %r (penalised %r) | best loss %r | %.4f s" % (
            controller.prune_counter, best_reward, penalised, best_loss, time.time() - t0))
        log("  best k: %s" % (best_arc,))
        if cfg.verbose > 1:
            log(best_node.state)
        history.append((epoch, best_arc, best_reward, penalised))
        recent.append(penalised)
        enough = epoch + 1 >= cfg.early_stop_lookback_count + cfg.num_warmup_iterations
        if enough and np.average(recent[-cfg.early_stop_lookback_count:]) >= cfg.early_stop_threshold:
            log("  early stop at epoch %d" % (epoch + 1))
            break
    return params, best_arc, best_node, best_reward, controller, history


def circuitModelTuning(initial_params=None, model=None, num_epochs: int = None, k=None, op_pool=None,
                       opt_callable=AdamOptimizer, lr=0.01, grad_noise_factor=1 / 100, verbose=1,
                       early_stop_threshold=1e-15):
    """Fixed structure list: loss -> gradient -> nan_to_num -> + noise -> optimiser step, until
    ``num_epochs`` or ``loss <= early_stop_threshold`` (reference :381-412).  Evaluator-backed
    models deliver the loss and the gradient of an epoch from a single launch."""
    p, c, l = initial_params.shape
    circuit = model(p, c, l, k, op_pool)
    optimizer = opt_callable(stepsize=lr)
    theta = np.array(initial_params, dtype=np.float64)
    losses = []
    fused = hasattr(model, "evaluate_batch")
    for epoch in range(num_epochs):
        if fused:
            out = model.evaluate_batch([k], theta, op_pool, want_grad=True, avg=True)
            loss, grad = float(out["energy"][0]), out["grad_dense"]
        else:
            loss, grad = circuit.getLoss(theta), np.nan_to_num(circuit.getGradient(theta))
        losses.append(loss)
        if verbose >= 1:
            print("tuning epoch %d/%d: loss %r" % (epoch + 1, num_epochs, loss))
        grad = grad + np.random.randn(p, c, l) * grad_noise_factor
        theta = np.array(optimizer.apply_grad(grad=grad, args=theta))
        if loss <= early_stop_threshold:
            print("early stop at epoch %d" % (epoch + 1))
            break
    return theta, losses


def circuitModelTuningDevice(initial_params=None, model=None, num_epochs: int = None, k=None, op_pool=None, lr=0.01,
                             grad_noise_factor=0.0, verbose=0, early_stop_threshold=1e-15, check_every=1, seed=None,
                             use_graph=True):
    """circuitModelTuning (reference :381-412) with the whole loop on the GPU: the parameter tensor,
    the Adam moments, the gradient, the step counter and the loss history never leave HBM.

    Without gradient noise (the default here) one epoch -- evaluation kernels, loss record, Adam step -- is
    captured ONCE into a CUDA graph and replayed; the host only reads the losses back every ``check_every``
    epochs for the early-stop test (``check_every=1`` reproduces the reference's stopping epoch exactly; a
    larger value may run up to ``check_every - 1`` epochs past it).  With noise, or ``use_graph=False``, the
    epochs are launched one by one (noise from torch's CUDA generator: a different stream of random numbers
    than np.random.randn upstream).  Returns (theta as ndarray, losses)."""
    import ctypes
    import torch
    from qas_b200.optim import DeviceAdam
    p, c, l = initial_params.shape
    ev = model.evaluator(op_pool, l)
    dev = torch.device("cuda", ev.device)
    theta = torch.as_tensor(np.ascontiguousarray(initial_params, dtype=np.float64)).to(dev)
    d_k = torch.as_tensor(np.ascontiguousarray(np.asarray(k, dtype=np.int32)[None, :])).to(dev)
    d_e = torch.empty(1, dtype=torch.float64, device=dev)
    d_g = torch.empty(p, c, l, dtype=torch.float64, device=dev)
    d_losses = torch.zeros(max(1, num_epochs), dtype=torch.float64, device=dev)
    opt = DeviceAdam(theta, stepsize=lr)
    check_every = max(1, int(check_every))
    if use_graph and not grad_noise_factor:
        lib = ev._lib
        d_t = torch.zeros(1, dtype=torch.int32, device=dev)
        d_step = torch.zeros(1, dtype=torch.float64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            ev.evaluate_device(d_k, theta, d_e, None, d_g, avg=True, stream=side)        # sizes every scratch buffer
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            cur = torch.cuda.current_stream(dev)
            sp = ctypes.c_void_p(cur.cuda_stream)
            ev.evaluate_device(d_k, theta, d_e, None, d_g, avg=True, stream=cur)
            rc = lib.qas_tuning_tick(d_t.data_ptr(), d_step.data_ptr(), d_e.data_ptr(), d_losses.data_ptr(), num_epochs,
                                     float(lr), opt.beta1, opt.beta2, sp)
            assert rc == 0
            rc = lib.qas_adam_step_dev(theta.data_ptr(), d_g.data_ptr(), opt.fm.data_ptr(), opt.sm.data_ptr(), theta.numel(),
                                       d_step.data_ptr(), opt.beta1, opt.beta2, opt.eps, sp)
            assert rc == 0
        done, stop_at = 0, None
        while done < num_epochs and stop_at is None:
            burst = min(check_every, num_epochs - done)
            for _ in range(burst):
                graph.replay()
            chunk = d_losses[done:done + burst].cpu().numpy()
            done += burst
            hit = np.nonzero(chunk <= early_stop_threshold)[0]
            if len(hit):
                stop_at = done - burst + int(hit[0]) + 1
        kept = done if stop_at is None else stop_at
        if stop_at is not None and verbose:
            print("early stop at epoch %d" % stop_at)
        losses = [float(x) for x in d_losses[:kept].cpu().numpy()]
        return theta.cpu().numpy(), losses
    stream = torch.cuda.current_stream(dev)
    gen = None
    if grad_noise_factor:
        gen = torch.Generator(device=dev)
        gen.manual_seed(0 if seed is None else int(seed))
    done, stop_at = 0, None
    for epoch in range(num_epochs):
        ev.evaluate_device(d_k, theta, d_e, None, d_g, avg=True, stream=stream)
        d_losses[epoch:epoch + 1].copy_(d_e)
        noise = torch.randn(p, c, l, dtype=torch.float64, device=dev, generator=gen) if gen is not None else None
        opt.step(d_g, noise, grad_noise_factor, stream)
        done = epoch + 1
        if done % check_every == 0 or done == num_epochs:
            lo = done - ((done - 1) % check_every + 1)
            hit = np.nonzero(d_losses[lo:done].cpu().numpy() <= early_stop_threshold)[0]
            if len(hit):
                stop_at = lo + int(hit[0]) + 1
                if verbose:
                    print("early stop at epoch %d" % stop_at)
                break
    losses = [float(x) for x in d_losses[:(done if stop_at is None else stop_at)].cpu().numpy()]
    return theta.cpu().numpy(), losses
