"""Search tree of the nested MCTS: nodes, controller, tree policy (Python compatibility controller;
the product tree is qas_b200/csrc/mcts.cu).

Implements the policy of the reference controller (qas/mcts/mcts.py:18-201) -- expansion by a
uniformly random untried legal action (:83-92), back-propagation (:94-98), the UCT score
``Q/N + alpha * sqrt(2 ln N_parent / N)`` with random tie-breaks and the three child policies
(:125-151), pruning of under-performing, often-visited children (:153-174), arc sampling by rounds
of selection + simulation followed by a greedy descent (:186-193) and level-wise exploitation
(:195-201) -- under the reference's public names and call shapes, so experiment scripts keep
working.  How it is organised here is different:

  * every node keeps a pool of its untried legal actions and draws from it (the reference reshuffles
    the whole legal list on every expansion and takes the first action that has no child: the same
    distribution, without the quadratic rescans);
  * all descents share one helper, all rewards one entry point that sends many structure lists to
    the batched CUDA evaluator in one launch (``batchRewards``);
  * rewards are cached by ``tuple(k)`` under a fingerprint of the parameter array, so a different (or
    modified) parameter super-set can never be served stale rewards;
  * opt-in LEAF-PARALLEL rollouts (``leaf_parallel > 1``): the selection rounds of
    :176-184 are issued ``leaf_parallel`` at a time under a virtual loss, their rewards come back from one
    batched evaluation, and the virtual losses are replaced by the real rewards.  This changes the
    statistics the selections see (not semantics-preserving), so the default stays sequential.
"""
import random
from typing import Callable, List, Optional, Set

import numpy as np

from qas_b200.mcts.mcts_state import StateOfMCTS


class TreeNode():
    """One partial structure list with its visit statistics."""

    __slots__ = ("state", "parent", "children", "numVisits", "totalReward", "isTerminal", "isFullyExpanded", "_untried")

    def __init__(self, state: StateOfMCTS, parent):
        self.state, self.parent = state, parent
        self.children = {}
        self.numVisits, self.totalReward = 0, 0
        self.isTerminal = state.isTerminal()
        self.isFullyExpanded = self.isTerminal
        self._untried = None            # legal actions without a child yet (filled at the first expansion)

    def mean(self):
        return self.totalReward / self.numVisits

    def path_to_root(self):
        node = self
        while node is not None:
            yield node
            node = node.parent

    def __repr__(self):
        return "TreeNode(k=%s, visits=%d, total=%r, terminal=%s, expanded=%s)" % (
            self.state.getCurrK(), self.numVisits, self.totalReward, self.isTerminal, self.isFullyExpanded)


def _params_fingerprint(params):
    """Cheap identity of a parameter super-set: buffer address, shape and two checksums.  An optimiser step
    changes the checksums even when it writes in place."""
    flat = np.asarray(params).reshape(-1)
    stride = max(1, flat.size // 4096)
    sample = flat[::stride]
    return (flat.__array_interface__["data"][0], flat.size, float(sample.sum()),
            float(np.dot(sample, np.arange(1, sample.size + 1, dtype=np.float64))))


class MCTSController():
    def __init__(self, model, op_pool, target_circuit_depth: int, state_class=None,
                 reward_penalty_function: Optional[Callable] = None,
                 initial_legal_control_qubit_choice: Optional[Set] = None, alpha=1 / np.sqrt(2),
                 prune_reward_ratio=0.9, max_visits_prune_threshold=50, min_num_children=5,
                 sampling_execute_rounds=20, exploit_execute_rounds=200, sample_policy='local_optimal',
                 exploit_policy='local_optimal', gate_limit_dict: Optional[dict] = None, prune=True,
                 leaf_parallel=1, virtual_loss=None):
        self.model, self.pool = model, op_pool
        self.pool_dict = op_pool.pool
        self.pool_keys = list(self.pool_dict.keys())
        self.max_depth = target_circuit_depth
        self.penalty_func = reward_penalty_function
        self.initial_legal_control_qubit_choice = set(initial_legal_control_qubit_choice or ())
        self.alpha = alpha
        self.first_policy, self.second_policy = sample_policy, exploit_policy
        self.prune, self.prune_counter = prune, 0
        self.prune_reward_ratio = prune_reward_ratio
        self.max_visits_prune_threshold = max_visits_prune_threshold
        self.min_num_children = min_num_children
        self.sampling_execute_rounds = sampling_execute_rounds
        self.exploit_execute_rounds = exploit_execute_rounds
        self.initial_state = state_class(op_pool=self.pool, maxDepth=self.max_depth,
                                         qubit_with_actions=self.initial_legal_control_qubit_choice,
                                         gate_limit_dict=gate_limit_dict)
        # not upstream: leaf-parallel rollouts, reward cache, evaluation counters
        self.leaf_parallel = max(1, int(leaf_parallel))
        self.virtual_loss = virtual_loss           # reward booked for an in-flight leaf; None: the running minimum
        self._min_reward = 0.0
        self.reward_cache, self._cache_tag = {}, None
        self.num_reward_requests = 0
        self.num_reward_evaluations = 0

    # ------------------------------------------------------------------ tree bookkeeping
    def setRoot(self):
        self.root = TreeNode(self.initial_state, None)

    def _reset(self):
        self.setRoot()
        self.prune_counter = 0

    def getActionFromBestChild(self, root: TreeNode, best_child):
        return next((action for action, node in root.children.items() if node is best_child), None)

    def expand(self, node: TreeNode):
        if node._untried is None:
            node._untried = [a for a in node.state.getLegalActions() if a not in node.children]
        pool = node._untried
        if not pool:
            node.isFullyExpanded = True
            return None
        pick = random.randrange(len(pool))
        pool[pick], pool[-1] = pool[-1], pool[pick]
        action = pool.pop()
        child = node.children[action] = TreeNode(node.state.takeAction(action), node)
        node.isFullyExpanded = not pool
        return child

    def backPropagate(self, node: TreeNode, reward):
        for ancestor in node.path_to_root():
            ancestor.numVisits += 1
            ancestor.totalReward += reward

    # ------------------------------------------------------------------ rewards
    def clearRewardCache(self):
        self.reward_cache, self._cache_tag = {}, None

    def _cache_for(self, params):
        tag = _params_fingerprint(params)
        if tag != self._cache_tag:
            self.reward_cache, self._cache_tag = {}, tag
        return self.reward_cache

    def _evaluate_missing(self, missing, params):
        """rewards of structure lists that are not cached yet: one launch when the model can batch"""
        if hasattr(self.model, "evaluate_batch"):
            out = self.model.evaluate_batch(np.array(missing, dtype=np.int32), params, self.pool, want_grad=False)
            if hasattr(self.model, "rewards_from_losses"):      # non-linear reward (VQLS: exp(-10 loss))
                return [float(r) for r in self.model.rewards_from_losses(out["energy"])]
            sign = getattr(self.model, "reward_sign", -1.0)
            return [float(sign * e) for e in out["energy"]]
        p, c, l = params.shape
        return [self.model(p, c, l, list(k), self.pool).getReward(params) for k in missing]

    def batchRewards(self, ks, params):
        """Rewards of many structure lists in one launch (models that expose evaluate_batch)."""
        cache = self._cache_for(params)
        self.num_reward_requests += len(ks)
        keys = [tuple(k) for k in ks]
        missing = [k for k in dict.fromkeys(keys) if k not in cache]
        if missing:
            cache.update(zip(missing, self._evaluate_missing(missing, params)))
            self.num_reward_evaluations += len(missing)
        return [cache[k] for k in keys]

    def simulationWithSuperCircuitParamsAndK(self, k: List[int], params):
        assert len(k) == self.max_depth
        cache = self._cache_for(params)
        self.num_reward_requests += 1
        key = tuple(k)
        if key not in cache:
            p, c, l = params.shape
            cache[key] = self.model(p, c, l, k, self.pool).getReward(params)
            self.num_reward_evaluations += 1
        self.r = cache[key]
        return self.r

    # ------------------------------------------------------------------ descents
    def _descend(self, node, step):
        while not node.state.isTerminal():
            node = step(node)
        return node

    def randomSample(self):
        leaf = self._descend(self.root, lambda n: self.expand(n) if not n.isFullyExpanded
                             else random.choice(list(n.children.values())))
        return leaf.state.getCurrK(), leaf

    def uctSample(self, policy='local_optimal'):
        leaf = self._descend(self.root, lambda n: self.expand(n) if not n.isFullyExpanded
                             else self.getBestChild(n, self.alpha, policy=policy))
        return leaf.state.getCurrK(), leaf

    def getBestChild(self, node: TreeNode, alpha, policy='local_optimal'):
        children = list(node.children.values())
        if not children:
            raise ValueError("No Legal Child for Node at: %s" % node.state.getCurrK())
        if policy == 'random':
            return random.choice(children)
        if policy not in ('local_optimal', 'local_sub_optimal'):
            raise ValueError("No such policy: %s" % policy)
        visits = np.array([ch.numVisits for ch in children], dtype=np.float64)
        totals = np.array([ch.totalReward for ch in children], dtype=np.float64)
        scores = totals / visits + alpha * np.sqrt(2.0 * np.log(node.numVisits) / visits)
        lead = np.flatnonzero(scores == scores.max())
        if policy == 'local_optimal':
            return children[int(np.random.choice(lead))]
        # 'local_sub_optimal' (mcts.py:141-148): whatever held the lead, in child order, just before the final
        # maximum took it over; for a maximum that led from the start, the first child
        first_lead = int(lead[0])
        if first_lead == 0:
            return children[0]
        before = scores[:first_lead]
        return children[int(np.random.choice(np.flatnonzero(before == before.max())))]

    def _prune(self, node: TreeNode):
        """Drop children whose mean reward fell below ratio * (parent mean) after more than
        max_visits_prune_threshold visits, worst first, always keeping min_num_children of them."""
        if not self.prune:
            return
        cut = node.mean() * self.prune_reward_ratio
        doomed = sorted(((child.mean(), key) for key, child in node.children.items()
                         if child.mean() < cut and child.numVisits > self.max_visits_prune_threshold),
                        key=lambda item: item[0])
        for _, key in doomed[:max(0, len(doomed) - self.min_num_children + 1)]:
            del node.children[key]
            self.prune_counter += 1

    def selectNode(self, node: TreeNode):
        if not node.isFullyExpanded:
            return self.expand(node)
        self._prune(node)
        return self.getBestChild(node, self.alpha, policy=self.first_policy)

    def _penalised(self, reward, node):
        return reward if self.penalty_func is None else self.penalty_func(reward, node)

    def executeRoundWithSuperCircParamsFromAnyNode(self, node: TreeNode, params):
        leaf = self._descend(self.root if node is None else node, self.selectNode)
        reward = self.simulationWithSuperCircuitParamsAndK(leaf.state.getCurrK(), params)
        self.backPropagate(leaf, self._penalised(reward, leaf))
        return leaf

    # ------------------------------------------------------------------ leaf-parallel rounds (opt-in)
    def _rounds(self, start, params, rounds):
        """`rounds` selection + simulation rounds from `start`; with leaf_parallel > 1 they are issued in
        waves under a virtual loss and evaluated in one launch per wave."""
        if self.leaf_parallel <= 1:
            for _ in range(rounds):
                self.executeRoundWithSuperCircParamsFromAnyNode(node=start, params=params)
            return
        done = 0
        while done < rounds:
            wave = min(self.leaf_parallel, rounds - done)
            vl = self.virtual_loss if self.virtual_loss is not None else self._min_reward
            leaves = []
            for _ in range(wave):
                leaf = self._descend(start, self.selectNode)
                self.backPropagate(leaf, vl)                     # book the in-flight leaf
                leaves.append(leaf)
            rewards = self.batchRewards([leaf.state.getCurrK() for leaf in leaves], params)
            for leaf, reward in zip(leaves, rewards):
                real = self._penalised(reward, leaf)
                self._min_reward = min(self._min_reward, real)
                for ancestor in leaf.path_to_root():            # swap the virtual loss for the reward
                    ancestor.totalReward += real - vl
            done += wave

    def sampleArcWithSuperCircParams(self, params):
        self._rounds(self.root, params, self.sampling_execute_rounds)
        leaf = self._descend(self.root, lambda n: self.getBestChild(n, self.alpha, policy=self.first_policy))
        return leaf.state.getCurrK(), leaf

    def exploitArcWithSuperCircParams(self, params):
        curr = self.root
        while not curr.state.isTerminal():
            self._rounds(curr, params, self.exploit_execute_rounds)
            curr = self.getBestChild(curr, 0, policy=self.second_policy)
        return curr.state.getCurrK(), curr
