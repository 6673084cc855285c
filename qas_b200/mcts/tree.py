"""Search tree of the nested MCTS: nodes, controller, tree policy.

Behavioural restatement of the reference controller (qas/mcts/mcts.py:18-201) -- expansion with a
shuffled legal-action list (:83-92), back-propagation (:94-98), the UCT score
``Q/N + alpha * sqrt(2 ln N_parent / N)`` with random tie-breaks and the three child policies
(:125-151), pruning of under-performing, often-visited children (:153-174), arc sampling by rounds
of selection + simulation followed by a greedy descent (:186-193) and level-wise exploitation
(:195-201) -- with the reward evaluation routed through the batched CUDA evaluator:

  * ``batchRewards`` evaluates many structure lists in one launch;
  * a reward cache keyed by ``tuple(k)`` is valid for as long as the parameters do not change
    (the driver clears it around every optimiser step), so repeated terminal states are free.
Public names and call shapes are the reference's, so experiment scripts keep working.
"""
import random
from typing import Callable, List, Optional, Set

import numpy as np

from qas_b200.mcts.mcts_state import StateOfMCTS

_NEG_INF = float("-inf")


class TreeNode():
    """One partial structure list with its visit statistics."""

    def __init__(self, state: StateOfMCTS, parent):
        self.state, self.parent = state, parent
        self.children = {}
        self.numVisits, self.totalReward = 0, 0
        self.isTerminal = state.isTerminal()
        self.isFullyExpanded = self.isTerminal

    def mean(self):
        return self.totalReward / self.numVisits

    def __repr__(self):
        return "TreeNode(k=%s, visits=%d, total=%r, terminal=%s, expanded=%s)" % (
            self.state.getCurrK(), self.numVisits, self.totalReward, self.isTerminal, self.isFullyExpanded)


def _uct_scores(parent: TreeNode, children, alpha):
    bonus = 2.0 * np.log(parent.numVisits)
    return [ch.totalReward / ch.numVisits + alpha * np.sqrt(bonus / ch.numVisits) for ch in children]

class MCTSController():
    def __init__(self, model, op_pool, target_circuit_depth: int, state_class=None,
                 reward_penalty_function: Optional[Callable] = None,
                 initial_legal_control_qubit_choice: Optional[Set] = None, alpha=1 / np.sqrt(2),
                 prune_reward_ratio=0.9, max_visits_prune_threshold=50, min_num_children=5,
                 sampling_execute_rounds=20, exploit_execute_rounds=200, sample_policy='local_optimal',
                 exploit_policy='local_optimal', gate_limit_dict: Optional[dict] = None, prune=True):
        self.model = model
        self.pool = op_pool
        self.pool_dict = op_pool.pool
        self.pool_keys = list(self.pool_dict.keys())
        self.max_depth = target_circuit_depth
        self.penalty_func = reward_penalty_function
        self.initial_legal_control_qubit_choice = set() if initial_legal_control_qubit_choice is None \
            else initial_legal_control_qubit_choice
        self.alpha = alpha
        self.first_policy = sample_policy
        self.second_policy = exploit_policy
        self.prune_reward_ratio = prune_reward_ratio
        self.max_visits_prune_threshold = max_visits_prune_threshold
        self.min_num_children = min_num_children
        self.sampling_execute_rounds = sampling_execute_rounds
        self.exploit_execute_rounds = exploit_execute_rounds
        self.prune_counter = 0
        self.prune = prune
        self.initial_state = state_class(op_pool=self.pool, maxDepth=self.max_depth,
                                         qubit_with_actions=self.initial_legal_control_qubit_choice,
                                         gate_limit_dict=gate_limit_dict)
        # evaluation bookkeeping (not upstream): rewards requested / evaluated, per-epoch cache
        self.reward_cache = {}
        self.num_reward_requests = 0
        self.num_reward_evaluations = 0

    # ------------------------------------------------------------------ tree bookkeeping
    def _reset(self):
        self.root = TreeNode(self.initial_state, None)
        self.prune_counter = 0

    def setRoot(self):
        self.root = TreeNode(self.initial_state, None)

    def getActionFromBestChild(self, root: TreeNode, best_child):
        for action, node in root.children.items():
            if node is best_child:
                return action

    def expand(self, node: TreeNode):
        actions = node.state.getLegalActions()
        random.shuffle(actions)
        for action in actions:
            if action not in node.children:
                child = TreeNode(node.state.takeAction(action), node)
                node.children[action] = child
                if len(actions) == len(node.children):
                    node.isFullyExpanded = True
                return child

    def backPropagate(self, node: TreeNode, reward):
        while node is not None:
            node.numVisits += 1
            node.totalReward += reward
            node = node.parent

    # ------------------------------------------------------------------ rewards
    def clearRewardCache(self):
        self.reward_cache = {}

    def simulationWithSuperCircuitParamsAndK(self, k: List[int], params):
        assert len(k) == self.max_depth
        self.num_reward_requests += 1
        key = tuple(k)
        if key in self.reward_cache:
            self.r = self.reward_cache[key]
            return self.r
        p, c, l = params.shape[0], params.shape[1], params.shape[2]
        self.r = self.model(p, c, l, k, self.pool).getReward(params)
        self.num_reward_evaluations += 1
        self.reward_cache[key] = self.r
        return self.r

    def batchRewards(self, ks, params):
        """Rewards of many structure lists in one launch (models that expose evaluate_batch)."""
        self.num_reward_requests += len(ks)
        missing = [k for k in {tuple(k) for k in ks} if k not in self.reward_cache]
        if missing:
            if hasattr(self.model, "evaluate_batch"):
                out = self.model.evaluate_batch(np.array(missing, dtype=np.int32), params, self.pool, want_grad=False)
                if hasattr(self.model, "rewards_from_losses"):      # non-linear reward (VQLS: exp(-10 loss))
                    rewards = self.model.rewards_from_losses(out["energy"])
                else:
                    rewards = getattr(self.model, "reward_sign", -1.0) * np.asarray(out["energy"])
                for k, r in zip(missing, rewards):
                    self.reward_cache[k] = float(r)
            else:
                p, c, l = params.shape
                for k in missing:
                    self.reward_cache[k] = self.model(p, c, l, list(k), self.pool).getReward(params)
            self.num_reward_evaluations += len(missing)
        return [self.reward_cache[tuple(k)] for k in ks]

    # ------------------------------------------------------------------ sampling policies
    def randomSample(self):
        node = self.root
        while not node.state.isTerminal():
            if not node.isFullyExpanded:
                node = self.expand(node)
            else:
                _, node = random.choice(list(node.children.items()))
        return node.state.getCurrK(), node

    def uctSample(self, policy='local_optimal'):
        node = self.root
        while not node.state.isTerminal():
            node = self.expand(node) if not node.isFullyExpanded else self.getBestChild(node, self.alpha, policy=policy)
        return node.state.getCurrK(), node

    def getBestChild(self, node: TreeNode, alpha, policy='local_optimal'):
        children = list(node.children.values())
        if not children:
            raise ValueError("No Legal Child for Node at: %s" % node.state.getCurrK())
        if policy == 'random':
            return random.choice(children)
        if policy not in ('local_optimal', 'local_sub_optimal'):
            raise ValueError("No such policy: %s" % policy)
        # running maximum with ties; the runner-up set is whatever held the lead just before the
        # final maximum took over (for a single child it is that child)
        lead, runner_up, top = [], [], _NEG_INF
        for child, score in zip(children, _uct_scores(node, children, alpha)):
            if score > top:
                runner_up = [child] if top == _NEG_INF else lead
                lead, top = [child], score
            elif score == top:
                lead.append(child)
        return np.random.choice(lead if policy == 'local_optimal' else runner_up)

    def _prune(self, node: TreeNode):
        """Drop children whose mean reward fell below ratio * (parent mean) after more than
        max_visits_prune_threshold visits, worst first, always keeping min_num_children of them."""
        if not self.prune:
            return
        cut = node.mean() * self.prune_reward_ratio
        doomed = sorted(((child.mean(), key) for key, child in node.children.items()
                         if child.mean() < cut and child.numVisits > self.max_visits_prune_threshold),
                        key=lambda item: item[0])
        keep_from = len(doomed) - self.min_num_children
        for rank, (_, key) in enumerate(doomed):
            if rank > keep_from:
                break
            del node.children[key]
            self.prune_counter += 1

    def selectNode(self, node: TreeNode):
        if not node.isFullyExpanded:
            return self.expand(node)
        self._prune(node)
        return self.getBestChild(node, self.alpha, policy=self.first_policy)

    def executeRoundWithSuperCircParamsFromAnyNode(self, node: TreeNode, params):
        if node is None:
            node = self.root
        while not node.state.isTerminal():
            node = self.selectNode(node)
        reward = self.simulationWithSuperCircuitParamsAndK(node.state.getCurrK(), params)
        reward = reward if self.penalty_func is None else self.penalty_func(reward, node)
        self.backPropagate(node, reward)
        return node

    def sampleArcWithSuperCircParams(self, params):
        curr = self.root
        for _ in range(self.sampling_execute_rounds):
            self.executeRoundWithSuperCircParamsFromAnyNode(node=curr, params=params)
        while not curr.state.isTerminal():
            curr = self.getBestChild(curr, self.alpha, policy=self.first_policy)
        return curr.state.getCurrK(), curr

    def exploitArcWithSuperCircParams(self, params):
        curr = self.root
        while not curr.state.isTerminal():
            for _ in range(self.exploit_execute_rounds):
                self.executeRoundWithSuperCircParamsFromAnyNode(node=curr, params=params)
            curr = self.getBestChild(curr, 0, policy=self.second_policy)
        return curr.state.getCurrK(), curr
