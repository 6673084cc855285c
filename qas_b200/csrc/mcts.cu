// mcts.cu -- native tree policy of the nested MCTS (host code; C ABI in include/qas_b200_mcts.h).
//
// Behavioural restatement of the reference controller (qas/mcts/mcts.py:32-201) and of the
// legality rules of QMLStateBasicGates (qas/mcts/qml_mcts_state.py:58-148).  What is different
// from the reference: nodes live in one arena and carry their legality bookkeeping (last operator
// on every qubit, touched-qubit mask, limited-gate counts) incrementally, the legal-action list of
// a node is computed once, rewards are cached by structure list while the parameters stand, and
// single rewards go to the GPU evaluator with the parameter super-set resident on the device.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/qas_b200_mcts.h"
#include "tape.h"

namespace {

constexpr int kMaxQubits = 32;

struct Node {
  int parent = -1;
  int action = -1;
  int depth = 0;
  long long visits = 0;
  double total = 0.0;
  bool terminal = false, fully_expanded = false, restricted = true;
  uint32_t touched = 0;                       // qubit_with_actions
  int16_t last[kMaxQubits];                   // last non-PlaceHolder pool key stacked on each qubit
  std::vector<int16_t> limited_count;         // per limited gate id (index into limited_ids)
  std::vector<std::pair<int, int>> children;  // (action, node) in insertion order (Python dict order)
  std::vector<int> legal;                     // legal actions, computed at the first expansion
  bool legal_known = false;
  double mean() const { return total / (double)visits; }
};

struct KeyHash {
  size_t operator()(const std::vector<int32_t> &k) const {
    uint64_t h = 1469598103934665603ull;
    for (int32_t v : k) { h ^= (uint64_t)(uint32_t)v; h *= 1099511628211ull; }
    return (size_t)h;
  }
};

bool is_pauli_like(int g) { return g == QAS_G_PAULIX || g == QAS_G_PAULIY || g == QAS_G_PAULIZ || g == QAS_G_HADAMARD; }
int inverse_of(int g) {      // T <-> Tdg, S <-> Sdg
  switch (g) { case QAS_G_T: return QAS_G_TDG; case QAS_G_TDG: return QAS_G_T; case QAS_G_S: return QAS_G_SDG;
               case QAS_G_SDG: return QAS_G_S; default: return -1; }
}
int square_of(int g) {       // T.T = S, Tdg.Tdg = Sdg, S.S = Sdg.Sdg = Z
  switch (g) { case QAS_G_T: return QAS_G_S; case QAS_G_TDG: return QAS_G_SDG; case QAS_G_S: case QAS_G_SDG: return QAS_G_PAULIZ;
               default: return -1; }
}

}  // namespace

struct qas_mcts {
  qas_mcts_config cfg{};
  std::vector<qas_pool_entry> pool;
  int c = 0;
  std::vector<int> limited_ids;            // gate ids with a limit
  int limited_index[QAS_G_COUNT];
  bool single_present[QAS_G_COUNT];        // gate id appears as a one-qubit pool entry
  qas_energy_fn energy_fn = nullptr;
  void *user = nullptr;
  qas_ctx *ctx = nullptr;
  double reward_sign = -1.0;
  std::vector<Node> nodes;
  int root = 0;
  int prune_counter = 0;
  std::mt19937_64 rng;
  std::unordered_map<std::vector<int32_t>, double, KeyHash> cache;
  uint64_t reward_requests = 0, reward_evaluations = 0, expansions = 0, reward_batches = 0;
  // leaf-parallel rounds (opt-in, qas_mcts_set_leaf_parallel): selection rounds are issued `leaf_parallel` at a
  // time under a virtual loss and rewarded by ONE batched evaluation
  int leaf_parallel = 1;
  bool vl_fixed = false;
  double virtual_loss = 0.0, min_reward = 0.0;
  mutable std::string err;

  int fail(int code, const std::string &msg) const { err = msg; return code; }

  // ---- state bookkeeping ---------------------------------------------------------------------
  Node make_root() const {
    Node n;
    n.restricted = cfg.restricted_root != 0;
    n.touched = cfg.init_touched;
    for (int q = 0; q < kMaxQubits; ++q) n.last[q] = -1;
    n.limited_count.assign(limited_ids.size(), 0);
    n.terminal = cfg.max_depth <= 0;
    n.fully_expanded = n.terminal;
    return n;
  }
  // verifyDesirableAction (qml_mcts_state.py:58-134)
  bool desirable(const Node &n, int action) const {
    const qas_pool_entry &e = pool[action];
    const int g = e.gate_id, npar = qas_gate_nparams(g);
    const bool two = qas_gate_nwires(g) == 2;
    if (!two) {
      const int prev = n.last[e.wire0];
      if (npar > 0 || is_pauli_like(g)) {
        if (prev == action) return false;
      } else if (inverse_of(g) >= 0 && prev >= 0) {
        if (pool[prev].gate_id == inverse_of(g)) return false;
        if (prev == action && single_present[square_of(g)]) return false;
      }
    } else {
      if (n.last[e.wire0] == action && n.last[e.wire1] == action)
        if (npar > 0 || g == QAS_G_CNOT || g == QAS_G_CZ || g == QAS_G_CY) return false;
      if (!((n.touched >> e.wire0) & 1u)) return false;
    }
    if (g == QAS_G_PLACEHOLDER && !((n.touched >> e.wire0) & 1u)) return false;
    const int li = limited_index[g];
    if (li >= 0 && n.limited_count[li] >= cfg.gate_limit[g]) return false;
    return true;
  }
  void legal_actions(Node &n) const {
    if (n.legal_known) return;
    n.legal.clear();
    if (n.depth < cfg.max_depth)
      for (int a = 0; a < c; ++a)
        if (!n.restricted || desirable(n, a)) n.legal.push_back(a);
    n.legal_known = true;
  }
  // takeAction (qml_mcts_state.py:136-148); children of any node are restricted states
  Node child_of(const Node &par, int parent_index, int action) const {
    Node ch;
    ch.parent = parent_index;
    ch.action = action;
    ch.depth = par.depth + 1;
    ch.restricted = true;
    ch.touched = par.touched;
    std::memcpy(ch.last, par.last, sizeof ch.last);
    ch.limited_count = par.limited_count;
    const qas_pool_entry &e = pool[action];
    const bool two = qas_gate_nwires(e.gate_id) == 2;
    if (e.gate_id != QAS_G_PLACEHOLDER) {
      ch.touched |= 1u << (two ? e.wire1 : e.wire0);
      ch.last[e.wire0] = (int16_t)action;
      if (two) ch.last[e.wire1] = (int16_t)action;
    }
    const int li = limited_index[e.gate_id];
    if (li >= 0) ch.limited_count[li]++;
    ch.terminal = ch.depth >= cfg.max_depth;
    ch.fully_expanded = ch.terminal;
    return ch;
  }
  std::vector<int32_t> path_of(int node) const {
    std::vector<int32_t> k(nodes[node].depth);
    for (int i = node; nodes[i].parent >= 0; i = nodes[i].parent) k[nodes[i].depth - 1] = nodes[i].action;
    return k;
  }
  int placeholders(const std::vector<int32_t> &k) const {
    int n = 0;
    for (int32_t a : k) n += pool[a].gate_id == QAS_G_PLACEHOLDER;
    return n;
  }
  double penalised(double r, const std::vector<int32_t> &k) const {
    if (cfg.penalty_kind == QAS_PENALTY_PLACEHOLDER) {
      const int ph = placeholders(k);
      if (ph >= cfg.penalty_limit) return r - cfg.penalty_scale * (double)(ph - cfg.penalty_limit);
    }
    return r;
  }

  // ---- tree policy ------------------------------------------------------------------------------
  void reset() {
    nodes.clear();
    nodes.push_back(make_root());
    root = 0;
    prune_counter = 0;
  }
  size_t rand_below(size_t n) { return (size_t)(rng() % (uint64_t)n); }

  // expand (mcts.py:83-92): a uniformly random legal action that has no child yet
  int expand(int ni) {
    legal_actions(nodes[ni]);
    Node &n = nodes[ni];
    std::vector<int> open;
    for (int a : n.legal) {
      bool has = false;
      for (auto &ch : n.children) if (ch.first == a) { has = true; break; }
      if (!has) open.push_back(a);
    }
    if (open.empty()) return -1;
    const int a = open[rand_below(open.size())];
    Node ch = child_of(n, ni, a);
    const int ci = (int)nodes.size();
    nodes.push_back(std::move(ch));          // may reallocate: do not touch `n` below
    nodes[ni].children.push_back({a, ci});
    if (nodes[ni].legal.size() == nodes[ni].children.size()) nodes[ni].fully_expanded = true;
    ++expansions;
    return ci;
  }
  void back_propagate(int ni, double reward) {
    for (int i = ni; i >= 0; i = nodes[i].parent) { nodes[i].visits += 1; nodes[i].total += reward; }
  }
  // getBestChild (mcts.py:125-151)
  int best_child(int ni, double alpha, int policy) {
    const Node &n = nodes[ni];
    if (n.children.empty()) return -1;
    if (policy == QAS_POLICY_RANDOM) return n.children[rand_below(n.children.size())].second;
    const double bonus = 2.0 * std::log((double)n.visits);
    std::vector<int> lead, runner_up;
    double top = -std::numeric_limits<double>::infinity();
    for (auto &ch : n.children) {
      const Node &cn = nodes[ch.second];
      const double score = cn.total / (double)cn.visits + alpha * std::sqrt(bonus / (double)cn.visits);
      if (score > top) {
        if (top == -std::numeric_limits<double>::infinity()) runner_up = {ch.second}; else runner_up = lead;
        lead = {ch.second};
        top = score;
      } else if (score == top) {
        lead.push_back(ch.second);
      }
    }
    const std::vector<int> &pick = policy == QAS_POLICY_LOCAL_OPTIMAL ? lead : runner_up;
    if (pick.empty()) return n.children[0].second;      // all scores NaN: keep going
    return pick[rand_below(pick.size())];
  }
  // _prune (mcts.py:153-174)
  void prune(int ni) {
    if (!cfg.prune) return;
    Node &n = nodes[ni];
    const double cut = n.mean() * cfg.prune_reward_ratio;
    std::vector<std::pair<double, int>> doomed;      // (mean, action)
    for (auto &ch : n.children) {
      const Node &cn = nodes[ch.second];
      if (cn.mean() < cut && cn.visits > cfg.max_visits_prune_threshold) doomed.push_back({cn.mean(), ch.first});
    }
    std::stable_sort(doomed.begin(), doomed.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first < b.first; });
    const int keep_from = (int)doomed.size() - cfg.min_num_children;
    for (int rank = 0; rank < (int)doomed.size(); ++rank) {
      if (rank > keep_from) break;
      const int act = doomed[rank].second;
      for (size_t i = 0; i < n.children.size(); ++i)
        if (n.children[i].first == act) { n.children.erase(n.children.begin() + i); break; }
      ++prune_counter;
    }
  }
  int select_node(int ni) {
    if (!nodes[ni].fully_expanded) return expand(ni);
    prune(ni);
    return best_child(ni, cfg.alpha, cfg.sample_policy);
  }

  // ---- rewards ------------------------------------------------------------------------------------
  int energies(const int32_t *ks, int B, double *out) {
    if (ctx) {
      const int rc = qas_eval_batch(ctx, B, cfg.max_depth, ks, nullptr, out, nullptr, nullptr, QAS_F_PARAMS_RESIDENT, nullptr);
      if (rc != QAS_OK) return fail(rc, std::string("evaluator: ") + qas_last_error(ctx));
      return QAS_OK;
    }
    if (!energy_fn) return fail(QAS_ERR_INVALID, "no reward source: pass an energy callback or bind an evaluator");
    if (energy_fn(user, ks, B, cfg.max_depth, out) != 0) return fail(QAS_ERR_INVALID, "energy callback failed");
    return QAS_OK;
  }
  // simulationWithSuperCircuitParamsAndK (mcts.py:100-105) behind the per-parameter cache
  int reward_of(const std::vector<int32_t> &k, double *r) {
    ++reward_requests;
    auto it = cache.find(k);
    if (it != cache.end()) { *r = it->second; return QAS_OK; }
    double e = 0.0;
    const int rc = energies(k.data(), 1, &e);
    if (rc != QAS_OK) return rc;
    ++reward_evaluations;
    *r = reward_sign * e;
    cache.emplace(k, *r);
    return QAS_OK;
  }
  // executeRoundWithSuperCircParamsFromAnyNode (mcts.py:176-184)
  int execute_round(int from) {
    int ni = from;
    while (!nodes[ni].terminal) {
      ni = select_node(ni);
      if (ni < 0) return fail(QAS_ERR_INVALID, "no legal child: the legality rules left a node without actions");
    }
    const std::vector<int32_t> k = path_of(ni);
    double r = 0.0;
    const int rc = reward_of(k, &r);
    if (rc != QAS_OK) return rc;
    back_propagate(ni, penalised(r, k));
    return QAS_OK;
  }
  // rewards of many arcs: the ones the cache does not hold go to the evaluator in ONE batch
  int rewards_of(const std::vector<std::vector<int32_t>> &ks, std::vector<double> &out) {
    const int p = cfg.max_depth;
    std::vector<int32_t> need;
    std::vector<const std::vector<int32_t> *> need_keys;
    std::unordered_map<std::vector<int32_t>, int, KeyHash> pending;
    for (const auto &k : ks) {
      ++reward_requests;
      if (cache.count(k) || pending.count(k)) continue;
      pending.emplace(k, (int)need_keys.size());
      need.insert(need.end(), k.begin(), k.end());
      need_keys.push_back(&k);
    }
    if (!need_keys.empty()) {
      std::vector<double> e(need_keys.size());
      const int rc = energies(need.data(), (int)need_keys.size(), e.data());
      if (rc != QAS_OK) return rc;
      (void)p;
      reward_evaluations += need_keys.size();
      ++reward_batches;
      for (size_t i = 0; i < need_keys.size(); ++i) cache.emplace(*need_keys[i], reward_sign * e[i]);
    }
    out.resize(ks.size());
    for (size_t i = 0; i < ks.size(); ++i) out[i] = cache[ks[i]];
    return QAS_OK;
  }
  // `rounds` rounds of executeRoundWithSuperCircParamsFromAnyNode from `from` (mcts.py:176-184, the loops at
  // :187-188 and :198-199).  leaf_parallel > 1: in waves -- every selected leaf is booked at once with a virtual
  // loss (one visit, reward = virtual_loss), so the following selections of the wave steer away from it; the
  // wave's rewards come from one batched evaluation and replace the virtual losses.
  int run_rounds(int from, int rounds) {
    if (leaf_parallel <= 1) {
      for (int r = 0; r < rounds; ++r) { const int rc = execute_round(from); if (rc != QAS_OK) return rc; }
      return QAS_OK;
    }
    std::vector<int> leaves;
    std::vector<std::vector<int32_t>> ks;
    std::vector<double> rs;
    for (int done = 0; done < rounds;) {
      const int wave = std::min(leaf_parallel, rounds - done);
      const double vl = vl_fixed ? virtual_loss : min_reward;
      leaves.clear();
      ks.clear();
      for (int w = 0; w < wave; ++w) {
        int ni = from;
        while (!nodes[ni].terminal) {
          ni = select_node(ni);
          if (ni < 0) return fail(QAS_ERR_INVALID, "no legal child: the legality rules left a node without actions");
        }
        back_propagate(ni, vl);
        leaves.push_back(ni);
        ks.push_back(path_of(ni));
      }
      const int rc = rewards_of(ks, rs);
      if (rc != QAS_OK) return rc;
      for (int w = 0; w < wave; ++w) {
        const double real = penalised(rs[w], ks[w]);
        min_reward = std::min(min_reward, real);
        for (int i = leaves[w]; i >= 0; i = nodes[i].parent) nodes[i].total += real - vl;
      }
      done += wave;
    }
    return QAS_OK;
  }
};

// ------------------------------------------------------------------------------------- C ABI
static thread_local std::string g_mcts_create_error;

extern "C" const char *qas_mcts_last_error(const qas_mcts *m) { return m ? m->err.c_str() : g_mcts_create_error.c_str(); }

extern "C" int qas_mcts_create(qas_mcts **out, const qas_mcts_config *cfg, const qas_pool_entry *pool, int c,
                               qas_energy_fn energy_fn, void *user, double reward_sign) {
  if (!out || !cfg || !pool || c <= 0 || c > 32767) { g_mcts_create_error = "bad arguments to qas_mcts_create"; return QAS_ERR_INVALID; }
  *out = nullptr;
  if (cfg->num_qubits < 1 || cfg->num_qubits > kMaxQubits || cfg->max_depth < 1) { g_mcts_create_error = "num_qubits / max_depth out of range"; return QAS_ERR_INVALID; }
  for (int i = 0; i < c; ++i) {
    const int g = pool[i].gate_id;
    if (g < 0 || g >= QAS_G_COUNT || pool[i].wire0 < 0 || pool[i].wire0 >= cfg->num_qubits ||
        (qas_gate_nwires(g) == 2 && (pool[i].wire1 < 0 || pool[i].wire1 >= cfg->num_qubits))) {
      g_mcts_create_error = "pool entry " + std::to_string(i) + " invalid"; return QAS_ERR_INVALID;
    }
  }
  qas_mcts *m = new qas_mcts();
  m->cfg = *cfg;
  m->pool.assign(pool, pool + c);
  m->c = c;
  for (int g = 0; g < QAS_G_COUNT; ++g) {
    m->limited_index[g] = -1;
    m->single_present[g] = false;
    if (cfg->gate_limit[g] >= 0) { m->limited_index[g] = (int)m->limited_ids.size(); m->limited_ids.push_back(g); }
  }
  for (int i = 0; i < c; ++i) if (qas_gate_nwires(pool[i].gate_id) == 1) m->single_present[pool[i].gate_id] = true;
  m->energy_fn = energy_fn;
  m->user = user;
  m->reward_sign = reward_sign;
  m->rng.seed(cfg->seed);
  m->reset();
  *out = m;
  return QAS_OK;
}

extern "C" int qas_mcts_destroy(qas_mcts *m) { delete m; return QAS_OK; }

extern "C" int qas_mcts_bind_evaluator(qas_mcts *m, qas_ctx *ctx) {
  if (!m) return QAS_ERR_INVALID;
  m->ctx = ctx;
  m->cache.clear();
  return QAS_OK;
}

extern "C" int qas_mcts_set_params(qas_mcts *m, const double *params, int p, int c, int l) {
  if (!m) return QAS_ERR_INVALID;
  if (p != m->cfg.max_depth || c != m->c) return m->fail(QAS_ERR_INVALID, "parameter super-set shape does not match (max_depth, pool size)");
  m->cache.clear();
  if (m->ctx) {
    if (!params) return m->fail(QAS_ERR_INVALID, "params is null");
    if (l != qas_ctx_param_slots(m->ctx))      // the evaluator reads params with its own stride
      return m->fail(QAS_ERR_INVALID, "parameter slots per gate (l) differ from the evaluator context's");
    const int rc = qas_set_params(m->ctx, p, params);
    if (rc != QAS_OK) return m->fail(rc, std::string("evaluator: ") + qas_last_error(m->ctx));
  }
  return QAS_OK;
}

extern "C" int qas_mcts_reset(qas_mcts *m) { if (!m) return QAS_ERR_INVALID; m->reset(); return QAS_OK; }
extern "C" int qas_mcts_set_alpha(qas_mcts *m, double alpha) { if (!m) return QAS_ERR_INVALID; m->cfg.alpha = alpha; return QAS_OK; }
extern "C" int qas_mcts_set_leaf_parallel(qas_mcts *m, int leaves, int fixed_virtual_loss, double virtual_loss) {
  if (!m || leaves < 1) return QAS_ERR_INVALID;
  m->leaf_parallel = leaves;
  m->vl_fixed = fixed_virtual_loss != 0;
  m->virtual_loss = virtual_loss;
  return QAS_OK;
}
extern "C" int qas_mcts_set_prune_ratio(qas_mcts *m, double ratio) { if (!m) return QAS_ERR_INVALID; m->cfg.prune_reward_ratio = ratio; return QAS_OK; }

extern "C" int qas_mcts_warmup_batch(qas_mcts *m, int B, int32_t *out_ks, double *out_rewards) {
  if (!m || B <= 0 || !out_ks || !out_rewards) return QAS_ERR_INVALID;
  const int p = m->cfg.max_depth;
  std::vector<int> leaves(B);
  for (int b = 0; b < B; ++b) {       // randomSample (mcts.py:107-114)
    int ni = m->root;
    while (!m->nodes[ni].terminal) {
      if (!m->nodes[ni].fully_expanded) ni = m->expand(ni);
      else {
        const auto &ch = m->nodes[ni].children;
        ni = ch.empty() ? -1 : ch[m->rand_below(ch.size())].second;
      }
      if (ni < 0) return m->fail(QAS_ERR_INVALID, "no legal child during randomSample");
    }
    leaves[b] = ni;
    const std::vector<int32_t> k = m->path_of(ni);
    std::memcpy(out_ks + (size_t)b * p, k.data(), sizeof(int32_t) * p);
  }
  // one batched evaluation of the arcs that are not cached yet
  std::vector<int32_t> need;
  std::vector<std::vector<int32_t>> need_keys;
  std::unordered_map<std::vector<int32_t>, int, KeyHash> pending;
  for (int b = 0; b < B; ++b) {
    std::vector<int32_t> k(out_ks + (size_t)b * p, out_ks + (size_t)(b + 1) * p);
    ++m->reward_requests;
    if (m->cache.count(k) || pending.count(k)) continue;
    pending.emplace(k, (int)need_keys.size());
    need.insert(need.end(), k.begin(), k.end());
    need_keys.push_back(std::move(k));
  }
  if (!need_keys.empty()) {
    std::vector<double> e(need_keys.size());
    const int rc = m->energies(need.data(), (int)need_keys.size(), e.data());
    if (rc != QAS_OK) return rc;
    m->reward_evaluations += need_keys.size();
    for (size_t i = 0; i < need_keys.size(); ++i) m->cache.emplace(need_keys[i], m->reward_sign * e[i]);
  }
  for (int b = 0; b < B; ++b) {
    std::vector<int32_t> k(out_ks + (size_t)b * p, out_ks + (size_t)(b + 1) * p);
    const double r = m->cache[k];
    out_rewards[b] = r;
    m->back_propagate(leaves[b], m->penalised(r, k));
  }
  return QAS_OK;
}

extern "C" int qas_mcts_sample_arc(qas_mcts *m, int32_t *out_k, double *out_reward) {
  if (!m || !out_k) return QAS_ERR_INVALID;
  {                                                              // sampleArcWithSuperCircParams (mcts.py:186-193)
    const int rc = m->run_rounds(m->root, m->cfg.sampling_execute_rounds);
    if (rc != QAS_OK) return rc;
  }
  int ni = m->root;
  while (!m->nodes[ni].terminal) {
    ni = m->best_child(ni, m->cfg.alpha, m->cfg.sample_policy);
    if (ni < 0) return m->fail(QAS_ERR_INVALID, "No Legal Child for Node during the greedy descent");
  }
  const std::vector<int32_t> k = m->path_of(ni);
  std::memcpy(out_k, k.data(), sizeof(int32_t) * k.size());
  double r = 0.0;                                                // mcts.py:336-338
  const int rc = m->reward_of(k, &r);
  if (rc != QAS_OK) return rc;
  m->back_propagate(ni, m->penalised(r, k));
  if (out_reward) *out_reward = r;
  return QAS_OK;
}

extern "C" int qas_mcts_exploit_arc(qas_mcts *m, int32_t *out_k, double *out_reward, double *out_penalised) {
  if (!m || !out_k) return QAS_ERR_INVALID;
  int ni = m->root;
  while (!m->nodes[ni].terminal) {                               // exploitArcWithSuperCircParams (mcts.py:195-201)
    {
      const int rc = m->run_rounds(ni, m->cfg.exploit_execute_rounds);
      if (rc != QAS_OK) return rc;
    }
    ni = m->best_child(ni, 0.0, m->cfg.exploit_policy);
    if (ni < 0) return m->fail(QAS_ERR_INVALID, "No Legal Child for Node during exploitation");
  }
  const std::vector<int32_t> k = m->path_of(ni);
  std::memcpy(out_k, k.data(), sizeof(int32_t) * k.size());
  double r = 0.0;
  const int rc = m->reward_of(k, &r);
  if (rc != QAS_OK) return rc;
  if (out_reward) *out_reward = r;
  if (out_penalised) *out_penalised = m->penalised(r, k);
  return QAS_OK;
}

extern "C" int qas_mcts_legal_actions(qas_mcts *m, const int32_t *k, int depth, uint8_t *legal) {
  if (!m || depth < 0 || depth > m->cfg.max_depth || (depth > 0 && !k) || !legal) return QAS_ERR_INVALID;
  Node n = m->make_root();
  for (int i = 0; i < depth; ++i) {
    if (k[i] < 0 || k[i] >= m->c) return m->fail(QAS_ERR_INVALID, "structure list entry out of range");
    n = m->child_of(n, -1, k[i]);
  }
  m->legal_actions(n);
  std::memset(legal, 0, (size_t)m->c);
  for (int a : n.legal) legal[a] = 1;
  return QAS_OK;
}

extern "C" int qas_mcts_get_stats(const qas_mcts *m, qas_mcts_stats *out) {
  if (!m || !out) return QAS_ERR_INVALID;
  out->nodes = m->nodes.size();
  out->reward_requests = m->reward_requests;
  out->reward_evaluations = m->reward_evaluations;
  out->expansions = m->expansions;
  out->prune_counter = m->prune_counter;
  out->root_visits = (int32_t)m->nodes[m->root].visits;
  out->reward_batches = m->reward_batches;
  return QAS_OK;
}
