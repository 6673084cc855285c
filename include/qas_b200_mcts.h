/* qas_b200_mcts.h -- C ABI of the native tree policy of the nested MCTS (libqas_b200.so).
 *
 * With the evaluator on the GPU a search epoch is bound by the reference's Python tree policy:
 * MCTSController (qas/mcts/mcts.py:32-201: expand :83-92, backPropagate :94-98, randomSample
 * :107-114, getBestChild / UCT :125-151, pruning :153-174, executeRound :176-184, sampleArc
 * :186-193, exploitArc :195-201) and the legality rules of QMLStateBasicGates
 * (qas/mcts/qml_mcts_state.py:58-148).  This is a behavioural restatement of both in C++ behind a
 * handle.  It draws from its own random generator (std::mt19937_64), so it follows the same
 * policy as the Python controller, not the same random trajectory.
 *
 * Rewards come from a callback (any evaluator; used by the CPU test tier) or, after
 * qas_mcts_bind_evaluator, straight from a qas_ctx with the parameter super-set resident on the
 * GPU.  Rewards are cached by structure list until the parameters change, exactly like the
 * Python controller's per-epoch cache (valid: parameters are constant inside an epoch,
 * mcts.py:284-351).
 */
#ifndef QAS_B200_MCTS_H
#define QAS_B200_MCTS_H

#include <stdint.h>
#include "qas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qas_mcts qas_mcts;

/* child policies of getBestChild (mcts.py:125-151) */
#define QAS_POLICY_LOCAL_OPTIMAL 0
#define QAS_POLICY_LOCAL_SUB_OPTIMAL 1
#define QAS_POLICY_RANDOM 2

/* reward penalties the experiment scripts use (h2_sto-3g.py:55-64, search_H2O_near_cx.py:80-89):
 *   r - scale * (ph_count - limit)  when the arc holds ph_count >= limit PlaceHolders          */
#define QAS_PENALTY_NONE 0
#define QAS_PENALTY_PLACEHOLDER 1

typedef struct {
  int32_t max_depth;                 /* p */
  int32_t num_qubits;
  double alpha;                      /* UCT exploration weight (mutable: qas_mcts_set_alpha)     */
  double prune_reward_ratio;         /* mutable: qas_mcts_set_prune_ratio                        */
  int32_t max_visits_prune_threshold;
  int32_t min_num_children;
  int32_t sampling_execute_rounds;
  int32_t exploit_execute_rounds;
  int32_t sample_policy;             /* QAS_POLICY_* used while sampling (first policy)          */
  int32_t exploit_policy;            /* QAS_POLICY_* used by exploitArc (second policy)          */
  int32_t prune;                     /* 0/1                                                      */
  int32_t restricted_root;           /* 1 = QMLStateBasicGates, 0 = ...NoRestrictions (root only) */
  uint32_t init_touched;             /* bit q set = qubit q is in init_qubit_with_controls       */
  int32_t penalty_kind;              /* QAS_PENALTY_*                                            */
  int32_t penalty_limit;
  double penalty_scale;
  int32_t gate_limit[QAS_G_COUNT];   /* per gate id, -1 = unlimited (gate_limit_dict)            */
  uint64_t seed;
} qas_mcts_config;

/* energies[b] = E(k_b, current parameters) for B structure lists of max_depth keys each;
 * return 0 on success.  The reward of an arc is reward_sign * energy.                          */
typedef int (*qas_energy_fn)(void *user, const int32_t *ks, int B, int p, double *energies);

int qas_mcts_create(qas_mcts **out, const qas_mcts_config *cfg, const qas_pool_entry *pool, int c,
                    qas_energy_fn energy_fn, void *user, double reward_sign);
int qas_mcts_destroy(qas_mcts *m);
const char *qas_mcts_last_error(const qas_mcts *m);

/* Route rewards to a GPU evaluator: qas_mcts_set_params uploads the parameter super-set once,
 * builds the gate table once, and every reward is then one k upload + kernels + 8-byte readback. */
int qas_mcts_bind_evaluator(qas_mcts *m, qas_ctx *ctx);
/* New parameters: clears the reward cache (mcts.py: every reward is a function of params).      */
int qas_mcts_set_params(qas_mcts *m, const double *params, int p, int c, int l);

int qas_mcts_reset(qas_mcts *m);                       /* _reset: fresh root, prune counter = 0   */
int qas_mcts_set_alpha(qas_mcts *m, double alpha);
int qas_mcts_set_prune_ratio(qas_mcts *m, double ratio);
/* Leaf-parallel rounds (opt-in, not semantics-preserving): the selection + simulation
 * rounds of sampleArc / exploitArc (mcts.py:176-184,187-188,198-199) are issued `leaves` at a time; every
 * selected leaf is booked at once with one visit and a virtual loss (fixed_virtual_loss != 0: `virtual_loss`,
 * else the smallest penalised reward seen so far), the wave's rewards come from ONE batched evaluation, and
 * the virtual losses are then replaced by the real rewards.  leaves = 1 (default) = the reference's sequential
 * rounds.                                                                                         */
int qas_mcts_set_leaf_parallel(qas_mcts *m, int leaves, int fixed_virtual_loss, double virtual_loss);

/* warm-up epoch (mcts.py:298-305): draw B arcs with randomSample, evaluate them in ONE batch,
 * back-propagate the (penalised) rewards.  out_ks: int32 [B][p]; out_rewards: double [B] raw.   */
int qas_mcts_warmup_batch(qas_mcts *m, int B, int32_t *out_ks, double *out_rewards);
/* search epoch body (mcts.py:332-338): sampleArcWithSuperCircParams, then the arc's reward once
 * more, back-propagated from its leaf.  out_k: int32 [p].                                       */
int qas_mcts_sample_arc(qas_mcts *m, int32_t *out_k, double *out_reward);
/* exploitArcWithSuperCircParams (mcts.py:195-201).  out_k: int32 [p]; out_reward raw reward.     */
int qas_mcts_exploit_arc(qas_mcts *m, int32_t *out_k, double *out_reward, double *out_penalised);

/* legality of every pool key after the partial structure list k[0..depth) (test hook; restates
 * QMLStateBasicGates.getLegalActions).  legal: uint8 [c].                                        */
int qas_mcts_legal_actions(qas_mcts *m, const int32_t *k, int depth, uint8_t *legal);

typedef struct {
  uint64_t nodes;               /* nodes in the current tree                                     */
  uint64_t reward_requests;     /* rewards asked for since creation                              */
  uint64_t reward_evaluations;  /* ... of which were not served by the cache                     */
  uint64_t expansions;
  int32_t prune_counter;
  int32_t root_visits;
  uint64_t reward_batches;      /* batched evaluator calls issued by leaf-parallel waves / warm-up   */
} qas_mcts_stats;
int qas_mcts_get_stats(const qas_mcts *m, qas_mcts_stats *out);

#ifdef __cplusplus
}
#endif
#endif
