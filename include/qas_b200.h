/* qas_b200.h -- C ABI of libqas_b200.so: B200-native batched evaluator for the QAS hot path.
 *
 * The reference (peiyong-addwater/QAS) has no native code and no FFI: its hot path is the
 * Python call chain  MCTSController.simulationWithSuperCircuitParamsAndK (qas/mcts/mcts.py:100-105)
 * -> model(p,c,l,k,pool).getReward / getGradient / getLoss (qas/models.py:3-21,
 * qas/qml_models/hamiltonian_model.py:55-91) -> PennyLane QNode.  Each entry point below
 * names the reference interface it replaces.  All functions return an int status
 * (QAS_OK == 0); no exception crosses this boundary; pointers are caller-owned and never
 * retained past the call.
 *
 * Bit convention: wire w of an n-qubit register is bit (n-1-w) of a basis index / mask
 * (wire 0 = most significant bit, as in PennyLane).
 */
#ifndef QAS_B200_H
#define QAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QAS_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------ */
#define QAS_OK 0
#define QAS_ERR_INVALID 1     /* bad argument (the reference's bare asserts: utils.py:12-15,    */
                              /* hamiltonian_model.py:56-58)                                     */
#define QAS_ERR_CUDA 2        /* CUDA runtime error; see qas_last_error                          */
#define QAS_ERR_UNSUPPORTED 3 /* configuration outside what the kernels cover                    */
#define QAS_ERR_NOMEM 4

/* ---- gate ids: the gate-name set of qas/qml_models/qml_gate_ops.py:66-105 -------------- */
enum qas_gate_id {
  QAS_G_PLACEHOLDER = 0,
  QAS_G_HADAMARD, QAS_G_PAULIX, QAS_G_PAULIY, QAS_G_PAULIZ, QAS_G_S, QAS_G_SDG, QAS_G_T,
  QAS_G_TDG, QAS_G_SX,
  QAS_G_ROT, QAS_G_RX, QAS_G_RY, QAS_G_RZ, QAS_G_PHASESHIFT, QAS_G_U1, QAS_G_U2, QAS_G_U3,
  QAS_G_CNOT, QAS_G_CZ, QAS_G_CY, QAS_G_SWAP, QAS_G_ISWAP, QAS_G_SISWAP,
  QAS_G_CRX, QAS_G_CRY, QAS_G_CRZ, QAS_G_CROT, QAS_G_CPHASE,
  QAS_G_ISINGXX, QAS_G_ISINGYY, QAS_G_ISINGZZ,
  QAS_G_COUNT
};

/* One operator-pool entry: key -> {gate: wires}  (QMLPool, qml_gate_ops.py:136-185).
 * wire1 = -1 for one-wire gates; for controlled gates wire0 = control, wire1 = target.    */
typedef struct {
  int32_t gate_id;
  int32_t wire0;
  int32_t wire1;
} qas_pool_entry;

/* One Pauli term  coeff * P,  P = tensor product with X where xmask&~zmask, Z where
 * zmask&~xmask, Y where xmask&zmask  (replaces the qml.Hamiltonian objects at
 * qml_models_legacy.py:1237-1256, kagome_heisenberg_model.py:41-52, the per-edge
 * Hermitian(ZZ) of :2377-2393).                                                            */
typedef struct {
  uint64_t xmask;
  uint64_t zmask;
  double coeff;
} qas_pauli_term;

/* initial state of the register (qml.BasisState(hf) at qml_models_legacy.py:1298, the
 * Hadamard layer at :2380-2381, or nothing = |0..0>)                                       */
#define QAS_INIT_ZERO 0
#define QAS_INIT_BASIS 1
#define QAS_INIT_PLUS 2

/* flags of qas_eval_batch */
#define QAS_F_DEVICE_PTRS 0x1u /* k, params, energy, grad_* are device pointers; the call is    */
                               /* asynchronous on `stream`. Otherwise host pointers: they are    */
                               /* copied straight from / to the caller's memory with              */
                               /* cudaMemcpyAsync (truly asynchronous only for pinned memory;     */
                               /* pageable memory works and is staged by the CUDA driver) and    */
                               /* the call returns synchronously.  Two exceptions: pageable      */
                               /* int32 structure lists of large batches are narrowed to one     */
                               /* byte per key into a page-locked buffer of the context (range   */
                               /* asserts in the same pass) and uploaded from there; energies    */
                               /* going to page-locked memory are read back on a second stream   */
                               /* while the batch reduction still runs.                          */
#define QAS_F_GRAD_SUM 0x2u    /* grad_dense = sum over the batch instead of mean                */
                               /* (search(..., avg_gradients=False), mcts.py:347)                */

#define QAS_F_PARAMS_RESIDENT 0x4u /* host-pointer calls only: ignore `params`, use the super-set (and   */
                                  /* its gate table) that qas_set_params left on the device -- the   */
                                  /* one-reward-at-a-time rollouts of mcts.py:100-105,181 then cost   */
                                  /* one k upload, the kernels and an 8-byte read-back                */

#define QAS_F_K_U8 0x8u            /* `k` holds one byte per key (uint8_t[B][p]; pools of <= 256 entries): a     */
                                  /* quarter of the structure-list upload; widened to int32 on the device      */

/* Threading / stream contract: a context is NOT thread-safe and owns one set of scratch buffers (tape records,
 * gate tables, cross matrices, bin queues, reduction partials).  Use it from one thread and one stream at a
 * time: a second call may only be issued on another stream after the first one has completed (device-pointer
 * calls are asynchronous -- synchronise or chain the streams with an event).  A call without
 * QAS_F_PARAMS_RESIDENT rebuilds the gate table and so invalidates a table left by qas_set_params.  Use one
 * context per stream for concurrent work.
 * Work queues: one call launches a kernel per support bin on forked streams.  Loading the library sets
 * CUDA_DEVICE_MAX_CONNECTIONS=32 (if unset) so that those streams get their own hardware queues; a host that
 * creates its CUDA context BEFORE loading the library must export the variable itself, or the bins serialise
 * behind each other and behind NCCL's streams (measured: +9 % per LiH step under torchrun).               */
typedef struct qas_ctx qas_ctx;

/* Library / build info. */
int qas_abi_version(void);
const char *qas_build_info(void);

/* Create an evaluator for one model family: n qubits, an initial state, a Pauli-sum
 * Hamiltonian and an operator pool.  Replaces the per-evaluation construction of
 * model + qml.device + QNode (mcts.py:103, qml_models_legacy.py:1269,1295-1301).
 * l = number of parameter slots per (depth, key) (3 everywhere in the reference).          */
int qas_ctx_create(qas_ctx **out, int device, int n_qubits, int init_kind, uint64_t init_basis,
                   const qas_pauli_term *terms, int num_terms,
                   const qas_pool_entry *pool, int c, int l);
int qas_ctx_destroy(qas_ctx *ctx);

/* Human-readable description of the last error on this context (or of a failed create
 * when ctx == NULL).  The string stays valid until the next call on the same context.      */
const char *qas_last_error(const qas_ctx *ctx);

/* Evaluate a batch of B candidate circuits that share one parameter super-set.
 *   k        int32  [B][p]     structure lists (pool keys), row-major
 *   params   double [p][c][l]  shared parameter super-set (read-only)
 *   energy   double [B]        out: E_b = <psi_b|H|psi_b>  (getLoss; getReward = -E)
 *   grad_compact double [B][p][l] or NULL
 *                              out: dE_b/dparams[i][k_b[i]][j]; zero where gate k_b[i] has
 *                              fewer than j+1 parameters (getGradient's non-zero entries,
 *                              hamiltonian_model.py:75-91)
 *   grad_dense   double [p][c][l] or NULL
 *                              out: mean (or sum) over the batch of the dense gradients after
 *                              nan_to_num -- the reduction at mcts.py:345-347
 * stream: a cudaStream_t.  With QAS_F_DEVICE_PTRS the work is enqueued on exactly that stream
 * (NULL = the CUDA default stream); with host pointers NULL selects the context's own stream.
 * Replaces the serial loops mcts.py:298-305 (rewards) and :340-347 (gradients).            */
int qas_eval_batch(qas_ctx *ctx, int B, int p, const int32_t *k, const double *params,
                   double *energy, double *grad_compact, double *grad_dense, uint32_t flags,
                   void *stream);

/* Upload the parameter super-set params[p][c][l] once and build its gate table; later
 * qas_eval_batch calls with QAS_F_PARAMS_RESIDENT reuse both.  Any qas_eval_batch call without the
 * flag invalidates the resident copy.  (The reference passes the same `params` to every
 * getReward of an epoch, mcts.py:284-351.)                                                   */
int qas_set_params(qas_ctx *ctx, int p, const double *params);

/* Host-side run of the tape compiler on ONE structure list (no GPU needed): lowers k to the
 * generalised-gate tape the kernels execute (CNOT/SWAP folded into the index frame).
 * out_ops receives up to max_ops records of 5 uint32 {xor_mask, target_parity_mask,
 * control_parity_mask, table_index, meta}; *num_ops the count; *init_index the physical
 * basis index of the initial state when init_kind == QAS_INIT_BASIS.
 * Restates backboneCirc (hamiltonian_model.py:22-41).                                      */
int qas_compile_tape(const qas_ctx *ctx, int p, const int32_t *k, uint32_t *out_ops, int max_ops,
                     int *num_ops, uint32_t *init_index);

/* Same, but usable without a context/GPU (pool passed directly).                            */
int qas_compile_tape_host(int n_qubits, const qas_pool_entry *pool, int c, int p,
                          const int32_t *k, uint64_t init_basis, uint32_t *out_ops, int max_ops,
                          int *num_ops, uint32_t *init_index);

/* Pass planner (host run of the code the tape-compile kernel executes on a compiled tape).
 * (1) groups consecutive uncontrolled 2x2 ops into register-fused passes of up to gmax <= 3 gates
 *     (group sizes go into meta bits 24-25 of the first and 26-27 of the last op of a group);
 * (2) tracks the affine support  init_index ^ span(xor masks applied so far)  of the state and
 *     writes one descriptor {P0, dc, X_0 .. X_{n-1-g}} of 2 + n_qubits words per group to gdesc
 *     (tape order of the groups): coset representative t of a group is P0 ^ xor of the X_i selected
 *     by the bits of t, and only the cosets t < 2^dc can hold non-zero amplitudes after the group.
 * init_kind / init_index as in qas_compile_tape; swizzled = order the X_i for the shared-memory
 * slot layout.  gdesc may be NULL (grouping only).  No reference counterpart (PennyLane applies
 * one gate per tensordot to the dense state).                                                 */
int qas_plan_passes_host(int n_qubits, uint32_t *ops, int num_ops, int gmax, int init_kind,
                         uint32_t init_index, int swizzled, uint32_t *gdesc, int gdesc_words,
                         int *num_groups);

/* Tile-pass planner of the tiled kernel for 13..16-qubit full-register circuits (host run of the code the kernel
 * executes right before every pass; qas_b200/csrc/tape.h qas_plan_tile_pass).  ops must carry group marks
 * (qas_plan_passes_host).  Starting at tape index `start` and walking in direction dir (+1 = adjoint sweep, -1 =
 * forward sweep) it takes as many whole groups as keep span(xor masks) + the clow lowest index bits within tb
 * dimensions (at most max_ops ops), writes the frame columns col[n_qubits] (pass coordinate x <-> physical index
 * xor of the columns selected by the bits of x; columns 0..tb-1 span the tile, tb.. number the tiles) and the
 * pivot table piv[n_qubits] (highest set bit -> column), and returns the number of ops taken in *num_ops.
 * No reference counterpart (lightning.qubit applies one gate per pass over the dense state).            */
int qas_plan_tile_pass_host(int n_qubits, const uint32_t *ops, int num_ops_total, int start, int dir, int tb,
                            int clow, int max_ops, uint32_t *col, int *piv, int *num_ops);

/* Host-side range asserts on structure lists (the reference's `min(k) >= 0, max(k) < c`, utils.py:12-15), exactly the
 * vectorised passes qas_eval_batch runs on large host batches: key_bytes = 1 (uint8) or 4 (int32).  *out_of_range = 1
 * if some key is < 0 or >= c.  With narrowed != NULL (int32 keys, c <= 256) the keys are also written there as one byte
 * each (the page-locked staging of pageable lists).  No GPU involved.                                                  */
int qas_check_keys_host(const void *k, uint64_t num_keys, int key_bytes, int c, uint8_t *narrowed, int *out_of_range);

/* Coset basis of one pass group inside a tile of the tiled kernel (qas_b200/csrc/tape.h qas_tile_group_basis): g <= 2
 * gates with tile-coordinate xor masks m[g] and parity masks r[g] (tb bits); kb (>= tb words) receives a basis of the
 * common kernel of the parity masks in lowest-set-bit echelon form, ascending pivots; returns its dimension tb - g in
 * *dim.  Coset c of the group is represented by the xor of the kb[i] selected by the bits of c.                  */
int qas_tile_group_basis_host(int g, const uint32_t *m, const uint32_t *r, int tb, uint32_t *kb, int *dim);

/* Final-support descriptor of the H application (host run of the planner code the tape-compile
 * kernel executes).  W = register subspace of the H blocking: w_rank <= 3 basis vectors in reduced
 * echelon form (highest-bit pivots w_pivots ascending).  hdesc (>= 32 words) receives
 * { ubar(init_index), dbar, X_0 .. X_{nu-1}, chk_0 .. chk_{nu-dbar-1} },  nu = n_qubits - w_rank:
 * quotient coordinates (index with W's pivot bits removed after reduction by W) of a basis adapted to
 * Dbar = image of the support span of the final state; dbar = nu means dense / not restricted.
 * No reference counterpart (PennyLane applies every Hamiltonian term to the full state).        */
int qas_plan_hsupport_host(int n_qubits, const uint32_t *ops, int num_ops, int init_kind,
                           uint32_t init_index, int w_rank, const int *w_pivots,
                           const uint32_t *w_basis, uint32_t *hdesc, int hdesc_words);

/* Support-compressed tape ("c-tape"; host run of the code the tape-compile kernel executes for all-zero
 * start states).  From |0..0> a circuit only populates D = span of the xor masks of its 2x2 ops, and after
 * its first j ops only the span D_j of the first j masks.  The c-tape rewrites the compiled, grouped tape in
 * a TIME-ADAPTED basis of D (B_k = mask of the k-th op, in time order, that enlarges the span): with
 * i(t) = xor_k t_k B_k the support after j ops is the prefix t < 2^{d_j} of a 2^d-amplitude register.
 * out_ops[j] = {m', r', rc', table index, meta} with meta bits 24-27 = group marks, bits 28-31 = d_j;
 * controlled ops whose control is constant on the support are dropped.  *dim = d = dim D, or -1 when
 * d == n_qubits (nothing to compress: out_ops is the grouped classic tape), -2 when d > max_d (<= 12).
 * basis_out (n_qubits words) = B_0 .. B_{d-1}; list_out[s'] = x' | s << 16 for every X mask
 * xmasks[s] inside D (x' = its coordinates); X masks outside D connect the support to amplitudes
 * that are exactly zero and drop out of the energy and of every derivative.
 * variant 0 = the reference formulation (echelon tables in memory; also what qas_eval_batch runs on the host
 * for batches of <= 16 circuits), 1 = the formulation the tape-compile kernel runs (packed ops, reduced
 * echelon basis in registers; n_qubits <= 16): both must produce identical records.
 * No reference counterpart (PennyLane keeps the dense 2^n state for every circuit).             */
int qas_compile_ctape_host(int n_qubits, const qas_pool_entry *pool, int c, int p, const int32_t *k,
                           int gmax, const uint32_t *xmasks, int num_xmasks, int max_d,
                           uint32_t *out_ops, int max_ops, int *num_ops, int *num_groups, int *dim,
                           uint32_t *basis_out, uint32_t *list_out, int *num_list, int variant);

/* Counters since context creation: evaluations and kernel launches; device time of the
 * last call's kernels (ms, CUDA events on the launch stream; host-pointer calls only) and of
 * the last evaluation kernel alone (any call; qas_get_stats waits for that kernel).         */
typedef struct {
  uint64_t evals;
  uint64_t launches;
  double last_kernel_ms;
  double last_eval_kernel_ms;
  int teams_per_cta;
  int threads_per_team;
  int ctas_per_sm;
  int smem_bytes_per_cta;
  int path;                 /* 0 = shared-memory resident, 1 = global-memory resident */
} qas_stats;
int qas_get_stats(const qas_ctx *ctx, qas_stats *out);

/* One optimiser step on the device: the update rule of qml.AdamOptimizer, the optimiser the
 * reference hands to search() and circuitModelTuning() (qas/mcts/mcts.py:226,351-352,387,407-408),
 * preceded by the caller-side gradient conditioning of mcts.py:346,349-350 / :404-406:
 *   g = nan_to_num(grad) + noise_factor * noise ;  m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g*g ;
 *   params -= stepsize * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)
 * All pointers are device pointers to n doubles (noise may be NULL); t = 1 for the first step.
 * Asynchronous on `stream`.  PennyLane's defaults: beta1 0.9, beta2 0.99, eps 1e-8.            */
int qas_adam_step(void *params, const void *grad, void *m, void *v, const void *noise, uint64_t n,
                  int t, double stepsize, double beta1, double beta2, double eps, double noise_factor,
                  void *stream);

/* The fine-tuning step of circuitModelTuning (mcts.py:381-412) in a form a CUDA graph can replay: nothing the
 * captured work reads may live on the host, so the step counter t, the bias-corrected Adam step size and the
 * loss history are device-resident.  qas_tuning_tick (one thread): t <- t + 1, *step_dev <- stepsize *
 * sqrt(1 - beta2^t) / (1 - beta1^t), losses_dev[t - 1] <- *loss_dev.  qas_adam_step_dev: the update of
 * qas_adam_step (no noise term) with the step size read from *step_dev.  Device pointers; asynchronous on
 * `stream`.  A graph of { qas_eval_batch(QAS_F_DEVICE_PTRS) ; qas_tuning_tick ; qas_adam_step_dev } replayed
 * num_epochs times is the whole loop without a host round trip per epoch (qas_eval_batch detects the capture
 * and records no timing events).                                                                   */
int qas_tuning_tick(void *t_dev, void *step_dev, const void *loss_dev, void *losses_dev, int max_losses,
                    double stepsize, double beta1, double beta2, void *stream);
int qas_adam_step_dev(void *params, const void *grad, void *m, void *v, uint64_t n, const void *step_dev,
                      double beta1, double beta2, double eps, void *stream);

/* Batch mean of per-circuit compact gradients that already sit on the device (device pointers; asynchronous on
 * `stream`): grad_dense[p][c][l] = mean | sum over b of scatter(nan_to_num(grad_compact[b])) by k[b] -- the
 * reduction of mcts.py:345-347 with the deterministic two-stage summation qas_eval_batch uses.  For costs whose
 * per-circuit gradient is assembled outside the evaluator (VQLS).                                  */
int qas_batch_mean(qas_ctx *ctx, int B, int p, const int32_t *k, const double *grad_compact, double *grad_dense,
                   uint32_t flags, void *stream);

/* VQLS cost from its two Pauli-sum expectations (VQLSDemo, qml_models_legacy.py:1880-1951,1969-1972):
 * loss[b] = 1/2 - |N_b| / (2 n |D_b|) and, when pl > 0, the quotient rule on the two compact gradients
 * gN, gD [B][pl] (gN is overwritten with dloss).  Device pointers; asynchronous on `stream`.      */
int qas_vqls_combine(const void *N, const void *D, void *gN, const void *gD, int B, int pl, int n_qubits, void *loss,
                     void *stream);

/* Fixed-order reduction of per-shard partial results: out[x] = scale * sum_r rows[r * row_stride + x],
 * x < total (device pointers; rows are summed lane-strided in ascending r, then by a fixed shuffle tree, so
 * every rank that reduces the same gathered buffer gets bit-identical numbers).  This is the multi-GPU end of
 * the batch mean of mcts.py:345-347: each rank writes [energies | gradient sum] into one buffer, ONE
 * all-gather moves it, and this kernel turns the gathered gradient sums into the mean (scale = 1 / B_total).
 * Asynchronous on `stream`.                                                                       */
int qas_reduce_rows(const void *rows, int nrows, uint64_t row_stride, uint64_t total, double scale, void *out,
                    void *stream);

/* The same exchange over NVLink peer memory, without a collective library on the data path (csrc/exchange.cu).
 * Every rank owns a receive buffer of `world` rows plus `world` 64-bit flags (zero-initialised) that is mapped into
 * every peer's address space (CUDA IPC / torch symmetric memory; the host passes the mapped pointers).
 *   qas_peer_push_row: one kernel stores `count` doubles (even; 16-byte aligned) of the local packed row into
 *     peer_rows[r] for every r < world -- the slot of THIS rank in rank r's receive buffer, r = own rank included --
 *     and then release-stores `seq` (>= 1, increasing by one per step) into peer_flags[r].
 *   qas_peer_wait_reduce: waits on the LOCAL flags until every rank's row of step `seq` has landed (status[0], a
 *     device uint32 the host zero-initialises, becomes 1 if one has not after 5 s -- the call never hangs), then
 *     out[x] = scale * sum_r rows[r * row_stride + offset + x], x < total, in the fixed order of qas_reduce_rows
 *     (total = 0: wait only).
 * Contract: every rank issues push(s), wait_reduce(s), push(s + 1), ... on ONE stream and alternates between two
 * receive buffers (s & 1); a slot is then never overwritten before its reader is done, without acknowledgements.
 * Both calls are asynchronous on `stream`.  Replaces all_gather (rewards) + all_reduce (gradient) of the batch mean
 * of mcts.py:345-347 spread over ranks; bench.py --gpus N and qas_b200.distributed.PeerExchange use it.          */
int qas_peer_push_row(const void *row, uint64_t count, void *const *peer_rows, void *const *peer_flags, int world,
                      uint64_t seq, void *stream);
int qas_peer_wait_reduce(const void *rows, int world, uint64_t row_stride, uint64_t offset, uint64_t total, double scale,
                         void *out, const void *flags, uint64_t seq, void *status, void *stream);

/* Diagnostic (contexts created with the environment variable QAS_B200_PHASE_CLOCKS=1 only): sums,
 * over all circuits evaluated so far, of the SM clock cycles a team spent in the four phases of the
 * evaluation kernel -- out4 = {prologue, forward sweep, H application + energy, adjoint sweep}.
 * Synchronises the device.  Used by tools/phase_clocks.py to see where a circuit's latency goes.   */
/* parameter slots per gate (the l of the (p, c, l) super-set) the context was created with */
int qas_ctx_param_slots(const qas_ctx *ctx);
int qas_debug_phase_clocks(qas_ctx *ctx, uint64_t *out4, int reset);
/* ... and of the compressed kernel, per bin: out[bin * 8 + 0..3] = the four phase sums, out[bin * 8 + 4] = circuits
 * evaluated; bin i holds the circuits whose support dimension is *dmin + i (the first bin also the smaller ones). */
int qas_debug_phase_clocks_c(qas_ctx *ctx, uint64_t *out, int max_bins, int *dmin, int reset);
/* ... of the tiled kernel for 13..16-qubit registers (stats.path == 4): out8 = summed cycles of the cluster's first
 * CTA [prologue, forward sweep, H + energy, adjoint sweep], forward tile passes, adjoint tile passes, circuits, 0 */
int qas_debug_phase_clocks_t(qas_ctx *ctx, uint64_t *out8, int reset);

/* Measured FP64 roofline denominator: runs a dependent-chain-free DFMA micro-benchmark on
 * `device` and returns the best of `reps` timings in TFLOP/s (2 flop per DFMA).            */
int qas_fp64_peak_tflops(int device, int reps, double *tflops);

/* ---- large-state mode (one statevector resident in HBM, optionally sharded) ---------- */
/* see DESIGN.md section "Large-state mode"; declared in qas_b200_sv.h                       */

#ifdef __cplusplus
}
#endif
#endif /* QAS_B200_H */
