/* qas_b200_sv.h -- C ABI of the large-state mode of libqas_b200.so.
 *
 * One complex128 statevector resident in HBM (one slab of 2^n_local amplitudes per GPU; the
 * top log2(world) qubits are the rank index and are exchanged by the caller with an NCCL
 * all-to-all, see qas_b200/largestate.py).  The reference has no counterpart: its largest
 * model is the 16-qubit kagome lattice on one CPU (qas/qml_models/kagome_heisenberg_model.py:62);
 * this mode covers BASELINE.json's "synthetic 26-30-qubit hardware-efficient ansatz from the
 * qas gate set" (Rot / CNOT layers = the gates of qas/qml_models/qml_gate_ops.py:66-105).
 *
 * All gates here are already expressed on LOCAL bit positions (bit b of the local amplitude
 * index); mapping logical wires to positions and resolving controls that sit on rank bits is
 * the caller's job.  `state` is a caller-owned device pointer to 2^n_local double2.
 */
#ifndef QAS_B200_SV_H
#define QAS_B200_SV_H

#include <stdint.h>
#include "qas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QAS_SV_U1 0      /* 2x2 unitary m on bit `target`, optionally controlled on bit `control` */
#define QAS_SV_PDIAG 1   /* amplitude *= (parity(i & mask) ? d1 : d0), d0 = m[0..1], d1 = m[6..7]  */

typedef struct {
  int32_t kind;
  int32_t target;     /* local bit position (QAS_SV_U1) */
  int32_t control;    /* local bit position, or -1 */
  int32_t reserved;
  uint64_t mask;      /* QAS_SV_PDIAG: parity mask over local bits */
  double m[8];        /* U00 U01 U10 U11 as (re, im) pairs */
} qas_sv_gate;

typedef struct {
  int passes;                 /* HBM passes (kernel launches) the gate list was blocked into */
  int tile_bits;              /* log2 amplitudes per shared-memory tile */
  double bytes_per_pass;      /* 32 * 2^n_local (read + write) */
  double last_ms;             /* device time of the whole call (CUDA events), host-sync calls only */
  int tma_passes;             /* passes whose tiles moved by cp.async.bulk.tensor (TMA); the others used cp.async */
} qas_sv_info;

/* |state> <- amp * |index>  (index < 0: every amplitude = amp, i.e. the uniform state) */
int qas_sv_init(void *state, int n_local, int64_t index, double amp_re, double amp_im, void *stream);

/* Apply `ngates` gates (host array) in order.  The call blocks them into passes: each pass
 * stages tiles of 2^tile_bits amplitudes (tile_bits <= 13; 0 = default) in shared memory,
 * applies every gate of the block and writes the tile back -- one HBM read+write per pass.
 * Tiles move by TMA (cp.async.bulk.tensor, completion on an mbarrier) whenever the tile bits can
 * be described by a 5-d tensor map of the slab, else by cp.async.  The work is enqueued on
 * `stream`, and the call returns after it has completed (it synchronises the stream: the
 * per-pass gate lists are staged through module-level constant memory, which also makes
 * concurrent calls serialise on an internal mutex); with info != NULL and sync != 0 the device
 * time of the call is measured with CUDA events and returned in info->last_ms.               */
int qas_sv_apply(void *state, int n_local, const qas_sv_gate *gates, int ngates, int tile_bits,
                 void *stream, int sync, qas_sv_info *info);

/* Host-only view of the pass planner behind qas_sv_apply (no GPU needed): block_of_gate[i] = the
 * pass gate i runs in; block_bits[b] = bit mask of the tile bits of pass b.                   */
int qas_sv_plan(int n_local, const qas_sv_gate *gates, int ngates, int tile_bits,
                int *block_of_gate, int *num_blocks, uint64_t *block_bits, int max_blocks);

/* partial <psi|H|psi> of this slab.  Terms are given over the FULL index (n_total bits); bits
 * >= n_local are rank bits whose value is `rank_bits` -- they may only carry Z (xmask must be
 * local).  Result (a double) is written to *out_host after a stream sync.                    */
int qas_sv_expval(const void *state, int n_local, int n_total, uint64_t rank_bits,
                  const qas_pauli_term *terms, int num_terms, double *out_host, void *stream);

/* squared norm of the slab (diagnostic) */
int qas_sv_norm2(const void *state, int n_local, double *out_host, void *stream);

const char *qas_sv_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
