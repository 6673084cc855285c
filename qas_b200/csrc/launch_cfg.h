// launch_cfg.h -- team / CTA configuration of the classic evaluation kernels by qubit count
#pragma once
#include "ctx.h"

// threads per team and threads per CTA by qubit count (n <= 5 normally run the thread-per-circuit
// kernel).  n = 10, 11: two-warp / four-warp teams in 128-thread CTAs -- six / three circuits in flight
// per SM instead of four / two, and less waiting inside a team for the warp that owns the in-support
// cosets of the adjoint sweep (LiH-10q: 2.46 -> 2.23 ms against 128-thread teams in 256-thread CTAs).
constexpr int team_threads(int n) {
  return n <= 4 ? 8 : n == 5 ? 16 : n <= 8 ? 32 : n <= 10 ? 64 : n == 11 ? 128 : 256;
}
constexpr int team_blk(int n) { return (n == 10 || n == 11) ? 128 : 256; }
constexpr int qas_gmax(int n, int T) { return qas_block_bits(n, T) >= 2 ? 2 : 1; }
// H application: 4 amplitudes per thread and a 2-deep prefetch ring where the team has the
// amplitudes for it (measured best on LiH-10q / H2O-8q among (3,1) (3,2) (2,2) (2,4))
constexpr int qas_hrb(int n, int T) { return qas_block_bits(n, T) >= 2 ? 2 : qas_block_bits(n, T); }
constexpr int qas_hpd(int n, int T) { return qas_hrb(n, T) >= 1 ? 2 : 1; }   // PD must divide 2^RB
static inline int team_threads_for(int n, int shift) {
  const int t = team_threads(n);
  // QAS_B200_TEAM_SHIFT=1: the doubled team in a 256-thread CTA (compiled for n = 8..11, see launch_eval)
  if (shift > 0 && n >= 8 && n <= 11) return t * 2;
  if (shift < 0 && n == 10) return 32;      // QAS_B200_TEAM_SHIFT=-1: one warp per circuit, no team barriers
  return t;
}


static inline size_t sdesc_bytes(int S) { return (((size_t)S * 4) + 15) & ~(size_t)15; }
// (N, T) instantiations: the default team size of each qubit count plus tuning alternatives
// n <= 5: thread-per-circuit kernel (eval_small_kernel); QAS_B200_SMALL=0 keeps the team kernel
static inline bool use_small_kernel(const qas_ctx *ctx) { return ctx->n <= 5 && ctx->small_kernel; }
// ... and the packed tape must fit next to the statevectors (very deep circuits fall back to the team kernel)
static inline bool small_fits(const qas_ctx *ctx, int ops_cap) {
  return eval_small_smem_bytes(ctx->n, ctx->n <= 4 ? 128 : 64, true, ops_cap, ctx->c) <= 200 * 1024;
}


// Which classic kernel evaluates this context's circuits, decided ONCE per call (the tape planner, the class
// descriptors and the launch must agree on the index layout): 0 = shared-memory resident (swizzled layout;
// *T / *blk = team and CTA size), 1 = global-memory slab (plain layout), -1 = neither is compiled for this
// configuration.  Deep tapes can push a team's shared memory over the 227 KB limit: the state then moves
// to the global-memory kernel, whose instantiation assumes 2-gate passes and 4-amplitude H rows.
static inline int qas_classic_path(const qas_ctx *ctx, bool grad, int ops_cap, int *T_out, int *blk_out) {
  int T = ctx->T, blk = ctx->blk;
  // energy-only launches at n = 10: one warp per circuit (no team barriers; same GMAX / RB / PD, so tapes
  // and sign tables are shared with the gradient launches).  LiH reward-only 72 -> 84 M evals/s; the
  // gradient kernel is slower that way (1.15 -> 1.44 ms: the dense tail of the batch runs on one warp).
  if (!grad && ctx->n == 10 && ctx->team_shift == 0 && T == 64 && blk == 128) { T = 32; blk = 64; }
  const bool forced_tiled = ctx->tiled && ctx->n >= ctx->tiled_force_n && ctx->n >= 11 && ctx->n <= QAS_TP_MAXQ;
  if (ctx->n <= (grad ? 12 : 13) && !forced_tiled) {
    const size_t smem = eval_team_smem_bytes(ctx->n, T, grad, ops_cap, true, qas_stage_u(ctx->n)) * (blk / T) +
                        sdesc_bytes(ctx->C_real + ctx->C_imag);
    if (smem <= 227 * 1024) { *T_out = T; *blk_out = blk; return 0; }
  }
  if (ctx->RB != 2 || ctx->PD != 2 || ctx->GMAX != 2) return -1;
  *T_out = 256; *blk_out = 256;
  return 1;
}
