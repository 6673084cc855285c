// launch_classic.inl -- launch logic of the classic evaluation kernels (kernels.cuh), included by
// launch_classic_grad.cu and launch_classic_nograd.cu with QAS_GRAD defined to true / false so that the two
// sets of instantiations compile in parallel.
#include "ctx.h"
#include "launch_cfg.h"

// ---------------------------------------------------------------------------------- launch

// CTAs per SM the register allocation is bounded for: the 3-gate adjoint pass holds 8 + 8
// amplitudes and a 2x2 cross matrix in registers (128-register budget)
constexpr int qas_minb(int n, int T, int blk) { return blk == 128 ? 3 : (qas_block_bits(n, T) >= 2 ? 2 : 3); }

template <int N, int T, bool GRAD, int RBH, int PD, int GMAX_ = 0, int MINB_ = 0, int BLK = 256>
static cudaError_t launch_smem(qas_ctx *ctx, const EvalParams &P, cudaStream_t st) {
  constexpr int TEAMS = BLK / T;
  const size_t smem = eval_team_smem_bytes(N, T, GRAD, P.ops_cap, true, qas_stage_u(N)) * TEAMS + sdesc_bytes(P.C_real + P.C_imag);
  auto kern = eval_kernel<N, T, GRAD, (GMAX_ ? GMAX_ : qas_gmax(N, T)), RBH, PD, (MINB_ ? MINB_ : qas_minb(N, T, BLK)), BLK>;
  static QasLaunchCache cache;
  int occ = 0;
  cudaError_t e = qas_prepare_launch(cache, ctx->device, kern, BLK, smem, &occ);
  if (e != cudaSuccess) return e;
  const int grid = std::min((P.B + TEAMS - 1) / TEAMS, std::max(1, occ) * ctx->num_sms);   // persistent CTAs, queue-fed
  ctx->stats.teams_per_cta = TEAMS;
  ctx->stats.threads_per_team = T;
  ctx->stats.ctas_per_sm = occ;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 0;
  kern<<<grid, BLK, smem, st>>>(P);
  return cudaGetLastError();
}

template <bool GRAD>
[[maybe_unused]] static cudaError_t launch_gmem(qas_ctx *ctx, EvalParams &P, cudaStream_t st) {
  const size_t smem = eval_team_smem_bytes(P.n, 256, GRAD, P.ops_cap, false, true) + sdesc_bytes(P.C_real + P.C_imag);
  auto kern = eval_kernel<0, 256, GRAD, 2, 2, 2, 2>;
  static QasLaunchCache cache;
  int occ = 0;
  cudaError_t e = qas_prepare_launch(cache, ctx->device, kern, 256, smem, &occ);
  if (e != cudaSuccess) return e;
  occ = std::max(1, std::min(occ, 4));
  const size_t per_cta = ((size_t)1 << P.n) * (GRAD ? 2 : 1);
  int grid = std::min(P.B, ctx->num_sms * occ);
  while (grid > 1 && per_cta * grid * 16 > ((size_t)8 << 30)) grid /= 2;
  e = ensure(&ctx->d_gstate, &ctx->cap_gstate, per_cta * grid);
  if (e != cudaSuccess) return e;
  P.gstate = ctx->d_gstate;
  ctx->stats.teams_per_cta = 1;
  ctx->stats.threads_per_team = 256;
  ctx->stats.ctas_per_sm = occ;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 1;
  kern<<<grid, 256, smem, st>>>(P);
  return cudaGetLastError();
}

template <int N, bool GRAD>
static cudaError_t launch_small(qas_ctx *ctx, const EvalParams &P, cudaStream_t st) {
  constexpr int BLK = N <= 4 ? 128 : 64;
  const size_t smem = eval_small_smem_bytes(N, BLK, GRAD, P.ops_cap, P.c);
  auto kern = eval_small_kernel<N, GRAD, BLK>;
  static QasLaunchCache cache;
  int occ = 0;
  cudaError_t e = qas_prepare_launch(cache, ctx->device, kern, BLK, smem, &occ);
  if (e != cudaSuccess) return e;
  ctx->stats.teams_per_cta = BLK;
  ctx->stats.threads_per_team = 1;
  ctx->stats.ctas_per_sm = occ;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 2;
  kern<<<(P.B + BLK - 1) / BLK, BLK, smem, st>>>(P);
  return cudaGetLastError();
}

template <bool GRAD>
static cudaError_t launch_eval(qas_ctx *ctx, EvalParams &P, cudaStream_t st) {
  if (ctx->small_now) {
    switch (P.n) {
      case 2: return launch_small<2, GRAD>(ctx, P, st);
      case 3: return launch_small<3, GRAD>(ctx, P, st);
      case 4: return launch_small<4, GRAD>(ctx, P, st);
      default: return launch_small<5, GRAD>(ctx, P, st);
    }
  }
  int T = 0, blk = 0;
  const int path = qas_classic_path(ctx, GRAD, P.ops_cap, &T, &blk);
  if (path < 0) return cudaErrorInvalidConfiguration;
  if (path == 0) {
#define QAS_CASE(N_, T_, BLK_) \
  if (P.n == N_ && T == T_ && blk == BLK_) return launch_smem<N_, T_, GRAD, qas_hrb(N_, T_), qas_hpd(N_, T_), 0, 0, BLK_>(ctx, P, st);
    QAS_CASE(2, 8, 256) QAS_CASE(3, 8, 256) QAS_CASE(4, 8, 256) QAS_CASE(5, 16, 256) QAS_CASE(6, 32, 256)
    QAS_CASE(7, 32, 256) QAS_CASE(8, 32, 256) QAS_CASE(9, 64, 256) QAS_CASE(10, 64, 128) QAS_CASE(11, 128, 128)
    QAS_CASE(12, 256, 256)
    QAS_CASE(10, 32, 64)                                                                          // energy-only launches; QAS_B200_TEAM_SHIFT=-1
    QAS_CASE(8, 64, 256) QAS_CASE(9, 128, 256) QAS_CASE(10, 128, 256) QAS_CASE(11, 256, 256)     // QAS_B200_TEAM_SHIFT=1
#undef QAS_CASE
    if constexpr (!GRAD) { if (P.n == 13 && T == 256) return launch_smem<13, 256, false, qas_hrb(13, 256), qas_hpd(13, 256)>(ctx, P, st); }
    return cudaErrorInvalidConfiguration;      // no instantiation for this (n, T, CTA) combination
  }
  if (qas_tiled_applies(ctx)) return qas_launch_tiled(ctx, P, GRAD, st);
  return launch_gmem<GRAD>(ctx, P, st);
}

#if QAS_GRAD
cudaError_t qas_launch_eval_grad(qas_ctx *ctx, EvalParams &P, cudaStream_t st) { return launch_eval<true>(ctx, P, st); }
#else
cudaError_t qas_launch_eval_nograd(qas_ctx *ctx, EvalParams &P, cudaStream_t st) { return launch_eval<false>(ctx, P, st); }
#endif

