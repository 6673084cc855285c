// launch_c.cu -- launches of the support-compressed path (kernels_c.cuh): c-tape compile kernel,
// compressed evaluation kernel per bin, fused gradient finalize + batch reduction.
#include "ctx.h"

size_t qas_ccompile_smem_bytes(const qas_ctx *ctx, int ops_cap, int p) {
  return qas_ccompile_smem_words(ctx->n, ops_cap, p) * 4 * 64 + (size_t)ctx->c * sizeof(qas_pool_entry);
}

template <int NQ>
static cudaError_t launch_compile_c_t(qas_ctx *ctx, const CompileCParams &P, cudaStream_t st) {
  constexpr int CT = 64;
  const size_t smem = qas_ccompile_smem_words(ctx->n, P.ops_cap, P.p) * 4 * CT + (size_t)P.c * sizeof(qas_pool_entry);
  static QasLaunchCache cache;
  cudaError_t e = qas_prepare_launch(cache, ctx->device, compile_ctapes_kernel<CT, NQ>, CT, smem, nullptr);
  if (e != cudaSuccess) return e;
  compile_ctapes_kernel<CT, NQ><<<(P.B + CT - 1) / CT + P.gate_blocks, CT, smem, st>>>(P);
  return cudaGetLastError();
}
cudaError_t qas_launch_compile_c(qas_ctx *ctx, const CompileCParams &P, cudaStream_t st) {
  if (ctx->n <= 10) return launch_compile_c_t<10>(ctx, P, st);
  if (ctx->n <= 12) return launch_compile_c_t<12>(ctx, P, st);
  return launch_compile_c_t<16>(ctx, P, st);
}

template <int T, bool GRAD, int BLK, int MINB, int AG = 2>
static cudaError_t launch_evalc_t(qas_ctx *ctx, const EvalCParams &P, int max_work, cudaStream_t st) {
  constexpr int TEAMS = BLK / T;
  const size_t smem = evalc_team_smem_bytes(P.slot_bits, T, GRAD, P.ops_cap, P.S) * TEAMS;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  auto kern = evalc_kernel<T, GRAD, BLK, MINB, AG>;
  static QasLaunchCache cache;
  int occ = 0;
  cudaError_t e = qas_prepare_launch(cache, ctx->device, kern, BLK, smem, &occ);
  if (e != cudaSuccess) return e;
  const int grid = std::min((max_work + TEAMS - 1) / TEAMS, std::max(1, occ) * ctx->num_sms);   // persistent CTAs, queue-fed
  ctx->stats.teams_per_cta = TEAMS;
  ctx->stats.threads_per_team = T;
  ctx->stats.ctas_per_sm = occ;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 3;
  kern<<<grid, BLK, smem, st>>>(P);
  return cudaGetLastError();
}

// threads per team of a compressed bin by its slot size (log2 amplitudes)
// Sub-warp teams for the small registers (several circuits per warp): the register file -- not the lanes -- bounds the
// circuits in flight, and a G = 2 pass of a 2^6-amplitude register has 16 cosets, so half a warp's lanes were idle.
// QAS_B200_C_T16 (tuning knob): registers of <= 2^(5 + value) amplitudes run as 16-lane teams (0 = off; default 1 for
// energy-only launches, whose forward sweep and H phase are full-width, 2 for gradient launches).
// QAS_B200_C_T8: registers of <= 2^6 amplitudes of gradient launches as 8-lane teams (default 0).
int qas_evalc_team_threads(int slot_bits, bool grad) {
  static const int t16 = []() { const char *e = getenv("QAS_B200_C_T16"); return e ? atoi(e) : -1; }();
  static const int t8 = []() { const char *e = getenv("QAS_B200_C_T8"); return e ? atoi(e) : 0; }();
  const int lim16 = t16 >= 0 ? t16 : (grad ? 2 : 1);
  if (t8 > 0 && grad && slot_bits <= 6) return 8;
  return (lim16 > 0 && slot_bits <= 5 + lim16) ? 16 : slot_bits <= 8 ? 32 : slot_bits <= 10 ? 64 : slot_bits == 11 ? 128 : 256;
}

cudaError_t qas_launch_evalc(qas_ctx *ctx, const EvalCParams &P, int max_work, bool grad, cudaStream_t st) {
  const int T = qas_evalc_team_threads(P.slot_bits, grad);
  // QAS_B200_C_MINB (tuning knob, one-warp teams): CTAs per SM the register allocation is bounded for -- 4 (128
  // registers, default), 5 (102), 6 (85) or 8 (64): more circuits in flight against spills in the two-gate adjoint pass
  static const int minb32 = []() { const char *e = getenv("QAS_B200_C_MINB"); const int v = e ? atoi(e) : 4; return (v == 5 || v == 6 || v == 8) ? v : 4; }();
#define QAS_CCASE(T_, BLK_, MINB_) \
  return grad ? launch_evalc_t<T_, true, BLK_, MINB_>(ctx, P, max_work, st) : launch_evalc_t<T_, false, BLK_, MINB_>(ctx, P, max_work, st);
  // QAS_B200_C_AG1 (tuning knob, bit mask, default 3): one-gate adjoint passes, registers bounded for more CTAs per SM, for the
  // gradient launches of bit 0: the 2^6-amplitude slot (16-lane teams; QAS_B200_C_AG1_MINB = 5 | 6 CTAs per SM),
  // bit 1: the one-warp teams (5 CTAs per SM), bit 2: the two-warp teams (5 CTAs per SM).  The 2^7 slot is bound by shared memory at 4 CTAs either way.
  static const int ag1 = []() { const char *e = getenv("QAS_B200_C_AG1"); return e ? atoi(e) : 3; }();
  static const int ag1_minb = []() { const char *e = getenv("QAS_B200_C_AG1_MINB"); return e && atoi(e) == 6 ? 6 : 5; }();
  if ((ag1 & 1) && grad && T == 16 && P.slot_bits <= 6)
    return ag1_minb == 6 ? launch_evalc_t<16, true, 128, 6, 1>(ctx, P, max_work, st) : launch_evalc_t<16, true, 128, 5, 1>(ctx, P, max_work, st);
  if ((ag1 & 2) && grad && T == 32 && minb32 == 4) return launch_evalc_t<32, true, 128, 5, 1>(ctx, P, max_work, st);
  if ((ag1 & 4) && grad && T == 64) return launch_evalc_t<64, true, 128, 5, 1>(ctx, P, max_work, st);
  if (T == 8) { QAS_CCASE(8, 128, 4) }
  if (T == 16) { QAS_CCASE(16, 128, 4) }
  if (T == 32 && minb32 == 5) { QAS_CCASE(32, 128, 5) }
  if (T == 32 && minb32 == 6) { QAS_CCASE(32, 128, 6) }
  if (T == 32 && minb32 == 8) { QAS_CCASE(32, 128, 8) }
  if (T == 32) { QAS_CCASE(32, 128, 4) }
  if (T == 64) { QAS_CCASE(64, 128, 4) }
  if (T == 128) { QAS_CCASE(128, 128, 3) }
  if (T == 256) { QAS_CCASE(256, 256, 1) }
#undef QAS_CCASE
  return cudaErrorInvalidConfiguration;
}
