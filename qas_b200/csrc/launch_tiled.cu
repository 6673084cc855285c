// launch_tiled.cu -- launch of the tiled evaluation kernel (kernels_t.cuh): full-register circuits on 13..16 qubits,
// one thread-block cluster per circuit, state slab in L2 / HBM, tape blocked into tile passes.
#include "ctx.h"
#include "kernels_t.cuh"

// tuning knobs (qas_ctx::tiled_*, read at context creation): log2 amplitudes per tile (11: two CTAs of 256 threads
// per SM, 12: one CTA of 512), CTAs per cluster (1, 2, 4; 0 = by register size)

template <bool GRAD, int TB, int NT, int CS, int MINB>
static cudaError_t launch_tiled_t(qas_ctx *ctx, EvalParams &P, cudaStream_t st) {
  constexpr int CLOW = 2;
  const size_t smem = evalt_smem_bytes(TB, NT, GRAD, P.ops_cap, P.C_real + P.C_imag);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  auto kern = eval_tiled_kernel<GRAD, TB, NT, CS, CLOW, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nclusters = 0;
  cfg.gridDim = dim3(CS * ctx->num_sms);          // the occupancy query wants a grid that is a multiple of the cluster
  e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
  if (e != cudaSuccess) return e;
  if (nclusters < 1) return cudaErrorInvalidConfiguration;
  const size_t per_cluster = ((size_t)1 << P.n) * (GRAD ? 2 : 1);
  int grid_clusters = std::min(P.B, nclusters);
  e = ensure(&ctx->d_gstate, &ctx->cap_gstate, per_cluster * grid_clusters);
  if (e != cudaSuccess) return e;
  P.gstate = ctx->d_gstate;
  cfg.gridDim = dim3(grid_clusters * CS);
  ctx->stats.teams_per_cta = 1;
  ctx->stats.threads_per_team = NT * CS;
  ctx->stats.ctas_per_sm = (nclusters * CS + ctx->num_sms - 1) / ctx->num_sms;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 4;
  return cudaLaunchKernelEx(&cfg, kern, P);
}

template <bool GRAD>
static cudaError_t launch_tiled_g(qas_ctx *ctx, EvalParams &P, cudaStream_t st) {
  const int tb = (ctx->tiled_tb == 12 && P.n >= 12) ? 12 : 11;
  // CTAs per cluster: one CTA per circuit is the most efficient split when the batch fills the machine (no cluster
  // barriers, no idle ranks); the small batches of a real search (25..100 circuits per epoch) get as many CTAs per
  // circuit as it takes to occupy every SM -- the cluster is what strong-scales a single evaluation.
  int cs = ctx->tiled_cs;
  if (cs <= 0) {
    const int resident = (tb == 11 ? 2 : 1) * ctx->num_sms;
    // the full bin of a device-compiled batch (P.count on the device, unknown here) is a small tail of the batch
    // (LiH 0.8 %, H2O-14q a few dozen circuits): largest clusters, so that it does not outlast the compressed bins
    const int work = P.count ? 1 : P.B;
    cs = 1;
    while (cs < 8 && work * cs * 2 <= resident) cs <<= 1;
    if (tb == 12 && cs > 4) cs = 4;
  }
  while (cs > 1 && (1 << (P.n - tb)) < cs) cs >>= 1;
#define QAS_TCASE(TB_, NT_, CS_, MINB_) if (tb == TB_ && cs == CS_) return launch_tiled_t<GRAD, TB_, NT_, CS_, MINB_>(ctx, P, st);
  QAS_TCASE(11, 256, 1, 2) QAS_TCASE(11, 256, 2, 2) QAS_TCASE(11, 256, 4, 2) QAS_TCASE(11, 256, 8, 2)
  QAS_TCASE(12, 512, 1, 1) QAS_TCASE(12, 512, 2, 1) QAS_TCASE(12, 512, 4, 1)
#undef QAS_TCASE
  return cudaErrorInvalidConfiguration;
}

bool qas_tiled_applies(const qas_ctx *ctx) {
  return ctx->tiled && ctx->n >= 11 && ctx->n <= QAS_TP_MAXQ && ctx->RB == 2 && ctx->PD == 2 && ctx->GMAX == 2;
}

cudaError_t qas_launch_tiled(qas_ctx *ctx, EvalParams &P, bool grad, cudaStream_t st) {
  return grad ? launch_tiled_g<true>(ctx, P, st) : launch_tiled_g<false>(ctx, P, st);
}
