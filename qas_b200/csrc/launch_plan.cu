// launch_plan.cu -- launches of the device-side classic tape compiler / pass planner (kernels_plan.cuh)
#include "ctx.h"
#include "kernels_plan.cuh"

// classic path: one thread per circuit compiles and plans its tape (all circuits of the batch)
int qas_launch_compile_classic(qas_ctx *ctx, const int32_t *dk, int B, int p, int ops_cap, int swz_plan, size_t rec_words,
                               cudaStream_t st) {
  const int cthr = 64, c = ctx->c;
  // ops in shared memory only while that keeps >= 8 CTAs per SM resident (short tapes)
  size_t ops_smem_limit = 24 * 1024;
  if (const char *oe = getenv("QAS_B200_COMPILE_SMEM_KB")) ops_smem_limit = (size_t)std::max(0, atoi(oe)) * 1024;   // tuning knob
  const bool ops_smem = qas_compile_smem_words(ctx->n, ops_cap, p, true) * 4 * cthr <= ops_smem_limit;
  const size_t csm = qas_compile_smem_words(ctx->n, ops_cap, p, ops_smem) * 4 * cthr;
  if (csm > 200 * 1024) return fail(ctx, QAS_ERR_UNSUPPORTED, "structure lists too long for the tape-compile kernel");
  if (ops_smem) {
    static QasLaunchCache cache;
    CK(qas_prepare_launch(cache, ctx->device, compile_tapes_kernel<true>, cthr, csm, nullptr));
    compile_tapes_kernel<true><<<(B + cthr - 1) / cthr, cthr, csm, st>>>(dk, ctx->d_pool, B, p, c, ctx->n, ops_cap, ctx->init_kind,
                                                                       ctx->init_basis, ctx->GMAX, swz_plan, ctx->hinfo, ctx->d_tape, rec_words);
  } else {
    static QasLaunchCache cache;
    CK(qas_prepare_launch(cache, ctx->device, compile_tapes_kernel<false>, cthr, csm, nullptr));
    compile_tapes_kernel<false><<<(B + cthr - 1) / cthr, cthr, csm, st>>>(dk, ctx->d_pool, B, p, c, ctx->n, ops_cap, ctx->init_kind,
                                                                        ctx->init_basis, ctx->GMAX, swz_plan, ctx->hinfo, ctx->d_tape, rec_words);
  }
  CK(cudaGetLastError());
  return QAS_OK;
}

cudaError_t qas_launch_plan_full(qas_ctx *ctx, uint32_t *tape, size_t rec_stride, const int *count, const int *order, int B,
                                 int ops_cap, int swizzled, cudaStream_t st) {
  const int thr = 32;
  const size_t smem = (size_t)thr * ((ops_cap * QAS_OP_WORDS) | 1) * 4;
  static QasLaunchCache cache;
  cudaError_t e = qas_prepare_launch(cache, ctx->device, plan_full_kernel, thr, smem, nullptr);
  if (e != cudaSuccess) return e;
  plan_full_kernel<<<(B + thr - 1) / thr, thr, smem, st>>>(tape, rec_stride, count, order, ctx->n, ops_cap, ctx->init_kind, swizzled, ctx->hinfo);
  return cudaGetLastError();
}
