// ctx.h -- the evaluator context shared by the translation units of libqas_b200.so (host side).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels_c.cuh"

extern thread_local std::string g_create_error;

struct qas_ctx {
  int device = 0, n = 0, init_kind = 0, c = 0, l = 0, num_sms = 0;
  uint64_t init_basis = 0;
  std::vector<qas_pool_entry> pool;
  int max_expansion = 1;
  int S = 0;                       // number of sign-table slices
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ek0 = nullptr, ek1 = nullptr;
  bool ek_pending = false;
  // device buffers owned by the context
  qas_pool_entry *d_pool = nullptr;
  uint8_t *d_npar_key = nullptr;   // [c] parameters per pool key (sparse batch reduction)
  bool pool_ctrl_param = false;    // the pool has controlled parametrised gates (the c-tape translation may drop such ops)
  double *d_vtab = nullptr;
  uint32_t *d_sdesc = nullptr;
  int team_shift = 0;              // QAS_B200_TEAM_SHIFT: threads per team = default << shift (tuning)
  int T = 0, RB = 0, GMAX = 1, rb[3] = {0, 0, 0};
  WComb wc{};                      // register subspace W of the H-apply blocking (kernels.cuh)
  int num_classes = 0, C_real = 0, C_imag = 0;
  uint32_t *d_cdesc_plain = nullptr;
  uint32_t *d_cdesc_u = nullptr;   // class representatives in quotient coordinates (support-restricted H application)
  QasHInfo hinfo{};                // register subspace W for the planner's final-support descriptor
  int hs_min = 1;                  // QAS_B200_HSPARSE: minimum log2 saving for the support-restricted path (0 = off)
  double *d_mout = nullptr; size_t cap_mout = 0;
  unsigned long long *d_phase = nullptr;   // QAS_B200_PHASE_CLOCKS=1: per-phase clock sums (diagnostic)
  bool stage_times = false;                // QAS_B200_STAGE_TIMES=1: print per-kernel stream times of every call
  cudaEvent_t sev[8] = {nullptr};
  int PD = 1;                      // H-apply prefetch depth (rows)
  bool small_kernel = true;        // n <= 5: thread-per-circuit kernel (QAS_B200_SMALL=0 disables)
  bool small_now = false;          // ... and chosen for the call in progress
  // tiled kernel (kernels_t.cuh) for registers that do not fit shared memory: QAS_B200_TILED (0 = keep the round-1
  // global-memory kernel), QAS_B200_TILED_FORCE_N (registers >= this skip the shared-memory kernels: tests),
  // QAS_B200_TILED_TB (log2 amplitudes per tile, 11 | 12), QAS_B200_TILED_CS (CTAs per cluster, 0 = by register size)
  bool tiled = true;
  int tiled_force_n = 99, tiled_tb = 11, tiled_cs = 0;
  int *d_counter = nullptr;        // work-queue head of the classic team kernels (= d_bins + QAS_MAX_BINS + full bin)
  // support-compressed path (kernels_c.cuh): circuits binned by the dimension d of their support span
  bool compress = false;           // all-zero start, n >= 7, QAS_B200_COMPRESS != 0
  bool sparse_reduce = true;       // batch reduction without a zero-filled [B][p][l] (QAS_B200_SPARSE_REDUCE=0: memset + dense read)
  bool gate_in_compile = true;     // gate table built by spare CTAs of the c-tape compile kernel (QAS_B200_GATE_IN_COMPILE=0: own launch)
  int c_dmin = 6, c_maxd = QAS_CT_MAX_D;   // smallest / largest compressed slot (log2 amplitudes)
  int c_nbins = 0;                 // compressed bins: slot c_dmin + i, i < c_nbins; the full bin is index c_nbins
  int *d_bins = nullptr;           // [QAS_MAX_BINS] counts | [QAS_MAX_BINS] work-queue heads
  int *d_order = nullptr; size_t cap_order = 0;   // [QAS_MAX_BINS][B] circuit indices per bin
  int *h_bins = nullptr; size_t cap_hbins = 0;    // pinned staging of counts + order for host-compiled batches
  double *d_vflat = nullptr;       // [S_flat][2^n] sign tables by physical index, one row per (X mask, real|imag) slice
  uint32_t *d_sxflat = nullptr;    // [S_flat] X masks of those rows
  uint4 *d_sinfo = nullptr;        // [S_flat] Z-span rank and basis of every slice (kernels_c.cuh c_hphase)
  double *d_slut = nullptr;        // [S_flat][16] sign of the slice as a function of the span parities
  std::vector<uint32_t> sxflat;
  int S_flat = 0, S_real_flat = 0;
  cudaStream_t side[QAS_MAX_BINS] = {nullptr};
  cudaEvent_t fork_ev = nullptr, join_ev[QAS_MAX_BINS] = {nullptr};
  int blk = 256;                   // threads per CTA of the team kernel
  GateTab *d_U = nullptr;
  GateDTab *d_dU = nullptr;
  int tab_p = 0;
  int resident_p = 0;              // depth of the parameter super-set left on the device by qas_set_params (0: none)
  int32_t *d_k = nullptr;  size_t cap_k = 0;
  uint8_t *d_k8 = nullptr;  size_t cap_k8 = 0;     // QAS_F_K_U8 uploads (widened into d_k on the device)
  cudaStream_t copy_stream = nullptr;            // structure-list upload of host-pointer calls (overlaps memsets + gate table)
  cudaEvent_t copy_gate = nullptr, copy_done = nullptr;
  double *d_params = nullptr; size_t cap_params = 0;
  double *d_energy = nullptr; size_t cap_energy = 0;
  double *d_grad = nullptr; size_t cap_grad = 0;
  double *d_partial = nullptr; size_t cap_partial = 0;
  double *d_dense = nullptr; size_t cap_dense = 0;
  double2 *d_gstate = nullptr; size_t cap_gstate = 0;
  uint32_t *d_tape = nullptr; size_t cap_tape = 0;
  uint32_t *h_tape = nullptr; size_t cap_htape = 0;   // pinned staging for host-compiled tapes (tiny batches)
  uint8_t *h_k8 = nullptr; size_t cap_hk8 = 0;        // pinned staging: pageable int32 structure lists narrowed to one byte per key
  qas_stats stats{};
  mutable std::string err;
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),     \
               __FILE__, __LINE__);                                                           \
      if (ctx) ctx->err = buf_; else g_create_error = buf_;                                   \
      return QAS_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

static inline int fail(const qas_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}

template <typename Tp>
static inline cudaError_t ensure(Tp **ptr, size_t *cap, size_t need_elems) {
  if (*cap >= need_elems && *ptr) return cudaSuccess;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *cap = 0;
  size_t want = std::max<size_t>(need_elems, 16);
  cudaError_t e = cudaMalloc((void **)ptr, want * sizeof(Tp));
  if (e == cudaSuccess) *cap = want;
  return e;
}


// Launch attributes of ONE kernel instantiation (a function-local static per call site): the dynamic shared-memory cap is
// only ever raised, and the occupancy of the last (block, shared memory) pair is remembered, so that a call's ~10 launches
// do not cost ~15 extra driver calls (cudaFuncSetAttribute + cudaOccupancyMaxActiveBlocksPerMultiprocessor, a few
// microseconds each: most of the host time of a search-sized batch).  Process-wide, hence the mutex.
#define QAS_MAX_DEVICES 16
struct QasLaunchCache {
  std::mutex mu;
  size_t max_smem[QAS_MAX_DEVICES] = {0};
  size_t occ_smem[QAS_MAX_DEVICES] = {0};
  int occ_block[QAS_MAX_DEVICES] = {0}, occ[QAS_MAX_DEVICES] = {0};
};
template <class K>
static inline cudaError_t qas_prepare_launch(QasLaunchCache &c, int device, K kern, int block, size_t smem, int *occ) {
  const int d = (device >= 0 && device < QAS_MAX_DEVICES) ? device : 0;
  const bool cacheable = device >= 0 && device < QAS_MAX_DEVICES;
  std::lock_guard<std::mutex> lock(c.mu);
  if (!cacheable || smem > c.max_smem[d] || c.max_smem[d] == 0) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, cacheable ? c.max_smem[d] : smem));
    if (e != cudaSuccess) return e;
    if (cacheable) c.max_smem[d] = std::max(std::max(smem, c.max_smem[d]), (size_t)1);
  }
  if (occ) {
    if (!cacheable || c.occ_block[d] != block || c.occ_smem[d] != smem || c.occ[d] == 0) {
      int o = 0;
      const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, block, smem);
      if (e != cudaSuccess) return e;
      if (cacheable) { c.occ_block[d] = block; c.occ_smem[d] = smem; c.occ[d] = o; }
      *occ = o;
    } else {
      *occ = c.occ[d];
    }
  }
  return cudaSuccess;
}

// launch_classic_grad.cu / launch_classic_nograd.cu: the classic evaluation kernels (kernels.cuh)
cudaError_t qas_launch_eval_grad(qas_ctx *ctx, EvalParams &P, cudaStream_t st);
cudaError_t qas_launch_eval_nograd(qas_ctx *ctx, EvalParams &P, cudaStream_t st);
// launch_tiled.cu: tiled kernel for full-register circuits on 11..16 qubits (kernels_t.cuh)
bool qas_tiled_applies(const qas_ctx *ctx);
cudaError_t qas_launch_tiled(qas_ctx *ctx, EvalParams &P, bool grad, cudaStream_t st);
// launch_plan.cu: device-side classic tape compiler / pass planner (kernels_plan.cuh)
int qas_launch_compile_classic(qas_ctx *ctx, const int32_t *dk, int B, int p, int ops_cap, int swz_plan, size_t rec_words,
                               cudaStream_t st);
// launch_c.cu: compressed path (kernels_c.cuh)
size_t qas_ccompile_smem_bytes(const qas_ctx *ctx, int ops_cap, int p);
int qas_evalc_team_threads(int slot_bits, bool grad);
cudaError_t qas_launch_compile_c(qas_ctx *ctx, const CompileCParams &P, cudaStream_t st);
cudaError_t qas_launch_plan_full(qas_ctx *ctx, uint32_t *tape, size_t rec_stride, const int *count, const int *order, int B,
                                 int ops_cap, int swizzled, cudaStream_t st);
cudaError_t qas_launch_evalc(qas_ctx *ctx, const EvalCParams &P, int max_work, bool grad, cudaStream_t st);
