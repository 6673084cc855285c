// exchange.cu -- the exchange step of the sharded candidate batch over NVLink peer memory, without NCCL on the
// data path (C ABI: qas_peer_push_row / qas_peer_wait_reduce, include/qas_b200.h).
//
// What is exchanged (qas/mcts/mcts.py:340-352 spread over ranks): every rank holds one packed row
// [B energies | p*c*l gradient sum]; every rank needs all rows (the rewards feed its copy of the tree) and the
// mean of the gradient parts.  Each rank owns a receive buffer of `world` rows and `world` 64-bit flags that is
// mapped into every peer's address space (torch symmetric memory / CUDA IPC: the host passes the pointers).
//   push:  ONE kernel copies the local row straight into slot `rank` of every peer's receive buffer (plain
//          16-byte stores over NVLink / NVSwitch, and into its own), then -- once the last CTA has fenced its
//          stores -- release-stores the step number into its flag on every peer.
//   wait + reduce:  one warp spins (acquire loads, system scope) on the LOCAL flags until every rank's row of
//          this step has landed, then the fixed-order reduction of the gradient parts runs on the same stream
//          (grad_final_reduce_kernel: bit-identical on every rank).
// No channel CTAs sit on the SMs while a peer is late (an NCCL all-gather holds its CTAs spinning next to the
// persistent evaluation kernels), and nothing is staged through intermediate buffers.
// Ordering contract: every rank issues push(s), wait_reduce(s), push(s + 1), ... on ONE stream with two
// alternating receive buffers (s & 1).  push(s + 2) then follows this rank's wait(s + 1), which saw every peer's
// push(s + 1), which followed that peer's wait(s) on its stream -- so a slot is never overwritten before its
// reader is done, without acknowledgements.
#include "ctx.h"

#define QAS_PEER_MAX 16

struct PeerPtrs {
  double *row[QAS_PEER_MAX];                 // where this rank's row lives on rank r
  unsigned long long *flag[QAS_PEER_MAX];    // this rank's flag on rank r
};

__device__ unsigned int g_push_done[2];      // last-CTA election of the push kernel, per buffer parity

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// count doubles, count even; src and every destination 16-byte aligned
__global__ void __launch_bounds__(256) peer_push_kernel(const double2 *src, size_t count2, const PeerPtrs P, int world,
                                                        unsigned long long seq) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count2; i += (size_t)gridDim.x * blockDim.x) {
    const double2 v = src[i];
    for (int r = 0; r < world; ++r) reinterpret_cast<double2 *>(P.row[r])[i] = v;
  }
  __threadfence_system();                    // this thread's peer stores are ordered before what follows
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&g_push_done[seq & 1ull], 1u);
    if (prev == gridDim.x - 1) {             // every CTA has fenced its stores: publish the row
      g_push_done[seq & 1ull] = 0u;
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(P.flag[r], seq);
    }
  }
}

// lane r waits for rank r's row of step `seq`; status[0] = 1 if a row has not arrived after ~timeout_ns
__global__ void peer_wait_kernel(const unsigned long long *flags, int world, unsigned long long seq,
                                 unsigned long long timeout_ns, unsigned int *status) {
  const int r = threadIdx.x;
  if (r >= world) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(flags + r) < seq) {
    __nanosleep(200);
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) { atomicExch(status, 1u); break; }
  }
}

static __global__ void peer_reduce_kernel(const double *rows, int nrows, size_t stride, size_t total, double scale, double *out) {
  // same summation order as grad_final_reduce_kernel (kernels.cuh): lane-strided rows in ascending order, then a
  // fixed shuffle tree -- the result depends only on nrows
  const size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (x >= total) return;
  double s = 0.0;
  for (int ch = lane; ch < nrows; ch += 32) s += rows[(size_t)ch * stride + x];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[x] = s * scale;
}

extern "C" int qas_peer_push_row(const void *row, uint64_t count, void *const *peer_rows, void *const *peer_flags, int world,
                                 uint64_t seq, void *stream) {
  if (!row || !peer_rows || !peer_flags || world < 1 || world > QAS_PEER_MAX || count == 0 || (count & 1) || seq == 0) {
    g_create_error = "bad arguments to qas_peer_push_row (count must be even, 1 <= world <= 16, seq >= 1)";
    return QAS_ERR_INVALID;
  }
  PeerPtrs P{};
  for (int r = 0; r < world; ++r) {
    if (!peer_rows[r] || !peer_flags[r] || ((uintptr_t)peer_rows[r] & 15) || ((uintptr_t)peer_flags[r] & 7)) {
      g_create_error = "qas_peer_push_row: null or misaligned peer pointer";
      return QAS_ERR_INVALID;
    }
    P.row[r] = (double *)peer_rows[r];
    P.flag[r] = (unsigned long long *)peer_flags[r];
  }
  if ((uintptr_t)row & 15) { g_create_error = "qas_peer_push_row: row must be 16-byte aligned"; return QAS_ERR_INVALID; }
  const size_t count2 = count / 2;
  const int grid = (int)std::min<size_t>(64, (count2 + 255) / 256);
  peer_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2 *)row, count2, P, world, (unsigned long long)seq);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string("peer_push_kernel: ") + cudaGetErrorString(e); return QAS_ERR_CUDA; }
  return QAS_OK;
}

extern "C" int qas_peer_wait_reduce(const void *rows, int world, uint64_t row_stride, uint64_t offset, uint64_t total, double scale,
                                    void *out, const void *flags, uint64_t seq, void *status, void *stream) {
  if (!rows || !flags || !status || world < 1 || world > QAS_PEER_MAX || seq == 0 || (total && (!out || offset + total > row_stride))) {
    g_create_error = "bad arguments to qas_peer_wait_reduce";
    return QAS_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  peer_wait_kernel<<<1, 32, 0, st>>>((const unsigned long long *)flags, world, (unsigned long long)seq, 5000000000ull, (unsigned int *)status);
  if (total)
    peer_reduce_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, st>>>((const double *)rows + offset, world, (size_t)row_stride,
                                                                            (size_t)total, scale, (double *)out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string("qas_peer_wait_reduce: ") + cudaGetErrorString(e); return QAS_ERR_CUDA; }
  return QAS_OK;
}
