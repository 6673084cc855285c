// kernels_plan.cuh -- the kernels that run the classic pass planner (tape.h qas_plan_supports) on the device.
// The planner's pivot-table loops are fully unrolled for five register sizes, which makes these the slowest
// kernels to compile: they live in one translation unit (launch_plan.cu).
#pragma once
#include "kernels_c.cuh"

// OPS_SMEM: the ops are built and planned in shared memory (the planner reads them back many
// times; from global memory every read is a dependent L2 round trip) and copied out warp-
// cooperatively, coalesced.  Per thread: frame (2n+1 words) | ops (ops_cap*5 words, odd stride).
// plus the circuit's structure list (p words, odd stride), staged coalesced by the whole CTA.
__host__ __device__ __forceinline__ size_t qas_compile_smem_words(int n, int ops_cap, int p, bool ops_smem) {
  return (size_t)(2 * n + 1) + (size_t)(p | 1) + (ops_smem ? (size_t)((ops_cap * QAS_OP_WORDS) | 1) : 0);
}
template <bool OPS_SMEM>
__global__ void compile_tapes_kernel(const int32_t *k, const qas_pool_entry *pool, int B, int p, int c,
                                     int n, int ops_cap, int init_kind, uint64_t init_basis,
                                     int gmax, int swizzled, QasHInfo hinfo, uint32_t *tape, size_t rec_stride) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = b < B;
  // GF(2) frame (col / row, indexed by wire at run time): shared memory, odd stride per thread
  extern __shared__ uint32_t frame_smem[];
  const size_t stride = qas_compile_smem_words(n, ops_cap, p, OPS_SMEM);
  uint32_t *mine = frame_smem + (size_t)threadIdx.x * stride;
  uint32_t *col = mine, *row = col + n;
  int32_t *ks = reinterpret_cast<int32_t *>(mine + (2 * n + 1));
  uint32_t *rec = tape + (size_t)(live ? b : 0) * rec_stride;
  uint32_t *ops = OPS_SMEM ? mine + (2 * n + 1) + (p | 1) : rec + QAS_REC_HDR;
  {   // the CTA's structure lists are contiguous in global memory: coalesced copy, row t -> thread t
    const size_t first = (size_t)blockIdx.x * blockDim.x * p;
    const size_t count = (size_t)min((int)blockDim.x, B - (int)(blockIdx.x * blockDim.x)) * p;
    for (size_t w = threadIdx.x; w < count; w += blockDim.x) {
      const size_t t = w / p, i = w - t * p;
      reinterpret_cast<int32_t *>(frame_smem + t * stride + (2 * n + 1))[i] = k[first + w];
    }
    __syncthreads();
  }
  int nops = 0;
  if (live) {
    QasTapeSink sink{ops, ops_cap, 0, 0};
    const int st = qas_compile_circuit(ks, p, pool, c, n, col, row, sink);
    nops = st ? 0 : sink.n;
    const uint32_t init_index = (init_kind == QAS_INIT_BASIS && !st) ? qas_frame_map_basis(col, n, init_basis) : 0u;
    rec[0] = (uint32_t)nops;
    rec[1] = (uint32_t)st;
    rec[2] = init_index;
    rec[3] = (uint32_t)qas_plan_groups(ops, nops, gmax, n, init_kind, init_index,
                                       rec + QAS_REC_HDR + (size_t)ops_cap * QAS_OP_WORDS, swizzled,   // passes per sweep
                                       &hinfo, rec + qas_rec_hd_offset(ops_cap, n));
  }
  if constexpr (OPS_SMEM) {
    __syncwarp();
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    for (int t = 0; t < 32; ++t) {
      const int words = __shfl_sync(0xffffffffu, nops, t) * QAS_OP_WORDS;
      const int bt = blockIdx.x * blockDim.x + wbase + t;
      if (bt >= B) break;
      const uint32_t *src = frame_smem + (size_t)(wbase + t) * stride + (2 * n + 1) + (p | 1);
      uint32_t *dst = tape + (size_t)bt * rec_stride + QAS_REC_HDR;
      for (int w = lane; w < words; w += 32) dst[w] = src[w];
    }
  }
}

// Classic planning of the circuits the binning left in the full bin (d == n): support descriptors and
// final-support descriptor of the already grouped tape, written to the record (kernels.cuh layout).  The planner
// re-reads the ops many times: they are staged in shared memory (from global memory every read is an L2 round trip).
// dynamic shared memory: [blockDim.x][(ops_cap * 5) | 1] words
__global__ void plan_full_kernel(uint32_t *tape, size_t rec_stride, const int *bin_count, const int *order,
                                 int n, int ops_cap, int init_kind, int swizzled, QasHInfo hinfo) {
  extern __shared__ uint32_t psm[];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= *bin_count) return;
  uint32_t *rec = tape + (size_t)order[q] * rec_stride;
  const int nops = (int)rec[0];
  uint32_t *ops = psm + (size_t)threadIdx.x * ((ops_cap * QAS_OP_WORDS) | 1);
  for (int w = 0; w < nops * QAS_OP_WORDS; ++w) ops[w] = rec[QAS_REC_HDR + w];
  rec[2] = 0u;                                    // init_index of the all-zero start
  qas_plan_marked(ops, nops, (int)rec[3], n, init_kind, 0u,
                  rec + QAS_REC_HDR + (size_t)ops_cap * QAS_OP_WORDS, swizzled, &hinfo, rec + qas_rec_hd_offset(ops_cap, n));
  for (int w = 0; w < nops * QAS_OP_WORDS; ++w) rec[QAS_REC_HDR + w] = ops[w];
}
