// sv.cu -- large-state mode: one complex128 statevector slab in HBM, gates applied in
// shared-memory-tiled passes (one HBM read + write per pass), Pauli-sum expectation.
// C ABI: include/qas_b200_sv.h.  HBM-bandwidth bound: 32 * 2^n_local bytes per pass.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include "../../include/qas_b200_sv.h"

static thread_local std::string g_sv_error;
extern "C" const char *qas_sv_last_error(void) { return g_sv_error.c_str(); }

#define SVCK(call)                                                                            \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char buf_[512];                                                                         \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),     \
               __FILE__, __LINE__);                                                           \
      g_sv_error = buf_;                                                                      \
      return QAS_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define SV_MAX_TILE_BITS 13
#define SV_THREADS 256

struct SvTileGate {
  int kind;            // QAS_SV_U1 / QAS_SV_PDIAG
  int tbit;            // tile bit index of the target (U1)
  int cbit_tile;       // tile bit index of the control, -1 if none / outside
  int cpos_out;        // slab bit position of a control outside the tile, -1 if none
  uint32_t tmask;      // PDIAG: parity mask over tile bits
  uint32_t pad;
  unsigned long long omask;   // PDIAG: parity mask over slab bits outside the tile
  double2 u[4];
};

struct SvBlockParams {
  double2 *state;
  const SvTileGate *gates;
  int n_local, kb, c, ngates;
  unsigned char bits[16];      // ascending slab positions of the tile bits; bits[0..c-1] = 0..c-1
};

__device__ __forceinline__ double2 sv_cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 sv_cfma(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// One CTA per tile.  Tile element j (kb bits) lives at slab index base | spread(j): the low c
// bits are contiguous (coalesced 16*2^c-byte runs), the others are scattered by offs[].
__global__ void __launch_bounds__(SV_THREADS) sv_block_kernel(const SvBlockParams P) {
  extern __shared__ __align__(16) double2 tile[];
  __shared__ unsigned long long offs[1 << (SV_MAX_TILE_BITS - 5)];
  __shared__ unsigned long long s_base;
  const int kb = P.kb, c = P.c;
  const uint32_t tsize = 1u << kb, nhi = 1u << (kb - c), lowmask = (1u << c) - 1u;
  for (uint32_t h = threadIdx.x; h < nhi; h += SV_THREADS) {
    unsigned long long o = 0;
    for (int q = c; q < kb; ++q) o |= (unsigned long long)((h >> (q - c)) & 1u) << P.bits[q];
    offs[h] = o;
  }
  const unsigned long long ntiles = 1ull << (P.n_local - kb);
  for (unsigned long long tid_ = blockIdx.x; tid_ < ntiles; tid_ += gridDim.x) {
    if (threadIdx.x == 0) {
      unsigned long long base = 0, rest = tid_;
      int q = 0;
      for (int pos = 0; pos < P.n_local; ++pos) {
        if (q < kb && P.bits[q] == pos) { ++q; continue; }
        base |= (rest & 1ull) << pos;
        rest >>= 1;
      }
      s_base = base;
    }
    __syncthreads();
    const unsigned long long base = s_base;
    double2 *g = P.state + base;
    for (uint32_t j = threadIdx.x; j < tsize; j += SV_THREADS) tile[j] = g[(j & lowmask) | offs[j >> c]];
    __syncthreads();
    for (int gi = 0; gi < P.ngates; ++gi) {
      const SvTileGate *gt = P.gates + gi;
      if (gt->kind == QAS_SV_U1) {
        const bool on = gt->cpos_out < 0 || ((base >> gt->cpos_out) & 1ull);
        if (on) {
          const double2 u00 = gt->u[0], u01 = gt->u[1], u10 = gt->u[2], u11 = gt->u[3];
          const int tb = gt->tbit, cb = gt->cbit_tile;
          const uint32_t tbm = 1u << tb;
#pragma unroll 4
          for (uint32_t t = threadIdx.x; t < (tsize >> 1); t += SV_THREADS) {
            const uint32_t j0 = ((t >> tb) << (tb + 1)) | (t & (tbm - 1u));
            if (cb >= 0 && !((j0 >> cb) & 1u)) continue;
            const uint32_t j1 = j0 | tbm;
            const double2 a0 = tile[j0], a1 = tile[j1];
            tile[j0] = sv_cfma(u01, a1, sv_cmul(u00, a0));
            tile[j1] = sv_cfma(u11, a1, sv_cmul(u10, a0));
          }
        }
      } else {
        const bool flip = __popcll(base & gt->omask) & 1;
        const double2 d0 = flip ? gt->u[3] : gt->u[0], d1 = flip ? gt->u[0] : gt->u[3];
        const uint32_t tm = gt->tmask;
#pragma unroll 4
        for (uint32_t j = threadIdx.x; j < tsize; j += SV_THREADS)
          tile[j] = sv_cmul((__popc(j & tm) & 1) ? d1 : d0, tile[j]);
      }
      __syncthreads();
    }
    for (uint32_t j = threadIdx.x; j < tsize; j += SV_THREADS) g[(j & lowmask) | offs[j >> c]] = tile[j];
    __syncthreads();
  }
}

__global__ void sv_init_kernel(double2 *state, unsigned long long dim, long long index, double2 amp) {
  const unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = i0; i < dim; i += stride)
    state[i] = (index < 0 || (unsigned long long)index == i) ? amp : make_double2(0.0, 0.0);
}

// ---- expectation: groups by X mask, signs evaluated on the fly (no 2^n tables at this size)
struct SvExpParams {
  const double2 *state;
  const unsigned long long *gx;      // [G] x mask (local bits only)
  const int *gbegin;                 // [G+1]
  const int *gimag;                  // [G]
  const unsigned long long *tz;      // [T] z mask over the full index
  const double *tc;                  // [T]
  unsigned long long rank_hi;        // rank_bits << n_local
  unsigned long long dim;
  int G;
  double *partial;                   // [gridDim.x]
};

__global__ void __launch_bounds__(SV_THREADS) sv_expval_kernel(const SvExpParams P) {
  double acc = 0.0;
  const unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = i0; i < P.dim; i += stride) {
    const double2 a = P.state[i];
    for (int g = 0; g < P.G; ++g) {
      const unsigned long long x = P.gx[g];
      const unsigned long long jf = (i ^ x) | P.rank_hi;
      double w = 0.0;
      for (int t = P.gbegin[g]; t < P.gbegin[g + 1]; ++t) w += (__popcll(jf & P.tz[t]) & 1) ? -P.tc[t] : P.tc[t];
      const double2 b = x ? P.state[i ^ x] : a;
      // Re(conj(a) * w * b)   or, for imaginary slices, Re(conj(a) * i w * b) = -w Im(conj(a) b)
      if (P.gimag[g]) acc -= w * (a.x * b.y - a.y * b.x);
      else acc += w * (a.x * b.x + a.y * b.y);
    }
  }
  __shared__ double red[SV_THREADS / 32];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < SV_THREADS / 32; ++w) s += red[w];
    P.partial[blockIdx.x] = s;
  }
}

__global__ void sv_final_sum_kernel(const double *partial, int n, double *out) {
  __shared__ double red[SV_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += SV_THREADS) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int off = SV_THREADS / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// ------------------------------------------------------------------------------ host side
extern "C" int qas_sv_init(void *state, int n_local, int64_t index, double amp_re, double amp_im, void *stream) {
  if (!state || n_local < 1 || n_local > 34) { g_sv_error = "bad arguments"; return QAS_ERR_INVALID; }
  const unsigned long long dim = 1ull << n_local;
  const int grid = (int)std::min<unsigned long long>((dim + 255) / 256, 148 * 16);
  sv_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((double2 *)state, dim, (long long)index,
                                                         make_double2(amp_re, amp_im));
  SVCK(cudaGetLastError());
  return QAS_OK;
}

namespace {
struct Block {
  std::vector<int> bits;
  std::vector<int> gates;
};

// Greedy dependency-aware blocking: a gate joins the current block if its target fits into the
// kb tile bits and no earlier deferred gate shares a bit with it (so deferral never reorders
// non-commuting gates).  PDIAG gates need no pairing and fit any tile.
std::vector<Block> plan_blocks(const qas_sv_gate *gates, int ngates, int n_local, int kb, int c) {
  std::vector<Block> blocks;
  std::vector<int> remaining(ngates);
  for (int i = 0; i < ngates; ++i) remaining[i] = i;
  while (!remaining.empty()) {
    std::vector<char> active(n_local, 0), blocked(n_local, 0);
    int nact = 0;
    for (int b = 0; b < c; ++b) { active[b] = 1; ++nact; }
    Block blk;
    std::vector<int> next;
    for (int gi : remaining) {
      const qas_sv_gate &g = gates[gi];
      std::vector<int> touches;
      if (g.kind == QAS_SV_U1) {
        touches.push_back(g.target);
        if (g.control >= 0) touches.push_back(g.control);
      } else {
        for (int b = 0; b < n_local; ++b) if ((g.mask >> b) & 1ull) touches.push_back(b);
      }
      bool dep = false;
      for (int b : touches) dep = dep || blocked[b];
      const bool need_new = g.kind == QAS_SV_U1 && !active[g.target];
      if (!dep && (!need_new || nact < kb)) {
        if (need_new) { active[g.target] = 1; ++nact; }
        blk.gates.push_back(gi);
      } else {
        for (int b : touches) blocked[b] = 1;
        next.push_back(gi);
      }
    }
    // a control inside the tile is cheaper than a per-tile predicate only if it is already active;
    // pad the tile with the lowest inactive bits
    for (int b = 0; b < n_local && nact < kb; ++b) if (!active[b]) { active[b] = 1; ++nact; }
    for (int b = 0; b < n_local; ++b) if (active[b]) blk.bits.push_back(b);
    if (blk.gates.empty()) return {};          // cannot happen with c < kb; guards against a livelock
    blocks.push_back(std::move(blk));
    remaining.swap(next);
  }
  return blocks;
}
}  // namespace

// Host-only view of the pass planner (no GPU needed): block index of every gate and the tile bits
// of every block.  Used by the CPU test tier.
extern "C" int qas_sv_plan(int n_local, const qas_sv_gate *gates, int ngates, int tile_bits,
                           int *block_of_gate, int *num_blocks, uint64_t *block_bits, int max_blocks) {
  if (n_local < 1 || !gates || ngates <= 0 || !block_of_gate || !num_blocks) return QAS_ERR_INVALID;
  int kb = tile_bits > 0 ? tile_bits : 12;
  kb = std::min(std::min(kb, SV_MAX_TILE_BITS), n_local);
  const int c = std::max(0, std::min(5, kb - 1));
  std::vector<Block> blocks = plan_blocks(gates, ngates, n_local, kb, c);
  if (blocks.empty()) { g_sv_error = "gate blocking made no progress"; return QAS_ERR_UNSUPPORTED; }
  *num_blocks = (int)blocks.size();
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    for (int gi : blocks[bi].gates) block_of_gate[gi] = (int)bi;
    if (block_bits && (int)bi < max_blocks) {
      uint64_t m = 0;
      for (int b : blocks[bi].bits) m |= 1ull << b;
      block_bits[bi] = m;
    }
  }
  return QAS_OK;
}

extern "C" int qas_sv_apply(void *state, int n_local, const qas_sv_gate *gates, int ngates, int tile_bits,
                            void *stream, int sync, qas_sv_info *info) {
  if (!state || n_local < 1 || n_local > 34 || (ngates > 0 && !gates)) { g_sv_error = "bad arguments"; return QAS_ERR_INVALID; }
  for (int i = 0; i < ngates; ++i) {
    const qas_sv_gate &g = gates[i];
    if (g.kind == QAS_SV_U1) {
      if (g.target < 0 || g.target >= n_local || g.control >= n_local || g.control == g.target) {
        g_sv_error = "gate " + std::to_string(i) + ": bit position out of range"; return QAS_ERR_INVALID;
      }
    } else if (g.kind == QAS_SV_PDIAG) {
      if (n_local < 64 && (g.mask >> n_local)) { g_sv_error = "gate " + std::to_string(i) + ": mask out of range"; return QAS_ERR_INVALID; }
    } else { g_sv_error = "unknown gate kind"; return QAS_ERR_INVALID; }
  }
  cudaStream_t st = (cudaStream_t)stream;
  int kb = tile_bits > 0 ? tile_bits : 12;
  kb = std::min(std::min(kb, SV_MAX_TILE_BITS), n_local);
  // c = contiguous low bits every tile contains (coalescing); leave at least one free tile bit
  const int c = std::max(0, std::min(5, kb - 1));
  if (info) { info->passes = 0; info->tile_bits = kb; info->bytes_per_pass = 32.0 * (double)(1ull << n_local); info->last_ms = 0.0; }
  if (ngates == 0) return QAS_OK;
  std::vector<Block> blocks = plan_blocks(gates, ngates, n_local, kb, c);
  if (blocks.empty()) { g_sv_error = "gate blocking made no progress (tile_bits too small)"; return QAS_ERR_UNSUPPORTED; }

  // translate every block's gates into tile coordinates, upload once
  std::vector<SvTileGate> tg;
  std::vector<int> first(blocks.size() + 1, 0);
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    const Block &blk = blocks[bi];
    std::vector<int> tix(n_local, -1);
    for (size_t q = 0; q < blk.bits.size(); ++q) tix[blk.bits[q]] = (int)q;
    for (int gi : blk.gates) {
      const qas_sv_gate &g = gates[gi];
      SvTileGate t{};
      t.kind = g.kind;
      t.tbit = -1; t.cbit_tile = -1; t.cpos_out = -1;
      if (g.kind == QAS_SV_U1) {
        t.tbit = tix[g.target];
        if (g.control >= 0) { if (tix[g.control] >= 0) t.cbit_tile = tix[g.control]; else t.cpos_out = g.control; }
      } else {
        for (int b = 0; b < n_local; ++b)
          if ((g.mask >> b) & 1ull) { if (tix[b] >= 0) t.tmask |= 1u << tix[b]; else t.omask |= 1ull << b; }
      }
      for (int q = 0; q < 4; ++q) t.u[q] = make_double2(g.m[2 * q], g.m[2 * q + 1]);
      tg.push_back(t);
    }
    first[bi + 1] = (int)tg.size();
  }
  SvTileGate *d_gates = nullptr;
  SVCK(cudaMallocAsync((void **)&d_gates, tg.size() * sizeof(SvTileGate), st));
  SVCK(cudaMemcpyAsync(d_gates, tg.data(), tg.size() * sizeof(SvTileGate), cudaMemcpyHostToDevice, st));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (info && sync) { SVCK(cudaEventCreate(&e0)); SVCK(cudaEventCreate(&e1)); SVCK(cudaEventRecord(e0, st)); }
  const size_t smem = (size_t)16 << kb;
  SVCK(cudaFuncSetAttribute(sv_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned long long ntiles = 1ull << (n_local - kb);
  const int grid = (int)std::min<unsigned long long>(ntiles, 1u << 20);
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    SvBlockParams P{};
    P.state = (double2 *)state;
    P.gates = d_gates + first[bi];
    P.n_local = n_local; P.kb = kb; P.c = c; P.ngates = first[bi + 1] - first[bi];
    for (int q = 0; q < kb; ++q) P.bits[q] = (unsigned char)blocks[bi].bits[q];
    sv_block_kernel<<<grid, SV_THREADS, smem, st>>>(P);
    SVCK(cudaGetLastError());
  }
  // the upload buffer must outlive the kernels: tg is host memory copied asynchronously from a
  // pageable vector, so wait for the copy before it goes out of scope
  if (info && sync) SVCK(cudaEventRecord(e1, st));
  SVCK(cudaFreeAsync(d_gates, st));
  SVCK(cudaStreamSynchronize(st));
  if (info) {
    info->passes = (int)blocks.size();
    if (sync) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      info->last_ms = ms;
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
  }
  return QAS_OK;
}

extern "C" int qas_sv_expval(const void *state, int n_local, int n_total, uint64_t rank_bits,
                             const qas_pauli_term *terms, int num_terms, double *out_host, void *stream) {
  if (!state || !terms || num_terms <= 0 || !out_host || n_local < 1 || n_total < n_local || n_total > 62) {
    g_sv_error = "bad arguments"; return QAS_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long local_mask = (1ull << n_local) - 1ull;
  std::map<std::pair<unsigned long long, int>, std::vector<std::pair<unsigned long long, double>>> groups;
  for (int t = 0; t < num_terms; ++t) {
    if (terms[t].xmask & ~local_mask) { g_sv_error = "Pauli X/Y on a rank (global) qubit: swap it local first"; return QAS_ERR_UNSUPPORTED; }
    const int ny = __builtin_popcountll(terms[t].xmask & terms[t].zmask);
    groups[{terms[t].xmask, ny & 1}].push_back({terms[t].zmask, ((ny & 2) ? -1.0 : 1.0) * terms[t].coeff});
  }
  std::vector<unsigned long long> gx, tz;
  std::vector<int> gbegin, gimag;
  std::vector<double> tc;
  for (auto &kv : groups) {
    gx.push_back(kv.first.first); gimag.push_back(kv.first.second); gbegin.push_back((int)tz.size());
    for (auto &zc : kv.second) { tz.push_back(zc.first); tc.push_back(zc.second); }
  }
  gbegin.push_back((int)tz.size());
  const int G = (int)gx.size(), grid = 148 * 8;
  unsigned long long *d_gx = nullptr, *d_tz = nullptr; int *d_gbegin = nullptr, *d_gimag = nullptr; double *d_tc = nullptr, *d_partial = nullptr, *d_out = nullptr;
  SVCK(cudaMallocAsync((void **)&d_gx, 8 * G, st)); SVCK(cudaMallocAsync((void **)&d_tz, 8 * tz.size(), st));
  SVCK(cudaMallocAsync((void **)&d_gbegin, 4 * (G + 1), st)); SVCK(cudaMallocAsync((void **)&d_gimag, 4 * G, st));
  SVCK(cudaMallocAsync((void **)&d_tc, 8 * tc.size(), st)); SVCK(cudaMallocAsync((void **)&d_partial, 8 * grid, st));
  SVCK(cudaMallocAsync((void **)&d_out, 8, st));
  SVCK(cudaMemcpyAsync(d_gx, gx.data(), 8 * G, cudaMemcpyHostToDevice, st));
  SVCK(cudaMemcpyAsync(d_tz, tz.data(), 8 * tz.size(), cudaMemcpyHostToDevice, st));
  SVCK(cudaMemcpyAsync(d_gbegin, gbegin.data(), 4 * (G + 1), cudaMemcpyHostToDevice, st));
  SVCK(cudaMemcpyAsync(d_gimag, gimag.data(), 4 * G, cudaMemcpyHostToDevice, st));
  SVCK(cudaMemcpyAsync(d_tc, tc.data(), 8 * tc.size(), cudaMemcpyHostToDevice, st));
  SvExpParams P{};
  P.state = (const double2 *)state; P.gx = d_gx; P.gbegin = d_gbegin; P.gimag = d_gimag; P.tz = d_tz; P.tc = d_tc;
  P.rank_hi = (unsigned long long)rank_bits << n_local; P.dim = 1ull << n_local; P.G = G; P.partial = d_partial;
  sv_expval_kernel<<<grid, SV_THREADS, 0, st>>>(P);
  SVCK(cudaGetLastError());
  sv_final_sum_kernel<<<1, SV_THREADS, 0, st>>>(d_partial, grid, d_out);
  SVCK(cudaGetLastError());
  SVCK(cudaMemcpyAsync(out_host, d_out, 8, cudaMemcpyDeviceToHost, st));
  SVCK(cudaStreamSynchronize(st));
  cudaFreeAsync(d_gx, st); cudaFreeAsync(d_tz, st); cudaFreeAsync(d_gbegin, st); cudaFreeAsync(d_gimag, st);
  cudaFreeAsync(d_tc, st); cudaFreeAsync(d_partial, st); cudaFreeAsync(d_out, st);
  return QAS_OK;
}

extern "C" int qas_sv_norm2(const void *state, int n_local, double *out_host, void *stream) {
  qas_pauli_term id{0, 0, 1.0};
  return qas_sv_expval(state, n_local, n_local, 0, &id, 1, out_host, stream);
}
