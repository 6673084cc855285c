// sv.cu -- large-state mode: one complex128 statevector slab in HBM, gates applied in
// shared-memory-tiled passes (one HBM read + write per pass), Pauli-sum expectation.
// C ABI: include/qas_b200_sv.h.  HBM-bandwidth bound: 32 * 2^n_local bytes per pass.
#include <algorithm>
#include <array>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <mutex>

#include <cuda.h>
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

#define SV_K_U1 0       // dense 2x2 on the target (optionally controlled)
#define SV_K_PDIAG 1    // parity-diagonal phase
#define SV_K_SWAPX 2    // (controlled) Pauli-X: a pure exchange of the pair, no arithmetic

struct __align__(16) SvTileGate {
  int kind;            // SV_K_*
  int tslot;           // U1: register slot bit (0..2) of the target inside its group
  int cslot;           // U1: register slot bit of the control, -1 if the control is not a slot bit
  uint32_t crmask;     // U1: control = parity(tile index & crmask) for a tile-bit control that is not a slot bit, else 0
  int cpos_out;        // U1: slab bit position of a control outside the tile, -1 if none
  uint32_t tmask;      // PDIAG: parity mask over tile bits
  unsigned long long omask;   // PDIAG: parity mask over slab bits outside the tile
  double2 u[4];
};

// A group = consecutive gates of a pass whose targets fall on at most three logical tile bits:
// every thread holds the 8 amplitudes P ^ span(m[0..2]) and applies ALL gates of the group in
// registers, so the tile goes through shared memory once per group instead of once per gate (the
// un-grouped version was shared-memory-bandwidth bound at ~18 gates per pass).
//
// CNOT folding inside a pass: a CNOT whose control and target are both tile bits is the index map
// j -> C j, linear over GF(2).  It is never executed: the host keeps the frame A of the pass (tile
// position p holds the amplitude of logical tile index A p), a gate on logical bit q then pairs the
// positions (p, p ^ m), m = A^-1 e_q, and reads logical bits as parity(p & r), r = row q of A; the
// coset representatives of a group lie in the common kernel of its r's (host-built table), so slot
// index bits are logical bits.  The frame is undone once, by the pass's store to HBM
// (position of logical index j = A^-1 j).
struct SvGroup {
  int first, ngates;
  uint32_t m[3];       // xor masks (tile positions) of the three logical bits held in registers
  uint32_t w[3];       // swizzle(m[q])
  int simple;          // != 0: the group is at most one uncontrolled 2x2 per register slot (products formed on the
                       // host; bit q set = slot q has one, stored in slot order from `first`): straight-line fast path
};

// The gate list and the groups of the pass being executed live in constant memory: every thread of
// the grid walks the same list, so descriptor fields and matrix entries come through the constant
// cache and the uniform datapath instead of per-thread shared-memory loads (the per-thread
// interpretation of the list was the kernel's largest cost).
#define SV_CONST_GATES 448
#define SV_CONST_GROUPS 448
__constant__ SvTileGate c_sv_gates[SV_CONST_GATES];
__constant__ SvGroup c_sv_groups[SV_CONST_GROUPS];

struct SvBlockParams {
  double2 *state;
  const unsigned short *ctab;  // [ngroups][2^(kb-3)] swizzled tile index of every coset representative
  int n_local, kb, c, ngates, ngroups;
  unsigned char bits[16];      // ascending slab positions of the tile bits; bits[0..c-1] = 0..c-1
  unsigned short ainv[16];     // columns of A^-1 at the end of the pass (tile position of logical bit q)
  int framed;                  // 0: the frame is the identity at the end of the pass
  // TMA path: the tile is the box (run 0 = slab bits [0, a0)) x (run 1 = slab bits [s1, s1 + l1)) of a 5-d tensor
  // map over the slab, repeated for every combination of the nh remaining (higher) tile bits hib[]
  int a0, s1, l1, nh;
  unsigned char hib[16];
};

__device__ __forceinline__ double2 sv_cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 sv_cfma(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ uint32_t sv_ins0(uint32_t t, int b) {
  return ((t >> b) << (b + 1)) | (t & ((1u << b) - 1u));
}
// Shared-memory tiles are stored swizzled: the low three index bits are xor-ed with every higher
// 3-bit group of the index, so index bit p lands on bank-group bit p mod 3.  A group's eight lanes
// of a quarter-warp walk the coset counter bits 0..2; the host picks (SvGroup::ls) three non-register
// tile bits with distinct p mod 3 for them, which makes every register-slot access conflict-free
// whatever the target bits are (un-swizzled, gates on tile bits 0..2 were 8-way conflicted).
__device__ __forceinline__ uint32_t sv_swz(uint32_t j) {
  uint32_t f = j >> 3;
  f ^= f >> 3;
  f ^= f >> 6;
  return j ^ (f & 7u);
}
// TMA path: the layout cp.async.bulk.tensor writes with CU_TENSOR_MAP_SWIZZLE_128B -- the 16-byte chunk index inside
// a 128-byte row is xor-ed with the row index mod 8 (xor-linear and an involution, like sv_swz)
__device__ __forceinline__ uint32_t sv_swz_tma(uint32_t j) { return j ^ ((j >> 3) & 7u); }
template <bool TMA>
__device__ __forceinline__ uint32_t sv_sw(uint32_t j) {
  if constexpr (TMA) return sv_swz_tma(j);
  return sv_swz(j);
}
__device__ __forceinline__ uint32_t sv_swapbits(uint32_t t, int a, int b) {
  const uint32_t x = ((t >> a) ^ (t >> b)) & 1u;
  return t ^ ((x << a) | (x << b));
}
// target on register slot bit Q; cm = mask of the register slot bit that controls the gate (0: none)
template <int Q>
__device__ __forceinline__ void sv_apply_slot(double2 (&a)[8], const double2 (&u)[4], uint32_t cm) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if ((c >> Q) & 1) continue;
    if ((c & cm) != cm) continue;
    const int c1 = c | (1 << Q);
    const double2 a0 = a[c], a1 = a[c1];
    a[c] = sv_cfma(u[1], a1, sv_cmul(u[0], a0));
    a[c1] = sv_cfma(u[3], a1, sv_cmul(u[2], a0));
  }
}
// uncontrolled 2x2 on register slot bit Q (fast path of simple groups)
template <int Q>
__device__ __forceinline__ void sv_apply_slot_nc(double2 (&a)[8], const double2 (&u)[4]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if ((c >> Q) & 1) continue;
    const int c1 = c | (1 << Q);
    const double2 a0 = a[c], a1 = a[c1];
    a[c] = sv_cfma(u[1], a1, sv_cmul(u[0], a0));
    a[c1] = sv_cfma(u[3], a1, sv_cmul(u[2], a0));
  }
}
// (controlled) Pauli-X: exchange the pair, no arithmetic
template <int Q>
__device__ __forceinline__ void sv_swap_slot(double2 (&a)[8], uint32_t cm) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if ((c >> Q) & 1) continue;
    const bool sw = (c & cm) == cm;
    const int c1 = c | (1 << Q);
    const double2 a0 = a[c], a1 = a[c1];
    a[c] = sw ? a1 : a0;
    a[c1] = sw ? a0 : a1;
  }
}

// ---- TMA / mbarrier primitives (PTX as issued by CUTLASS' SM90_TMA_LOAD_5D / SM90_TMA_STORE_5D / ClusterBarrier) ----
__device__ __forceinline__ uint32_t sv_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sv_mbar_init(unsigned long long *mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(sv_smem_u32(mbar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sv_mbar_expect_tx(unsigned long long *mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(sv_smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void sv_mbar_wait(unsigned long long *mbar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "SV_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra SV_DONE;\n\t"
      "bra SV_WAIT;\n\t"
      "SV_DONE:\n\t"
      "}" ::"r"(sv_smem_u32(mbar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void sv_tma_load_5d(const void *tmap, unsigned long long *mbar, void *dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(sv_smem_u32(dst)), "l"((unsigned long long)tmap), "r"(sv_smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void sv_tma_store_5d(const void *tmap, const void *src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"((unsigned long long)tmap), "r"(sv_smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void sv_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the gate groups of a pass on one shared-memory tile (all threads of the CTA; ends with a CTA barrier) ----
template <bool TMA, int NT>
__device__ __forceinline__ void sv_apply_groups(double2 *tile, const SvBlockParams &P, unsigned long long base) {
  const int kb = P.kb;
  const uint32_t tsize = 1u << kb;
  for (int gr = 0; gr < P.ngroups; ++gr) {
    const SvGroup &grp = c_sv_groups[gr];
    const uint32_t s1 = grp.m[0], s2 = grp.m[1], s4 = grp.m[2];
    const uint32_t w1 = grp.w[0], w2 = grp.w[1], w4 = grp.w[2];           // the swizzle is xor-linear
    const int simple = grp.simple;
    const unsigned short *ctab = P.ctab + ((size_t)gr << (kb - 3));
    for (uint32_t t = threadIdx.x; t < (tsize >> 3); t += NT) {
      // coset representative from the host-built table (the index arithmetic -- bit swaps, zero
      // insertion, swizzle -- was a third of the kernel's instructions)
      const uint32_t Ps = __ldg(ctab + t);
      uint32_t sm[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) sm[q] = Ps ^ ((q & 1) ? w1 : 0u) ^ ((q & 2) ? w2 : 0u) ^ ((q & 4) ? w4 : 0u);
      double2 a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = tile[sm[q]];
      if (simple) {
        // at most one uncontrolled 2x2 per register slot, in slot order: no descriptor decoding
        const SvTileGate *gt = c_sv_gates + grp.first;
        if (simple & 1) {
          double2 u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) u[e] = gt->u[e];
          sv_apply_slot_nc<0>(a, u);
          ++gt;
        }
        if (simple & 2) {
          double2 u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) u[e] = gt->u[e];
          sv_apply_slot_nc<1>(a, u);
          ++gt;
        }
        if (simple & 4) {
          double2 u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) u[e] = gt->u[e];
          sv_apply_slot_nc<2>(a, u);
        }
      } else {
        const uint32_t Pj = sv_sw<TMA>(Ps);   // the swizzle is an involution
        for (int gi = grp.first; gi < grp.first + grp.ngates; ++gi) {
          const SvTileGate *gt = c_sv_gates + gi;
          const int kind = gt->kind;
          if (kind != SV_K_PDIAG) {
            const int cpo = gt->cpos_out;
            const uint32_t crm = gt->crmask;
            const bool on = (cpo < 0 || ((base >> cpo) & 1ull)) && (crm == 0u || (__popc(Pj & crm) & 1));
            if (on) {
              const uint32_t cm = gt->cslot >= 0 ? (1u << gt->cslot) : 0u;
              const int ts = gt->tslot;
              if (kind == SV_K_U1) {
                double2 u[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) u[e] = gt->u[e];
                if (ts == 0) sv_apply_slot<0>(a, u, cm);
                else if (ts == 1) sv_apply_slot<1>(a, u, cm);
                else sv_apply_slot<2>(a, u, cm);
              } else {
                if (ts == 0) sv_swap_slot<0>(a, cm);
                else if (ts == 1) sv_swap_slot<1>(a, cm);
                else sv_swap_slot<2>(a, cm);
              }
            }
          } else {
            const bool flip = (__popcll(base & gt->omask) ^ __popc(Pj & gt->tmask)) & 1;
            const double2 d0 = flip ? gt->u[3] : gt->u[0], d1 = flip ? gt->u[0] : gt->u[3];
            const uint32_t tm = gt->tmask;
            const uint32_t p1 = __popc(s1 & tm) & 1, p2 = __popc(s2 & tm) & 1, p4 = __popc(s4 & tm) & 1;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t par = ((q & 1) ? p1 : 0u) ^ ((q & 2) ? p2 : 0u) ^ ((q & 4) ? p4 : 0u);
              a[q] = sv_cmul(par ? d1 : d0, a[q]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) tile[sm[q]] = a[q];
    }
    __syncthreads();
  }
}
// Write a tile back with the frame of the pass undone: logical index j sits at tile position A^-1 j.  Thread
// `tid` stores the elements j = tid | (i << LB): their tile positions are sw(A^-1 tid) ^ sw(A^-1 (i << LB)) and their
// slab offsets offs(tid) ^ offs(i << LB) (both maps are xor-linear), so the i-dependent halves come from two small
// per-pass tables (spk / goff, filled once per CTA by sv_fill_store_tables) instead of being recomputed per element.
template <int NT>
struct SvStoreTables {
  static constexpr int LB = NT == 512 ? 9 : 8;
  uint32_t spk[1 << (SV_MAX_TILE_BITS - 8)];
  unsigned long long goff[1 << (SV_MAX_TILE_BITS - 8)];
};
template <bool TMA, int NT>
__device__ __forceinline__ void sv_fill_store_tables(SvStoreTables<NT> &tb, const SvBlockParams &P, const unsigned long long *offs) {
  constexpr int LB = SvStoreTables<NT>::LB;
  const int nhi = P.kb > LB ? 1 << (P.kb - LB) : 1;
  for (int i = threadIdx.x; i < nhi; i += NT) {
    uint32_t pk = 0;
    for (int q = LB; q < P.kb; ++q) pk ^= ((i >> (q - LB)) & 1) ? (uint32_t)P.ainv[q] : 0u;
    tb.spk[i] = sv_sw<TMA>(pk);
    tb.goff[i] = LB >= P.c ? offs[(uint32_t)i << (LB - P.c)] : 0ull;
  }
}
template <bool TMA, int NT>
__device__ __forceinline__ void sv_store_framed(const double2 *tile, double2 *g, const SvBlockParams &P, const unsigned long long *offs,
                                                const SvStoreTables<NT> &tb) {
  constexpr int LB = SvStoreTables<NT>::LB;
  const int c = P.c;
  const uint32_t tsize = 1u << P.kb, lowmask = (1u << c) - 1u;
  if (threadIdx.x >= tsize) return;              // tiles smaller than the CTA (test sizes)
  uint32_t pt = 0;
#pragma unroll
  for (int q = 0; q < LB; ++q) pt ^= ((threadIdx.x >> q) & 1u) ? (uint32_t)P.ainv[q] : 0u;
  const uint32_t sw0 = sv_sw<TMA>(pt);
  double2 *g0 = g + ((threadIdx.x & lowmask) | offs[threadIdx.x >> c]);
  const int nhi = P.kb > LB ? 1 << (P.kb - LB) : 1;
#pragma unroll 4
  for (int i = 0; i < nhi; ++i) g0[tb.goff[i]] = tile[sw0 ^ tb.spk[i]];
}
__device__ __forceinline__ unsigned long long sv_tile_base(const SvBlockParams &P, unsigned long long tid_) {
  unsigned long long base = 0, rest = tid_;
  int q = 0;
  for (int pos = 0; pos < P.n_local; ++pos) {
    if (q < P.kb && P.bits[q] == pos) { ++q; continue; }
    base |= (rest & 1ull) << pos;
    rest >>= 1;
  }
  return base;
}
// one warp issues / stores the TMA boxes of a tile
template <bool STORE>
__device__ __forceinline__ void sv_tma_tile(const CUtensorMap *tmap, const SvBlockParams &P, unsigned long long base, double2 *tile,
                                            unsigned long long *mbar) {
  const int ncombo = 1 << P.nh;
  const uint32_t box_amps = 1u << (P.a0 + P.l1);
  for (int ci = threadIdx.x & 31; ci < ncombo; ci += 32) {
    unsigned long long hi = base;
    for (int q = 0; q < P.nh; ++q) hi |= (unsigned long long)((ci >> q) & 1) << P.hib[q];
    const int c2 = (int)((hi >> P.a0) & ((1ull << (P.s1 - P.a0)) - 1ull));
    const int c4 = (int)(hi >> (P.s1 + P.l1));
    if constexpr (STORE) sv_tma_store_5d(tmap, tile + (size_t)ci * box_amps, 0, 0, c2, 0, c4);
    else sv_tma_load_5d(tmap, mbar, tile + (size_t)ci * box_amps, 0, 0, c2, 0, c4);
  }
}

// Persistent CTAs, one tile at a time.  Tile element j (kb bits) lives at slab index base | spread(j): the low c
// bits are contiguous (coalesced 16*2^c-byte runs), the others are scattered by offs[].
// TMA = true: the tile is fetched by cp.async.bulk.tensor from a 5-d tensor map of the slab (one box per
// combination of the tile bits beyond its first two runs of consecutive positions), completion on an mbarrier,
// shared-memory layout = the TMA 128-byte swizzle; a pass that ends with the identity frame is written back with
// cp.async.bulk.tensor as well, a framed pass by the threads (the frame is undone by the store addresses).
// TMA = false: cp.async staging with the fold-all swizzle (tiles the tensor map cannot describe, test sizes).
// dynamic shared memory: tile [2^kb] double2 (1024-byte aligned for the TMA swizzle)
template <int MINB, bool TMA>
__global__ void __launch_bounds__(SV_THREADS, MINB) sv_block_kernel(const SvBlockParams P, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) double2 tile[];
  __shared__ unsigned long long offs[1 << 8];
  __shared__ unsigned long long s_base;
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ SvStoreTables<SV_THREADS> s_tb;
  const int kb = P.kb, c = P.c;
  const uint32_t tsize = 1u << kb, nhi = 1u << (kb - c), lowmask = (1u << c) - 1u;
  for (uint32_t h = threadIdx.x; h < nhi; h += SV_THREADS) {
    unsigned long long o = 0;
    for (int q = c; q < kb; ++q) o |= (unsigned long long)((h >> (q - c)) & 1u) << P.bits[q];
    offs[h] = o;
  }
  if constexpr (TMA) {
    if (threadIdx.x == 0) sv_mbar_init(&s_mbar, 1);
  }
  __syncthreads();
  if (P.framed) sv_fill_store_tables<TMA, SV_THREADS>(s_tb, P, offs);
  uint32_t phase = 0;
  const unsigned long long ntiles = 1ull << (P.n_local - kb);
  for (unsigned long long tid_ = blockIdx.x; tid_ < ntiles; tid_ += gridDim.x) {
    if (threadIdx.x == 0) s_base = sv_tile_base(P, tid_);
    __syncthreads();
    const unsigned long long base = s_base;
    double2 *g = P.state + base;
    if constexpr (TMA) {
      // one warp issues the tile's boxes; every thread then waits on the mbarrier's phase
      if (threadIdx.x < 32) {
        if (threadIdx.x == 0) sv_mbar_expect_tx(&s_mbar, 16u << kb);
        __syncwarp();
        sv_tma_tile<false>(&tmap, P, base, tile, &s_mbar);
      }
      sv_mbar_wait(&s_mbar, phase);
      phase ^= 1u;
    } else {
      // HBM -> shared memory with cp.async: every thread has all of its 16-byte copies in flight at once
      const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
      const uint32_t sw0 = sv_swz(threadIdx.x) * 16u;
#pragma unroll 4
      for (uint32_t k = 0; k < tsize; k += SV_THREADS) {   // swz(tid | k) = swz(tid) ^ swz(k): disjoint bits
        const uint32_t j = threadIdx.x | k;
        if (j >= tsize) break;               // tiles smaller than the CTA (test sizes)
        const double2 *src = g + ((j & lowmask) | offs[j >> c]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile_s + (sw0 ^ (sv_swz(k) * 16u))), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    sv_apply_groups<TMA, SV_THREADS>(tile, P, base);
    if (!P.framed) {
      if constexpr (TMA) {
        // shared memory (generic-proxy writes) -> async proxy, then one warp stores the boxes back
        sv_fence_async_smem();
        __syncthreads();
        if (threadIdx.x < 32) {
          sv_tma_tile<true>(&tmap, P, base, tile, nullptr);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the tile may be overwritten again
        }
      } else {
        const uint32_t sw0 = sv_swz(threadIdx.x);
#pragma unroll 4
        for (uint32_t k = 0; k < tsize; k += SV_THREADS) {
          const uint32_t j = threadIdx.x | k;
          if (j >= tsize) break;
          g[(j & lowmask) | offs[j >> c]] = tile[sw0 ^ sv_swz(k)];
        }
      }
    } else {
      sv_store_framed<TMA, SV_THREADS>(tile, g, P, offs, s_tb);
      if constexpr (TMA) sv_fence_async_smem();     // these reads precede the next tile's async-proxy writes
    }
    __syncthreads();
  }
  if constexpr (TMA) {
    if (threadIdx.x < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA exits
  }
}

// Software-pipelined TMA variant: ONE CTA of 512 threads per SM owns TWO tile buffers; while the threads apply the
// gate groups to tile k, the TMA unit is already fetching tile k + 1 into the other buffer (its own mbarrier), so
// HBM reads overlap the arithmetic inside every SM by construction -- with single-buffered CTAs all tiles cost the
// same and the CTAs of the whole grid fall into lock-step (everybody loads, then everybody computes), leaving HBM
// idle during the arithmetic and the FP64 pipe idle during the copies.
// dynamic shared memory: 2 x tile [2^kb] double2
#define SV_PIPE_THREADS 512
__global__ void __launch_bounds__(SV_PIPE_THREADS, 1) sv_pipe_kernel(const SvBlockParams P, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) double2 tiles[];
  __shared__ unsigned long long offs[1 << 8];
  __shared__ unsigned long long s_base[2];
  __shared__ __align__(8) unsigned long long s_mbar[2];
  __shared__ SvStoreTables<SV_PIPE_THREADS> s_tb;
  const int kb = P.kb, c = P.c;
  const uint32_t tsize = 1u << kb, nhi = 1u << (kb - c);
  for (uint32_t h = threadIdx.x; h < nhi; h += SV_PIPE_THREADS) {
    unsigned long long o = 0;
    for (int q = c; q < kb; ++q) o |= (unsigned long long)((h >> (q - c)) & 1u) << P.bits[q];
    offs[h] = o;
  }
  const unsigned long long ntiles = 1ull << (P.n_local - kb);
  if (threadIdx.x == 0) {
    sv_mbar_init(&s_mbar[0], 1);
    sv_mbar_init(&s_mbar[1], 1);
  }
  __syncthreads();
  if (P.framed) sv_fill_store_tables<true, SV_PIPE_THREADS>(s_tb, P, offs);
  // prologue: fetch this CTA's first tile
  if (threadIdx.x < 32 && blockIdx.x < ntiles) {
    unsigned long long b0 = 0;
    if (threadIdx.x == 0) { b0 = sv_tile_base(P, blockIdx.x); s_base[0] = b0; sv_mbar_expect_tx(&s_mbar[0], 16u << kb); }
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    sv_tma_tile<false>(&tmap, P, b0, tiles, &s_mbar[0]);
  }
  uint32_t it = 0;
  for (unsigned long long tid_ = blockIdx.x; tid_ < ntiles; tid_ += gridDim.x, ++it) {
    const int cur = it & 1, nxt = cur ^ 1;
    double2 *tile = tiles + (size_t)cur * tsize;
    // prefetch the next tile into the other buffer (free since the barrier that ended iteration it - 1)
    const unsigned long long tnext = tid_ + gridDim.x;
    if (threadIdx.x < 32 && tnext < ntiles) {
      unsigned long long bn = 0;
      if (threadIdx.x == 0) { bn = sv_tile_base(P, tnext); s_base[nxt] = bn; sv_mbar_expect_tx(&s_mbar[nxt], 16u << kb); }
      bn = __shfl_sync(0xffffffffu, bn, 0);
      sv_tma_tile<false>(&tmap, P, bn, tiles + (size_t)nxt * tsize, &s_mbar[nxt]);
    }
    sv_mbar_wait(&s_mbar[cur], (it >> 1) & 1u);
    const unsigned long long base = s_base[cur];     // written before the barrier that preceded this tile's wait
    double2 *g = P.state + base;
    sv_apply_groups<true, SV_PIPE_THREADS>(tile, P, base);
    if (!P.framed) {
      sv_fence_async_smem();
      __syncthreads();
      if (threadIdx.x < 32) {
        sv_tma_tile<true>(&tmap, P, base, tile, nullptr);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else {
      sv_store_framed<true, SV_PIPE_THREADS>(tile, g, P, offs, s_tb);
      sv_fence_async_smem();
    }
    __syncthreads();
  }
  if (threadIdx.x < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

namespace {
// The gate list and the groups of the pass being executed live in __constant__ symbols of this module: calls
// from different threads / streams would overwrite each other's lists.  qas_sv_apply serialises itself.
std::mutex g_sv_apply_mutex;

typedef CUresult (*SvEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
SvEncodeTiledFn sv_encode_fn() {
  static SvEncodeTiledFn fn = []() -> SvEncodeTiledFn {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return (SvEncodeTiledFn)p;
  }();
  return fn;
}

// Tile geometry for the TMA path: run 0 = slab bits [0, a0), run 1 = [s1, s1 + l1), the other tile bits hib[].
struct SvTmaGeom { int a0, s1, l1, nh; unsigned char hib[16]; };
bool sv_tma_geometry(const std::vector<int> &bits, int kb, int n_local, SvTmaGeom *gm) {
  if (kb < 7 || n_local < kb) return false;
  int a0 = 0;
  while (a0 < kb && a0 < 11 && bits[a0] == a0) ++a0;
  if (a0 < 3) return false;                        // a 128-byte row = tile bits 0..2
  int s1 = a0, l1 = 0;
  if (a0 < kb) {
    s1 = bits[a0];
    while (a0 + l1 < kb && l1 < 8 && bits[a0 + l1] == s1 + l1) ++l1;
  }
  gm->a0 = a0; gm->s1 = s1; gm->l1 = l1; gm->nh = kb - a0 - l1;
  for (int q = 0; q < 16; ++q) gm->hib[q] = 0;
  for (int q = 0; q < gm->nh; ++q) gm->hib[q] = (unsigned char)bits[a0 + l1 + q];
  return gm->nh <= 10;
}
// 5-d tensor map over the slab, innermost first: [16 doubles = bits 0..2][bits 3..a0-1][gap a0..s1-1][run 1][bits above run 1],
// box = (16, 2^(a0-3), 1, 2^l1, 1), 128-byte swizzle
bool sv_make_tensor_map(void *state, int n_local, const SvTmaGeom &gm, CUtensorMap *out) {
  SvEncodeTiledFn enc = sv_encode_fn();
  if (!enc) return false;
  const int top = gm.s1 + gm.l1;
  cuuint64_t gdim[5] = {16, 1ull << (gm.a0 - 3), 1ull << (gm.s1 - gm.a0), 1ull << gm.l1, 1ull << (n_local - top)};
  cuuint64_t gstride[4] = {128, 16ull << gm.a0, 16ull << gm.s1, 16ull << top};
  cuuint32_t box[5] = {16, 1u << (gm.a0 - 3), 1, 1u << gm.l1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, state, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
struct SvFreeList {       // device allocations and events of one call, released on every exit path
  cudaStream_t st;
  std::vector<void *> ptrs;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ~SvFreeList() {
    for (void *q : ptrs) cudaFreeAsync(q, st);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};
}  // namespace

extern "C" int qas_sv_apply(void *state, int n_local, const qas_sv_gate *gates, int ngates, int tile_bits,
                            void *stream, int sync, qas_sv_info *info) {
  std::lock_guard<std::mutex> serialise(g_sv_apply_mutex);
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
  kb = std::max(3, std::min(std::min(kb, SV_MAX_TILE_BITS), n_local));
  if (n_local < 3) { g_sv_error = "large-state mode needs at least 3 local qubits"; return QAS_ERR_UNSUPPORTED; }
  // c = contiguous low bits every tile contains (coalescing); leave at least one free tile bit
  int cwant = 5;
  if (const char *ce = getenv("QAS_B200_SV_C")) cwant = std::max(0, atoi(ce));   // tuning knob
  const int c = std::max(std::max(0, kb - 8), std::min(cwant, kb - 1));     // the kernel's offs[] table holds 2^(kb-c) <= 256 entries
  if (info) { info->passes = 0; info->tile_bits = kb; info->bytes_per_pass = 32.0 * (double)(1ull << n_local); info->last_ms = 0.0; }
  if (ngates == 0) return QAS_OK;
  std::vector<Block> blocks = plan_blocks(gates, ngates, n_local, kb, c);
  if (blocks.empty()) { g_sv_error = "gate blocking made no progress (tile_bits too small)"; return QAS_ERR_UNSUPPORTED; }

  // translate every block's gates into tile coordinates and register groups, upload once
  std::vector<SvTileGate> tg;
  std::vector<SvGroup> groups;
  std::vector<unsigned short> ctab;      // per group: swizzled tile index of every coset representative
  int use_tma = 1;
  if (const char *te = getenv("QAS_B200_SV_TMA")) use_tma = atoi(te) != 0;     // tuning knob: 0 = cp.async staging for every pass
  bool tma_mode = false;         // layout of the pass being translated
  auto h_swz = [&tma_mode](uint32_t j) {
    if (tma_mode) return j ^ ((j >> 3) & 7u);
    uint32_t f = j >> 3; f ^= f >> 3; f ^= f >> 6; return j ^ (f & 7u);
  };
  std::vector<SvTmaGeom> geom(blocks.size());
  std::vector<char> tma_of(blocks.size(), 0);
  auto h_par = [](uint32_t v) { return (uint32_t)(__builtin_popcount(v) & 1); };
  std::vector<int> first(blocks.size() + 1, 0), gfirst(blocks.size() + 1, 0);
  std::vector<std::array<unsigned short, 16>> ainv_of(blocks.size());
  std::vector<int> framed_of(blocks.size(), 0);
  int fold = 1;
  if (const char *fe = getenv("QAS_B200_SV_FOLD")) fold = atoi(fe) != 0;   // tuning knob: 0 executes every CNOT
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    const Block &blk = blocks[bi];
    tma_of[bi] = use_tma && sv_tma_geometry(blk.bits, kb, n_local, &geom[bi]) && sv_encode_fn() != nullptr;
    tma_mode = tma_of[bi] != 0;
    std::vector<int> tix(n_local, -1);
    for (size_t q = 0; q < blk.bits.size(); ++q) tix[blk.bits[q]] = (int)q;
    // frame of the pass: tile position p holds logical tile index A p.  arow[q] = row q of A (reads
    // logical bit q of a position), acol[q] = column q of A^-1 (xor mask that flips logical bit q).
    std::vector<uint32_t> arow(kb), acol(kb);
    for (int q = 0; q < kb; ++q) arow[q] = acol[q] = 1u << q;
    const int gate0 = (int)tg.size();
    std::vector<int> cur_bits;      // logical tile bits (targets) of the open group
    int cur_first = 0;              // first gate (index relative to the block) of the open group
    auto close_group = [&]() {
      int end_rel = (int)tg.size() - gate0;
      if (end_rel == cur_first) { cur_bits.clear(); return; }
      // pad the register bits with the highest free logical bits (low bits stay with the lanes)
      for (int b = kb - 1; b >= 0 && (int)cur_bits.size() < 3; --b)
        if (std::find(cur_bits.begin(), cur_bits.end(), b) == cur_bits.end()) cur_bits.push_back(b);
      std::sort(cur_bits.begin(), cur_bits.end());
      SvGroup gr{};
      gr.first = cur_first; gr.ngates = end_rel - cur_first;
      uint32_t rr[3];
      for (int q = 0; q < 3; ++q) { gr.m[q] = acol[cur_bits[q]]; gr.w[q] = h_swz(gr.m[q]); rr[q] = arow[cur_bits[q]]; }
      // slot indices of the group's gates; tile-bit controls outside the group become parity masks
      for (int r = cur_first; r < end_rel; ++r) {
        SvTileGate &t = tg[gate0 + r];
        if (t.kind == SV_K_PDIAG) continue;
        const int tb_ = t.tslot;    // holds the logical tile bit until the group closes
        t.tslot = (int)(std::find(cur_bits.begin(), cur_bits.end(), tb_) - cur_bits.begin());
        if (t.cslot >= 0) {         // holds the logical tile bit of the control
          const int cb_ = t.cslot;
          auto it = std::find(cur_bits.begin(), cur_bits.end(), cb_);
          if (it != cur_bits.end()) { t.cslot = (int)(it - cur_bits.begin()); t.crmask = 0u; }
          else { t.cslot = -1; t.crmask = arow[cb_]; }
        }
      }
      // coset representatives: the common kernel of the three read masks, as xor-span of a basis whose
      // first three vectors move the three bank-group bits of the swizzled index independently when
      // the kernel allows it (they are the lanes of a quarter-warp, see sv_swz)
      std::vector<uint32_t> kern;
      {
        // eliminate: x in kernel <=> rr[q].x = 0; rr's are independent with rr[a].m[b] = delta_ab, so
        // projecting the unit vectors along span(m) onto the kernel spans it
        std::vector<uint32_t> ech;      // echelon basis (highest-bit pivots) of the projections
        for (int b = 0; b < kb; ++b) {
          uint32_t x = 1u << b;
          for (int q = 0; q < 3; ++q) if (h_par(x & rr[q])) x ^= gr.m[q];
          for (uint32_t e : ech) if ((x ^ e) < x) x ^= e;
          if (x) { ech.push_back(x); std::sort(ech.begin(), ech.end(), std::greater<uint32_t>()); }
        }
        // pick three with independent low swizzled bits first
        std::vector<uint32_t> lead, rest, keys;
        for (uint32_t v : ech) {
          uint32_t key = h_swz(v) & 7u;
          for (uint32_t kk : keys) if ((key ^ kk) < key) key ^= kk;
          if (key && lead.size() < 3) { lead.push_back(v); keys.push_back(key); std::sort(keys.begin(), keys.end(), std::greater<uint32_t>()); }
          else rest.push_back(v);
        }
        kern = lead;
        std::sort(rest.begin(), rest.end());
        kern.insert(kern.end(), rest.begin(), rest.end());
      }
      // simple group: nothing but uncontrolled 2x2s -> one product per register slot, in slot order
      {
        bool simple = true;
        for (int r = cur_first; r < end_rel; ++r) {
          const SvTileGate &t = tg[gate0 + r];
          if (t.kind == SV_K_PDIAG || t.cslot >= 0 || t.crmask != 0u || t.cpos_out >= 0) simple = false;
        }
        if (simple) {
          double2 acc[3][4];
          bool has[3] = {false, false, false};
          for (int r = cur_first; r < end_rel; ++r) {
            const SvTileGate &t = tg[gate0 + r];
            const int q = t.tslot;
            if (!has[q]) { for (int e = 0; e < 4; ++e) acc[q][e] = t.u[e]; has[q] = true; continue; }
            double2 n4[4];                       // later gate times what is there: U_r * acc
            for (int i = 0; i < 2; ++i)
              for (int j = 0; j < 2; ++j) {
                const double2 x = t.u[2 * i], y = acc[q][j], z = t.u[2 * i + 1], w = acc[q][2 + j];
                n4[2 * i + j] = make_double2(x.x * y.x - x.y * y.y + z.x * w.x - z.y * w.y, x.x * y.y + x.y * y.x + z.x * w.y + z.y * w.x);
              }
            for (int e = 0; e < 4; ++e) acc[q][e] = n4[e];
          }
          tg.resize((size_t)gate0 + cur_first);
          for (int q = 0; q < 3; ++q) {
            if (!has[q]) continue;
            SvTileGate t{};
            t.kind = SV_K_U1; t.tslot = q; t.cslot = -1; t.crmask = 0u; t.cpos_out = -1;
            for (int e = 0; e < 4; ++e) t.u[e] = acc[q][e];
            tg.push_back(t);
            gr.simple |= 1 << q;
          }
          end_rel = (int)tg.size() - gate0;
          gr.ngates = end_rel - cur_first;
        }
      }
      const int kd = (int)kern.size();       // = kb - 3
      for (uint32_t t = 0; t < (1u << (kb - 3)); ++t) {
        uint32_t P = 0;
        for (int i = 0; i < kd; ++i) if ((t >> i) & 1u) P ^= kern[i];
        ctab.push_back((unsigned short)h_swz(P));
      }
      groups.push_back(gr);
      cur_bits.clear();
      cur_first = end_rel;
    };
    for (int gi : blk.gates) {
      const qas_sv_gate &g = gates[gi];
      SvTileGate t{};
      t.kind = g.kind;
      t.tslot = -1; t.cslot = -1; t.crmask = 0u; t.cpos_out = -1;
      const bool is_x = g.kind == QAS_SV_U1 && g.m[0] == 0.0 && g.m[1] == 0.0 && g.m[6] == 0.0 && g.m[7] == 0.0 &&
                        g.m[2] == 1.0 && g.m[3] == 0.0 && g.m[4] == 1.0 && g.m[5] == 0.0;
      if (fold && is_x && g.control >= 0 && tix[g.control] >= 0) {
        // CNOT inside the tile: fold into the frame (A <- C A), nothing to execute
        close_group();
        const int qt = tix[g.target], qc = tix[g.control];
        arow[qt] ^= arow[qc];
        acol[qc] ^= acol[qt];
        framed_of[bi] = 1;
        continue;
      }
      if (g.kind == QAS_SV_U1) {
        const int tb_ = tix[g.target];
        if (std::find(cur_bits.begin(), cur_bits.end(), tb_) == cur_bits.end()) {
          if ((int)cur_bits.size() == 3) close_group();
          cur_bits.push_back(tb_);
        }
        t.tslot = tb_;
        if (g.control >= 0) { if (tix[g.control] >= 0) t.cslot = tix[g.control]; else t.cpos_out = g.control; }
      } else {
        uint32_t lmask = 0;
        for (int b = 0; b < n_local; ++b)
          if ((g.mask >> b) & 1ull) { if (tix[b] >= 0) lmask |= 1u << tix[b]; else t.omask |= 1ull << b; }
        for (int q = 0; q < kb; ++q) if ((lmask >> q) & 1u) t.tmask ^= arow[q];     // parity over positions
      }
      for (int q = 0; q < 4; ++q) t.u[q] = make_double2(g.m[2 * q], g.m[2 * q + 1]);
      if (is_x) t.kind = SV_K_SWAPX;        // PauliX / CNOT with an outside control: exchange the pair, no flops
      tg.push_back(t);
    }
    close_group();
    for (int q = 0; q < 16; ++q) ainv_of[bi][q] = q < kb ? (unsigned short)acol[q] : (unsigned short)0;
    first[bi + 1] = (int)tg.size();
    gfirst[bi + 1] = (int)groups.size();
  }
  SvFreeList owned{st};
  unsigned short *d_ctab = nullptr;
  SVCK(cudaMallocAsync((void **)&d_ctab, std::max<size_t>(ctab.size(), 1) * sizeof(unsigned short), st));
  owned.ptrs.push_back(d_ctab);
  SVCK(cudaMemcpyAsync(d_ctab, ctab.data(), ctab.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, st));
  if (info && sync) { SVCK(cudaEventCreate(&owned.e0)); SVCK(cudaEventCreate(&owned.e1)); SVCK(cudaEventRecord(owned.e0, st)); }
  size_t max_gates = 0;
  for (size_t bi = 0; bi < blocks.size(); ++bi) max_gates = std::max<size_t>(max_gates, first[bi + 1] - first[bi]);
  size_t max_groups = 0;
  for (size_t bi = 0; bi < blocks.size(); ++bi) max_groups = std::max<size_t>(max_groups, gfirst[bi + 1] - gfirst[bi]);
  const size_t smem_max = (size_t)16 << kb;
  int tma_passes = 0;
  if (max_gates > SV_CONST_GATES || max_groups > SV_CONST_GROUPS) { g_sv_error = "too many gates in one pass for the constant-memory gate list"; return QAS_ERR_UNSUPPORTED; }
  int minb = 3;
  if (const char *mb = getenv("QAS_B200_SV_MINB")) minb = atoi(mb) == 2 ? 2 : 3;   // tuning knob
  SVCK(cudaFuncSetAttribute(sv_block_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  SVCK(cudaFuncSetAttribute(sv_block_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  SVCK(cudaFuncSetAttribute(sv_block_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  SVCK(cudaFuncSetAttribute(sv_block_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  const unsigned long long ntiles = 1ull << (n_local - kb);
  // persistent CTAs: as many as stay resident, each walks its tiles
  int dev_id = 0, num_sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev_id);
  int occ = minb;
  if (minb == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sv_block_kernel<2, true>, SV_THREADS, smem_max);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sv_block_kernel<3, true>, SV_THREADS, smem_max);
  const int grid = (int)std::min<unsigned long long>(ntiles, (unsigned long long)std::max(1, occ) * num_sms);
  // double-buffered TMA pipeline (one 512-thread CTA per SM, two tile buffers).  Opt-in: measured SLOWER than three
  // single-buffered CTAs per SM (28 qubits: 49.3 vs 39.4 ms per 12 passes) -- the pass is bound by instruction issue
  // and by the CTA barrier after every gate group, and three independent CTAs hide each other's barriers.
  int use_pipe = 0;
  if (const char *pe = getenv("QAS_B200_SV_PIPE")) use_pipe = atoi(pe) != 0;      // tuning knob
  const bool pipe_fits = 2 * smem_max <= 200 * 1024 && kb >= 9;
  if (use_pipe && pipe_fits) SVCK(cudaFuncSetAttribute(sv_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem_max)));
  const int grid_pipe = (int)std::min<unsigned long long>(ntiles, (unsigned long long)num_sms);
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    SvBlockParams P{};
    P.state = (double2 *)state;
    P.ctab = d_ctab + ((size_t)gfirst[bi] << (kb - 3));
    P.n_local = n_local; P.kb = kb; P.c = c; P.ngates = first[bi + 1] - first[bi];
    P.ngroups = gfirst[bi + 1] - gfirst[bi];
    for (int q = 0; q < kb; ++q) P.bits[q] = (unsigned char)blocks[bi].bits[q];
    for (int q = 0; q < 16; ++q) P.ainv[q] = ainv_of[bi][q];
    P.framed = framed_of[bi];
    if (P.ngroups == 0 && !P.framed) continue;      // (cannot happen: a pass without gates)
    const size_t smem = (size_t)16 << kb;
    // stream-ordered: the copy waits for the previous pass's kernel, which still reads the old list
    SVCK(cudaMemcpyToSymbolAsync(c_sv_gates, tg.data() + first[bi], (size_t)P.ngates * sizeof(SvTileGate), 0, cudaMemcpyHostToDevice, st));
    SVCK(cudaMemcpyToSymbolAsync(c_sv_groups, groups.data() + gfirst[bi], (size_t)P.ngroups * sizeof(SvGroup), 0, cudaMemcpyHostToDevice, st));
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof tmap);
    bool tma = tma_of[bi] != 0;
    if (tma) {
      P.a0 = geom[bi].a0; P.s1 = geom[bi].s1; P.l1 = geom[bi].l1; P.nh = geom[bi].nh;
      for (int q = 0; q < 16; ++q) P.hib[q] = geom[bi].hib[q];
      if (!sv_make_tensor_map(state, n_local, geom[bi], &tmap)) { g_sv_error = "cuTensorMapEncodeTiled failed for a tile geometry"; return QAS_ERR_CUDA; }
    }
    if (tma && use_pipe && pipe_fits) {
      sv_pipe_kernel<<<grid_pipe, SV_PIPE_THREADS, 2 * smem, st>>>(P, tmap);
      ++tma_passes;
    } else if (tma) {
      if (minb == 2) sv_block_kernel<2, true><<<grid, SV_THREADS, smem, st>>>(P, tmap);
      else sv_block_kernel<3, true><<<grid, SV_THREADS, smem, st>>>(P, tmap);
      ++tma_passes;
    } else {
      if (minb == 2) sv_block_kernel<2, false><<<grid, SV_THREADS, smem, st>>>(P, tmap);
      else sv_block_kernel<3, false><<<grid, SV_THREADS, smem, st>>>(P, tmap);
    }
    SVCK(cudaGetLastError());
  }
  // the upload buffer must outlive the kernels: tg is host memory copied asynchronously from a
  // pageable vector, so wait for the copy before it goes out of scope
  if (info && sync) SVCK(cudaEventRecord(owned.e1, st));
  SVCK(cudaStreamSynchronize(st));
  if (info) {
    info->passes = (int)blocks.size();
    info->tma_passes = tma_passes;
    if (sync) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, owned.e0, owned.e1);
      info->last_ms = ms;
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
