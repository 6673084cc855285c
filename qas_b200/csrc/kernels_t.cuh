// kernels_t.cuh -- tiled evaluation kernel for full-register circuits on 13..16 qubits (sm_100a).
//
// A 2^n complex128 state (and, for gradients, lambda = H psi) of 13-16 qubits does not fit the shared memory of
// one SM (n = 16: 1 MiB + 1 MiB).  The round-1 fallback (eval_kernel<N = 0>) kept both vectors in a global slab and
// streamed the WHOLE slab through L2 for every pass of <= 2 gates: kagome-16q (p = 500, ~125 dense gates per
// circuit) moved ~0.8 GB per evaluation and ran at the L2 bandwidth of one SM (13 k evals/s on 148 SMs).
//
// This kernel keeps the slab but blocks the tape into TILE PASSES, the way the large-state kernel (sv.cu) does:
//   * a pass is a run of consecutive tape ops whose xor masks, together with the CLOW lowest index bits, span at
//     most TB dimensions of GF(2)^n.  The pass owns an invertible frame A = [tile basis | tile-number vectors]:
//     pass coordinate (t, tau) <-> physical index A (t | tau << TB).  A tile (2^TB amplitudes, fixed tau) is
//     closed under every xor mask of the pass, so all of the pass's gates run on it in shared memory and the slab
//     is read and written ONCE per pass instead of once per gate pair.  Ops are translated to pass coordinates on
//     the fly (m' = A^-1 m, r' = A^T r: the same change of basis as the c-tapes of tape.h); the part of a parity
//     mask that falls on the tile-number bits is a per-tile constant (a role flip / a constant control).
//   * the CLOW lowest physical bits are always tile coordinates 0..CLOW-1 and every other basis vector is free of
//     them, so a tile is a set of 2^CLOW-amplitude (64-byte) contiguous chunks: sector-exact global accesses.
//   * passes are planned greedily, by one thread, right before they run (qas_plan_tile_pass, tape.h: host/device,
//     tested on the CPU); the forward sweep plans from the end of the tape, the adjoint sweep from its start --
//     any partition of the tape into runs is valid for either sweep, so nothing is stored between them.
//   * a thread-block CLUSTER of CS CTAs evaluates one circuit: the tiles of a pass are dealt round-robin to the
//     CTAs, the hardware cluster barrier separates passes (it orders the slab accesses of the cluster), and the
//     2x2 cross matrices / energy partials of the CTAs are combined through distributed shared memory in fixed
//     rank order.  Fewer circuits in flight keep the slabs L2-resident (37 circuits x 2 MiB at CS = 4 against
//     296 MiB with one CTA per circuit and two CTAs per SM).
//     Why the state itself is NOT distributed over the cluster's shared memories: measured DSMEM bandwidth is
//     ~20 B/clk per SM (B300_MICROARCH.md, "CGA; DSMEM"), below what one SM draws from L2; a gate whose mask
//     crosses the CTA bits would move 64 KB per CTA through it, as slow as the L2 round trip it replaces.
// Arithmetic, gate conventions, cross-matrix adjoint and the dense H application are those of kernels.cuh.
// What the reference does instead: lightning.qubit's adjoint differentiation on the CPU, one circuit at a time
// (qas/qml_models/kagome_heisenberg_model.py:54-64, qchem_model.py:36-60).
#pragma once
#include <cooperative_groups.h>
#include "kernels.cuh"

#define QAS_T_MAXP 32          // ops per pass (shared-memory staging of their matrices / cross matrices)
#define QAS_T_KBW 12           // words per coset basis (tile bits - 1 at most)

__host__ __device__ inline size_t evalt_smem_bytes(int tb, int nt, bool grad, int ops_cap, int C) {
  const int nw = nt / 32;
  size_t b = ((size_t)16 << tb) * (grad ? 2 : 1);                        // psi tile | lambda tile
  b += (((size_t)ops_cap * QAS_OP_WORDS * 4) + 15) & ~(size_t)15;       // the circuit's tape
  b += (size_t)QAS_T_MAXP * 64;                                          // gate matrices of the pass
  b += (size_t)QAS_T_MAXP * 16;                                          // translated ops of the pass
  b += (size_t)QAS_T_MAXP * QAS_T_KBW * 4;                               // coset bases of the pass's groups
  if (grad) b += (size_t)QAS_T_MAXP * nw * 64 + (size_t)2 * QAS_T_MAXP * 64;   // per-warp cross matrices | CTA sums (two buffers)
  b += 2 * QAS_TP_MAXQ * 4;                                              // frame columns, pivot table
  b += (((size_t)nw * 8) + 15) & ~(size_t)15;                            // energy partials of the warps
  b += 64;                                                               // control words
  b += (((size_t)C * 4) + 15) & ~(size_t)15;                             // class descriptors of the H application
  return b;
}

struct TCtl { int q, cnt, j0, pad; double epart; };

// ---- one forward pass group on a tile: G gates on the 2^G amplitudes of every coset
// coset representative machinery shared by the two sweeps: the group's basis of K (qas_tile_group_basis, tape.h) is
// folded into one word per thread (the bits of tid select kb[0 .. LT-1]) and one word per loop iteration
template <int G, int TB, int NT>
struct TGeom {
  static constexpr int LT = (NT == 128) ? 7 : (NT == 256) ? 8 : (NT == 512) ? 9 : 10;
  static constexpr int NIT = (1 << (TB - G)) / NT;            // cosets per thread (1, 2 or 4)
  uint32_t m[G], xm[1 << G], base, k8, k9;     // k8, k9: the basis vectors selected by the iteration counter
  __device__ __forceinline__ void load(const uint4 *top, const uint32_t *kb, uint32_t tau, int tid) {
    uint32_t off = 0u;
#pragma unroll
    for (int q = 0; q < G; ++q) {
      m[q] = top[q].x;
      off ^= (__popc(tau & (top[q].y >> TB)) & 1) ? m[q] : 0u;      // tile-number part of the parity mask: a role flip
    }
    xm[0] = 0u;
#pragma unroll
    for (int c = 1; c < (1 << G); ++c) xm[c] = xm[c & ~(c >= 2 ? 2 : 1)] ^ m[c >= 2 ? 1 : 0];
    uint32_t b = off;
#pragma unroll
    for (int i = 0; i < LT; ++i) b ^= ((tid >> i) & 1) ? kb[i] : 0u;
    base = b;
    k8 = NIT > 1 ? kb[LT] : 0u;
    k9 = NIT > 2 ? kb[LT + 1] : 0u;
  }
  __device__ __forceinline__ uint32_t rep(int it) const { return base ^ ((it & 1) ? k8 : 0u) ^ ((it & 2) ? k9 : 0u); }
};

// ---- one forward pass group on a tile: G gates on the 2^G amplitudes of every coset
template <int G, int TB, int NT>
__device__ __forceinline__ void t_fwd_group(double2 *tile, const uint4 *top, const uint32_t *kb, const double2 *Us, uint32_t tau,
                                            int tid) {
  TGeom<G, TB, NT> gg;
  gg.load(top, kb, tau, tid);
  double2 u[G][4];
#pragma unroll
  for (int q = 0; q < G; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) u[q][e] = Us[4 * q + e];
  const uint32_t rc = (G == 1) ? top[0].z : 0u;
  const bool cflip = (G == 1) && (__popc(tau & (rc >> TB)) & 1);
  const uint32_t rcl = rc & ((1u << TB) - 1u);
#pragma unroll
  for (int it = 0; it < TGeom<G, TB, NT>::NIT; ++it) {
    const uint32_t Pp = gg.rep(it);
    if (G == 1 && rc && (((__popc(Pp & rcl) & 1) != 0) == cflip)) continue;     // control parity 0: the gate does not fire
    double2 a[1 << G];
#pragma unroll
    for (int s = 0; s < (1 << G); ++s) a[s] = tile[Pp ^ gg.xm[s]];
#pragma unroll
    for (int q = G - 1; q >= 0; --q) {       // time order = descending tape index
#pragma unroll
      for (int s = 0; s < (1 << G); ++s)
        if (!((s >> q) & 1)) apply_u(a[s], a[s | (1 << q)], u[q]);
    }
#pragma unroll
    for (int s = 0; s < (1 << G); ++s) tile[Pp ^ gg.xm[s]] = a[s];
  }
}

// ---- one adjoint pass group on a tile (tape order): psi <- U^ psi ; M += conj(lam) psi ; lam <- U^ lam.
// The cross matrices are reduced over the warp and ADDED to the warp's own slots Mw[op][warp][8] (one writer lane
// per slot: no atomics, fixed order); the CTA sums the warps after the last tile of the pass.
template <int G, int TB, int NT>
__device__ __forceinline__ void t_adj_group(double2 *tile, double2 *ltile, const uint4 *top, const uint32_t *kb, const double2 *Us,
                                            uint32_t tau, int tid, int lane, double *mw /* this warp's [8] of the group's first op */,
                                            int mw_op_stride) {
  TGeom<G, TB, NT> gg;
  gg.load(top, kb, tau, tid);
  int npar[G];
#pragma unroll
  for (int q = 0; q < G; ++q) npar[q] = qas_meta_npar(top[q].w);
  const uint32_t rc = (G == 1) ? top[0].z : 0u;
  const bool cflip = (G == 1) && (__popc(tau & (rc >> TB)) & 1);
  const uint32_t rcl = rc & ((1u << TB) - 1u);
  double M[G][8];
#pragma unroll
  for (int q = 0; q < G; ++q)
#pragma unroll
    for (int e = 0; e < 8; ++e) M[q][e] = 0.0;
  for (int it = 0; it < TGeom<G, TB, NT>::NIT; ++it) {
    const uint32_t Pp = gg.rep(it);
    if (G == 1 && rc && (((__popc(Pp & rcl) & 1) != 0) == cflip)) continue;
    double2 a[1 << G], l[1 << G];
#pragma unroll
    for (int s = 0; s < (1 << G); ++s) {
      a[s] = tile[Pp ^ gg.xm[s]];
      l[s] = ltile[Pp ^ gg.xm[s]];
    }
#pragma unroll
    for (int q = 0; q < G; ++q) {
      double2 u[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) u[e] = Us[4 * q + e];
#pragma unroll
      for (int s = 0; s < (1 << G); ++s) {
        if ((s >> q) & 1) continue;
        const int s1 = s | (1 << q);
        apply_udag(a[s], a[s1], u);
        if (npar[q]) macc(M[q], l[s], l[s1], a[s], a[s1]);
        apply_udag(l[s], l[s1], u);
      }
    }
#pragma unroll
    for (int s = 0; s < (1 << G); ++s) {
      tile[Pp ^ gg.xm[s]] = a[s];
      ltile[Pp ^ gg.xm[s]] = l[s];
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < G; ++q)
    if (npar[q]) reduce_store_m<32, true>(M[q], lane, 0xffffffffu, mw + (size_t)q * mw_op_stride, true);
}

// ---- dense H application on the slab: lambda = H psi for the cosets u = u0, u0 + ustride, ..; returns this thread's
// share of Re <psi|lambda>.  Same blocking and sign tables as the dense branch of eval_kernel (kernels.cuh); psi is
// read past L1 (its lines were written by other SMs of the cluster).
template <int RB, int PD, bool GRAD>
__device__ __forceinline__ double t_hphase(const EvalParams &P, const double2 *psi, double2 *lam, const uint32_t *sdesc,
                                           uint32_t u0, uint32_t ustride, int n) {
  constexpr int NB = 1 << RB;
  const uint32_t nrest = (1u << n) >> RB;
  const char *psib = reinterpret_cast<const char *>(psi);
  const int C = P.C_real + P.C_imag;
  double e_part = 0.0;
  for (uint32_t u = u0; u < nrest; u += ustride) {
    const uint32_t sb16 = qas_deposit_rest(u, RB, P.rb0, P.rb1, P.rb2) << 4;
    double2 acc[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[j] = make_double2(0.0, 0.0);
    const double *vp = P.vtab + (size_t)u * (NB >= 2 ? 2 : 1);
    double ring[PD][NB];
#pragma unroll
    for (int d = 0; d < PD; ++d) load_vrow<NB>(ring[d], vp + ((size_t)d << n), nrest);
    for (int cls = 0; cls < C; ++cls) {
      const uint32_t r16 = sb16 ^ sdesc[cls];
      double2 a[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) a[j] = __ldcg(reinterpret_cast<const double2 *>(psib + (r16 ^ P.wc16.w[j])));
      const bool imag = cls >= P.C_real;
#pragma unroll
      for (int jw = 0; jw < NB; ++jw) {
        double v[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) v[j] = ring[jw % PD][j];
        load_vrow<NB>(ring[jw % PD], vp + ((size_t)(jw + PD) << n), nrest);
        hacc_row<NB>(acc, a, v, jw, imag);
      }
      vp += (size_t)NB << n;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const uint32_t off = sb16 ^ P.wc16.w[j];
      const double2 ai = __ldcg(reinterpret_cast<const double2 *>(psib + off));
      e_part = fma(ai.x, acc[j].x, fma(ai.y, acc[j].y, e_part));
      if constexpr (GRAD) __stcg(reinterpret_cast<double2 *>(reinterpret_cast<char *>(lam) + off), acc[j]);
    }
  }
  return e_part;
}

// One cluster of CS CTAs per circuit, persistent, fed from the global work queue.  CLOW = log2 amplitudes per
// contiguous chunk of a tile.  Requires n >= TB and 2^(n - TB) >= CS.
template <bool GRAD, int TB, int NT, int CS, int CLOW, int MINB>
__global__ void __launch_bounds__(NT, MINB) eval_tiled_kernel(const EvalParams P) {
  namespace cg = cooperative_groups;
  constexpr int NW = NT / 32;
  constexpr int LT = (NT == 128) ? 7 : (NT == 256) ? 8 : (NT == 512) ? 9 : 10;
  static_assert((1 << LT) == NT, "CTA size must be 128, 256, 512 or 1024");
  constexpr int EPT = (1 << TB) / NT;              // tile elements per thread
  constexpr uint32_t LOWMASK = (1u << CLOW) - 1u;
  const int n = P.n;
  const uint32_t DIM = 1u << n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned rank = 0;
  if constexpr (CS > 1) rank = cg::this_cluster().block_rank();
  const int cluster_id = blockIdx.x / CS;
  const uint32_t ntiles = 1u << (n - TB);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *base = smem_raw;
  double2 *tile = reinterpret_cast<double2 *>(base);
  base += (size_t)16 << TB;
  double2 *ltile = nullptr;
  if constexpr (GRAD) { ltile = reinterpret_cast<double2 *>(base); base += (size_t)16 << TB; }
  uint32_t *ops = reinterpret_cast<uint32_t *>(base);
  base += (((size_t)P.ops_cap * QAS_OP_WORDS * 4) + 15) & ~(size_t)15;
  double2 *Us = reinterpret_cast<double2 *>(base);
  base += (size_t)QAS_T_MAXP * 64;
  uint4 *top = reinterpret_cast<uint4 *>(base);                 // translated ops: m', A^T r, A^T rc, meta
  base += (size_t)QAS_T_MAXP * 16;
  uint32_t *gkb = reinterpret_cast<uint32_t *>(base);           // coset basis of every group, at its first op
  base += (size_t)QAS_T_MAXP * QAS_T_KBW * 4;
  double *Mw = nullptr, *Ms = nullptr;
  if constexpr (GRAD) {
    Mw = reinterpret_cast<double *>(base); base += (size_t)QAS_T_MAXP * NW * 64;
    Ms = reinterpret_cast<double *>(base); base += (size_t)2 * QAS_T_MAXP * 64;
  }
  uint32_t *col = reinterpret_cast<uint32_t *>(base);
  int *piv = reinterpret_cast<int *>(col + QAS_TP_MAXQ);
  base += 2 * QAS_TP_MAXQ * 4;
  double *ered = reinterpret_cast<double *>(base);
  base += (((size_t)NW * 8) + 15) & ~(size_t)15;
  TCtl *ctl = reinterpret_cast<TCtl *>(base);
  base += 64;
  uint32_t *sdesc = reinterpret_cast<uint32_t *>(base);
  for (int s = tid; s < P.C_real + P.C_imag; s += NT) sdesc[s] = P.cdesc[s];

  double2 *gpsi = P.gstate + (size_t)cluster_id * DIM * (GRAD ? 2 : 1);
  double2 *glam = gpsi + DIM;
  const int count = P.count ? *P.count : P.B;

  auto csync = [&]() {
    if constexpr (CS > 1) cg::this_cluster().sync();
    else __syncthreads();
  };
  // frame of the pass in flight -> this thread's tile addressing: element idx = tid + k NT of the tile sits at
  // physical index abase ^ kx[k] ^ P0(tau)
  uint32_t abase = 0u, kx[EPT];
  auto frame_regs = [&]() {
    uint32_t hi = 0u;
#pragma unroll
    for (int i = CLOW; i < LT; ++i) hi ^= ((tid >> i) & 1) ? col[i] : 0u;
    abase = ((uint32_t)tid & LOWMASK) | hi;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      uint32_t x = 0u;
#pragma unroll
      for (int i = 0; i < TB - LT; ++i) x ^= ((k >> i) & 1) ? col[LT + i] : 0u;
      kx[k] = x;
    }
  };
  auto tile_origin = [&](uint32_t tau) {
    uint32_t p0 = 0u;
    for (int i = 0; i < n - TB; ++i) p0 ^= ((tau >> i) & 1u) ? col[TB + i] : 0u;
    return p0;
  };
  // translate the cnt ops of the pass that starts at tape index j0 to pass coordinates, stage their matrices
  auto translate = [&](int j0, int cnt) {
    if (tid < cnt) {
      const uint32_t *o = ops + (size_t)(j0 + tid) * QAS_OP_WORDS;
      uint32_t x = o[0] & ~LOWMASK, mp = o[0] & LOWMASK;
      for (int p = n - 1; p >= CLOW; --p)
        if ((x >> p) & 1u) { const int ci = piv[p]; x ^= col[ci]; mp ^= 1u << ci; }
      uint32_t rp = 0u, rcp = 0u;
      for (int i = 0; i < n; ++i) {
        rp |= (uint32_t)(__popc(col[i] & o[1]) & 1) << i;
        rcp |= (uint32_t)(__popc(col[i] & o[2]) & 1) << i;
      }
      top[tid] = make_uint4(mp, rp, rcp, o[4]);
    }
    for (int w = tid; w < cnt * 4; w += NT) Us[w] = P.U[ops[(size_t)(j0 + (w >> 2)) * QAS_OP_WORDS + 3]].u[w & 3];
  };

  // coset bases of the pass's groups (needs every translated op: call after a barrier), one thread per group
  auto group_bases = [&](int cnt) {
    if (tid < cnt) {
      const int g = (int)QAS_META_GFIRST(top[tid].w);
      if (g >= 1 && qas_meta_kind(top[tid].w) == QAS_K_U1) {
        uint32_t mm[2], rr[2];
        for (int q = 0; q < g && q < 2; ++q) { mm[q] = top[tid + q].x; rr[q] = top[tid + q].y & ((1u << TB) - 1u); }
        qas_tile_group_basis(g < 2 ? g : 2, mm, rr, TB, gkb + (size_t)tid * QAS_T_KBW);
      }
    }
  };

  for (;;) {
    // ---- work queue: rank 0 draws a circuit for the cluster and posts the slot into every CTA's control block
    if (rank == 0 && tid == 0) {
      const int q = atomicAdd(P.work_counter, 1);
      if constexpr (CS > 1) {
        auto cl = cg::this_cluster();
        for (unsigned r = 0; r < CS; ++r) cl.map_shared_rank(&ctl->q, r)[0] = q;
      } else {
        ctl->q = q;
      }
    }
    csync();
    const int qslot = ctl->q;
    if (qslot >= count) break;
    const bool clk = P.phase_clk && rank == 0 && tid == 0;      // diagnostic: per-phase clocks of the cluster's first CTA
    long long clk_prev = clk ? clock64() : 0;
    auto tick = [&](int slot) {
      if (clk) { const long long c_ = clock64(); atomicAdd(P.phase_clk + slot, (unsigned long long)(c_ - clk_prev)); clk_prev = c_; }
    };
    const int b = P.order ? P.order[qslot] : qslot;
    const uint32_t *grec = P.tape + (size_t)b * P.rec_stride;
    const int nops = (int)grec[0], status = (int)grec[1];
    const uint32_t init_index = grec[2];
    for (int w = tid; w < nops * QAS_OP_WORDS; w += NT) ops[w] = grec[QAS_REC_HDR + w];
    __syncthreads();
    tick(0);

    // ---- forward sweep: passes planned from the end of the tape; the first one creates the start state
    {
      const double amp0 = (P.init_kind == QAS_INIT_PLUS) ? exp2(-0.5 * (double)n) : 0.0;
      int e = nops - 1;
      bool first = true;
      while (first || e >= 0) {
        if (tid == 0) ctl->cnt = qas_plan_tile_pass(ops, nops, e, -1, n, TB, CLOW, QAS_T_MAXP, col, piv);
        __syncthreads();
        const int cnt = ctl->cnt, j0 = e - cnt + 1;
        translate(j0, cnt);
        frame_regs();
        __syncthreads();
        group_bases(cnt);
        __syncthreads();
        for (uint32_t tau = rank; tau < ntiles; tau += CS) {
          const uint32_t p0 = tile_origin(tau);
          if (first) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
              const uint32_t ga = abase ^ kx[k] ^ p0;
              tile[tid + k * NT] = make_double2(P.init_kind == QAS_INIT_PLUS ? amp0 : (ga == init_index ? 1.0 : 0.0), 0.0);
            }
          } else {
            double2 v[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) v[k] = __ldcg(gpsi + (abase ^ kx[k] ^ p0));
#pragma unroll
            for (int k = 0; k < EPT; ++k) tile[tid + k * NT] = v[k];
          }
          __syncthreads();
          for (int i = cnt - 1; i >= 0;) {
            const int g = max(1, (int)QAS_META_GLAST(top[i].w));
            const int f = i - g + 1;
            if (qas_meta_kind(top[f].w) == QAS_K_U1) {
              if (g == 2) t_fwd_group<2, TB, NT>(tile, top + f, gkb + (size_t)f * QAS_T_KBW, Us + 4 * f, tau, tid);
              else t_fwd_group<1, TB, NT>(tile, top + f, gkb + (size_t)f * QAS_T_KBW, Us + 4 * f, tau, tid);
            } else {
              const double2 d0 = Us[4 * f], d1 = Us[4 * f + 3];
              const uint32_t rl = top[f].y & ((1u << TB) - 1u);
              const bool fl = __popc(tau & (top[f].y >> TB)) & 1;
#pragma unroll
              for (int k = 0; k < EPT; ++k) {
                const uint32_t t = tid + k * NT;
                tile[t] = cmul((((__popc(t & rl) & 1) != 0) != fl) ? d1 : d0, tile[t]);
              }
            }
            i = f - 1;
            __syncthreads();
          }
#pragma unroll
          for (int k = 0; k < EPT; ++k) __stcg(gpsi + (abase ^ kx[k] ^ p0), tile[tid + k * NT]);
        }
        e -= cnt;
        first = false;
        csync();          // the slab is complete; the frame may be re-planned
        if (clk) atomicAdd(P.phase_clk + 4, 1ull);
      }
    }
    tick(1);

    // ---- lambda = H psi, E = Re <psi|lambda>
    {
      double e_part = t_hphase<2, 2, GRAD>(P, gpsi, glam, sdesc, rank * NT + tid, CS * NT, n);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) e_part += shfl_xor_d(e_part, off);
      if (lane == 0) ered[warp] = e_part;
      __syncthreads();
      if (tid == 0) {
        double e = 0.0;
        for (int w = 0; w < NW; ++w) e += ered[w];
        ctl->epart = e;
      }
      csync();            // lambda complete on the slab; the CTAs' energy partials are posted
      if (rank == 0 && tid == 0) {
        double e = 0.0;
        if constexpr (CS > 1) {
          auto cl = cg::this_cluster();
          for (unsigned r = 0; r < CS; ++r) e += cl.map_shared_rank(&ctl->epart, r)[0];
        } else {
          e = ctl->epart;
        }
        P.energy[b] = status ? __longlong_as_double(0x7ff8000000000000ll) : e;
      }
    }
    tick(2);

    // ---- adjoint sweep: passes planned from the start of the tape (reverse time order)
    if constexpr (GRAD) {
      int j = 0, pi = 0;
      while (j < nops) {
        if (tid == 0) ctl->cnt = qas_plan_tile_pass(ops, nops, j, +1, n, TB, CLOW, QAS_T_MAXP, col, piv);
        __syncthreads();
        const int cnt = ctl->cnt;
        translate(j, cnt);
        frame_regs();
        for (int w = tid; w < cnt * NW * 8; w += NT) Mw[w] = 0.0;
        __syncthreads();
        group_bases(cnt);
        __syncthreads();
        for (uint32_t tau = rank; tau < ntiles; tau += CS) {
          const uint32_t p0 = tile_origin(tau);
          {
            double2 v[EPT], lv[EPT];
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
              v[k] = __ldcg(gpsi + (abase ^ kx[k] ^ p0));
              lv[k] = __ldcg(glam + (abase ^ kx[k] ^ p0));
            }
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
              tile[tid + k * NT] = v[k];
              ltile[tid + k * NT] = lv[k];
            }
          }
          __syncthreads();
          for (int f = 0; f < cnt;) {
            const int g = max(1, (int)QAS_META_GFIRST(top[f].w));
            double *mw = Mw + ((size_t)f * NW + warp) * 8;
            if (qas_meta_kind(top[f].w) == QAS_K_U1) {
              if (g == 2) t_adj_group<2, TB, NT>(tile, ltile, top + f, gkb + (size_t)f * QAS_T_KBW, Us + 4 * f, tau, tid, lane, mw, NW * 8);
              else t_adj_group<1, TB, NT>(tile, ltile, top + f, gkb + (size_t)f * QAS_T_KBW, Us + 4 * f, tau, tid, lane, mw, NW * 8);
            } else {
              const double2 d0 = Us[4 * f], d1 = Us[4 * f + 3];
              const uint32_t rl = top[f].y & ((1u << TB) - 1u);
              const bool fl = __popc(tau & (top[f].y >> TB)) & 1;
              const int npar = qas_meta_npar(top[f].w);
              double acc[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) acc[q] = 0.0;
#pragma unroll
              for (int k = 0; k < EPT; ++k) {
                const uint32_t t = tid + k * NT;
                const bool hi = ((__popc(t & rl) & 1) != 0) != fl;
                const double2 dd = hi ? d1 : d0;
                const double2 q = cmul_conj(dd, tile[t]);
                const double2 lv = ltile[t];
                tile[t] = q;
                if (npar) {
                  const double re = fma(lv.x, q.x, lv.y * q.y), im = fma(lv.x, q.y, -lv.y * q.x);
                  if (hi) { acc[6] += re; acc[7] += im; } else { acc[0] += re; acc[1] += im; }
                }
                ltile[t] = cmul_conj(dd, lv);
              }
              __syncwarp();
              if (npar) reduce_store_m<32, true>(acc, lane, 0xffffffffu, mw, true);
            }
            f += g;
            __syncthreads();
          }
          const bool last = j + cnt >= nops;      // nothing reads the slab after the last adjoint pass
          if (!last) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
              __stcg(gpsi + (abase ^ kx[k] ^ p0), tile[tid + k * NT]);
              __stcg(glam + (abase ^ kx[k] ^ p0), ltile[tid + k * NT]);
            }
          }
        }
        // cross matrices of the pass: warps in fixed order -> CTA sum, CTAs in fixed rank order -> global
        double *ms = Ms + (size_t)(pi & 1) * QAS_T_MAXP * 8;
        for (int w = tid; w < cnt * 8; w += NT) {
          const int op = w >> 3, v = w & 7;
          double s = 0.0;
          for (int wp = 0; wp < NW; ++wp) s += Mw[((size_t)op * NW + wp) * 8 + v];
          ms[w] = s;
        }
        csync();
        if (rank == 0) {
          for (int w = tid; w < cnt * 8; w += NT) {
            const int op = w >> 3;
            if (!qas_meta_npar(top[op].w)) continue;
            double s = 0.0;
            if constexpr (CS > 1) {
              auto cl = cg::this_cluster();
              for (unsigned r = 0; r < CS; ++r) s += cl.map_shared_rank(ms, r)[w];
            } else {
              s = ms[w];
            }
            P.mout[((size_t)b * P.ops_cap + j + op) * 8 + (w & 7)] = s;
          }
        }
        __syncthreads();        // top / Mw are rewritten by the next pass
        j += cnt;
        ++pi;
        if (clk) atomicAdd(P.phase_clk + 5, 1ull);
      }
    }
    tick(3);
    if (clk) atomicAdd(P.phase_clk + 6, 1ull);
    // (the queue draw's cluster barrier separates this circuit's last DSMEM reads from the next circuit)
  }
}
