// kernels_c.cuh -- support-compressed evaluation path (sm_100a): c-tape compile kernel, the compressed
// evaluation kernel and the fused gradient finalize + batch reduction.
//
// For all-zero start states a circuit with support span D (dim d) is evaluated as a d-qubit register in
// the time-adapted basis of D (tape.h, "support-compressed tapes"): the support after j ops is the
// prefix t < 2^{d_j} of a 2^d-amplitude register, so
//   * a LiH-10q circuit (d = 3..10, mean 7) owns 2^d instead of 2^10 amplitudes of psi and lambda,
//     one warp evaluates it without team barriers, and an SM holds 16-20 of them instead of 6;
//   * passes need no planner descriptors: the cosets of a group are enumerated by inserting zeros at the
//     pivots of its masks inside the prefix and projected onto the common kernel of the parity masks
//     with two popcounts;
//   * lambda = H psi reads the sign tables at the physical index i(t) = xor_k t_k B_k, only for the X masks
//     inside D (the others connect the support to exact zeros).
// Circuits are binned by d (one launch per bin, each with its own work queue, launched on forked streams
// so that a small bin fills the tail of a large one); circuits with d == n keep the classic path
// (kernels.cuh), fed from the same binning.
// What the reference does instead: the dense 2^n state of PennyLane's default.qubit for every circuit
// (qas/qml_models/hamiltonian_model.py:43-91).
#pragma once
#include "kernels.cuh"

#define QAS_MAX_BINS 16

struct CompileCParams {
  const int32_t *k;
  const qas_pool_entry *pool;
  const uint32_t *sxflat;      // [S] X masks of the flat sign-table rows (real slices first)
  uint32_t *tape;              // [B][rec_stride]
  int *bin_count;              // [QAS_MAX_BINS]
  int *bin_order;              // [QAS_MAX_BINS][B]
  int B, p, c, n, ops_cap, S, gmax, dmin, max_d, full_bin;
  size_t rec_stride;
  // gate table of this call, built by `gate_blocks` extra CTAs at the end of the grid (0: a separate launch did it):
  // the table only has to be complete before the evaluation kernels start, i.e. when this kernel ends
  const double *params;
  GateTab *U;
  GateDTab *dU;
  int l, gate_blocks;
};

// per-thread shared-memory words of the c-tape compile kernel: col[n] row[n] | packed ops [cap * 3]  (odd stride)
__host__ __device__ __forceinline__ size_t qas_ccompile_smem_words(int n, int ops_cap, int p) {
  (void)p;
  return ((size_t)(2 * n) + (size_t)ops_cap * QAS_PK_WORDS) | 1;
}

// One thread per circuit: CNOT-folded tape (tape.h) -> record (5-word ops) + packed copy in shared memory, group
// marks, c-tape translation with register-resident tables (qas_compress_packed), bin assignment.
// Record: [nops, status, d (0x80000000 for a classic record), ngroups | ops | cinfo = nsl, B_0..B_{n-1}, list[nsl]].
// The frame (col / row, indexed by wire at run time) and the packed ops are the only per-thread shared memory;
// structure lists and pool entries come straight from global memory (L1-resident).
template <int CTHREADS, int NQ>
__global__ void __launch_bounds__(CTHREADS) compile_ctapes_kernel(const CompileCParams P) {
  extern __shared__ uint32_t csm[];
  if (blockIdx.x >= gridDim.x - P.gate_blocks) {        // spare CTAs: U and dU/dtheta of every (depth, key)
    qas_gate_table_entry(P.pool, P.c, P.p, P.l, P.params, P.U, P.dU, (int)(blockIdx.x - (gridDim.x - P.gate_blocks)) * CTHREADS + threadIdx.x);
    return;
  }
  const int b = blockIdx.x * CTHREADS + threadIdx.x;
  const bool live = b < P.B;
  const int n = P.n;
  const size_t stride = qas_ccompile_smem_words(n, P.ops_cap, P.p);
  uint32_t *mine = csm + (size_t)threadIdx.x * stride;
  uint32_t *col = mine, *row = col + n, *pk = row + n;
  // the pool table is read at every depth step of every thread: one shared copy per CTA
  qas_pool_entry *spool = reinterpret_cast<qas_pool_entry *>(csm + (size_t)CTHREADS * stride);
  for (int x = threadIdx.x; x < P.c; x += CTHREADS) spool[x] = P.pool[x];
  __syncthreads();
  int bin = -1;
  if (live) {
    uint32_t *rec = P.tape + (size_t)b * P.rec_stride;
    uint32_t *gops = rec + QAS_REC_HDR;
    uint32_t *gci = gops + (size_t)P.ops_cap * QAS_OP_WORDS;
    QasPk3Sink sink{pk, P.ops_cap, 0, 0};
    const int32_t *kb = P.k + (size_t)b * P.p;
    const int st = qas_compile_circuit(kb, P.p, spool, P.c, n, col, row, sink);
    int nops = st ? 0 : sink.n;
    int ngroups = qas_mark_groups_packed(pk, nops, P.gmax);
    const int nops_raw = nops, ngroups_raw = ngroups;
    const int d = qas_compress_packed<NQ>(pk, gops, &nops, &ngroups, n, P.sxflat, P.S, gci, gci + 1 + n, P.max_d);
    if (d < 0) {       // classic path: the grouped original tape (the translation above stopped half-way: compile again)
      QasPk3Sink again{pk, P.ops_cap, 0, 0};
      qas_compile_circuit(kb, P.p, spool, P.c, n, col, row, again);
      qas_mark_groups_packed(pk, nops_raw, P.gmax);
      qas_unpack_ops(pk, nops_raw, gops);
      nops = nops_raw;
      ngroups = ngroups_raw;
    }
    rec[0] = (uint32_t)nops;
    rec[1] = (uint32_t)st;
    rec[2] = d >= 0 ? (uint32_t)d : 0x80000000u;
    rec[3] = (uint32_t)ngroups;
    bin = d >= 0 ? max(d, P.dmin) - P.dmin : P.full_bin;
  }
  // bin assignment, one atomic per (warp, bin)
  const uint32_t act = __ballot_sync(0xffffffffu, live);
  if (live) {
    const uint32_t peers = __match_any_sync(act, bin);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(P.bin_count + bin, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    P.bin_order[(size_t)bin * P.B + base + __popc(peers & ((1u << lane) - 1u))] = b;
  }
}

// ------------------------------------------------------------------------------ compressed evaluation
struct EvalCParams {
  const uint32_t *tape;        // [B][rec_stride]
  const GateTab *U;
  const double *vflat;         // [S][2^n] flat sign tables, row s = slice s (real slices first)
  const uint4 *sinfo;          // [S] per slice: x = rank r of the span of its Z masks (255: look the sign up in vflat),
                               //     y = zeta_0 | zeta_1 << 16, z = zeta_2 | zeta_3 << 16 (basis of that span)
  const double *slut;          // [S][16] V_s as a function of the r parities parity(i & zeta_q) (x_s folded in)
  const int *order;            // this bin's circuit indices
  const int *count;            // ... and how many
  int *work_counter;           // this bin's queue head (zeroed before the launch)
  double *energy;              // [B]
  const GateDTab *dU;          // [NCONST + p*c] dU/dtheta table (gradient launches)
  double *grad;                // [B][p][l] compact gradients, zero-filled beforehand (gradient launches)
  int p, l;
  unsigned long long *phase_clk;   // diagnostic: summed per-team clock64 deltas [prologue, forward, H, adjoint] at +4*.. or null
  size_t rec_stride;
  int n, ops_cap, S, S_real, slot_bits;
};

__host__ __device__ inline size_t evalc_team_smem_bytes(int slot_bits, int T, bool grad, int ops_cap, int S) {
  const int wpt = T >= 32 ? T / 32 : 1;
  size_t b = ((size_t)16 << slot_bits) * (grad ? 2 : 1);
  b += (((size_t)(QAS_REC_HDR + ops_cap * QAS_OP_WORDS) * 4) + 15) & ~(size_t)15;
  b += 32 * 4;                                         // nsl + basis (<= QAS_CT_MAX_D words) | physical offsets of the HA strides
  b += (size_t)S * 16;                                 // slice records (CSlice)
  b += (size_t)ops_cap * 64;                           // staged gate matrices; gradient launches overwrite the slot of an op with its
                                                       // cross matrix M once the adjoint sweep has passed the op (last use of the gate)
  if (grad && wpt > 1) b += (size_t)2 * 2 * wpt * 64;  // cross-matrix ring (two buffers, <= 2 gates per pass)
  b += (((size_t)wpt * 8) + 15) & ~(size_t)15;
  b += 16;                                             // the team's work-queue slot
  return b;
}

// coset geometry of a pass of G gates inside the support prefix: pivots of the echelonised masks
template <int G>
struct CGeom {
  uint32_t m[G], r[G], xm[1 << G];
  int qlo, qhi;
  __device__ __forceinline__ void load(const uint32_t *o) {
#pragma unroll
    for (int q = 0; q < G; ++q) { m[q] = o[q * QAS_OP_WORDS]; r[q] = o[q * QAS_OP_WORDS + 1]; }
    const int q0 = 31 - __clz(m[0]);
    qlo = qhi = q0;
    if constexpr (G == 2) {
      const uint32_t e1 = ((m[1] >> q0) & 1u) ? (m[1] ^ m[0]) : m[1];
      const int q1 = 31 - __clz(e1);
      qlo = min(q0, q1);
      qhi = max(q0, q1);
    }
    xm[0] = 0u;
#pragma unroll
    for (int c = 1; c < (1 << G); ++c) {
      const int q = c >= 2 ? 1 : 0;
      xm[c] = xm[c & ~(1 << q)] ^ m[q];
    }
  }
  // representative of coset c, projected onto the common kernel of the r's (slot bits = logical bits)
  __device__ __forceinline__ uint32_t rep(uint32_t c) const {
    uint32_t t = insert_zero(c, qlo);
    if constexpr (G == 2) t = insert_zero(t, qhi);
    uint32_t P = t;
#pragma unroll
    for (int q = 0; q < G; ++q) P ^= (__popc(t & r[q]) & 1) ? m[q] : 0u;
    return P;
  }
};

template <int G, int T>
__device__ __forceinline__ void c_fwd_group(double2 *psi, const uint32_t *o, int dj, const double2 *Us, int tid) {
  CGeom<G> gg;
  gg.load(o);
  double2 u[G][4];
#pragma unroll
  for (int q = 0; q < G; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) u[q][e] = Us[4 * q + e];
  const uint32_t rc = (G == 1) ? o[2] : 0u;
  const uint32_t ncos = 1u << (dj - G);
  for (uint32_t c = tid; c < ncos; c += T) {
    const uint32_t P = gg.rep(c);
    if (G == 1 && rc && !(__popc(P & rc) & 1)) continue;
    double2 a[1 << G];
#pragma unroll
    for (int s = 0; s < (1 << G); ++s) a[s] = psi[P ^ gg.xm[s]];
#pragma unroll
    for (int q = G - 1; q >= 0; --q) {       // time order = descending tape index
#pragma unroll
      for (int s = 0; s < (1 << G); ++s)
        if (!((s >> q) & 1)) apply_u(a[s], a[s | (1 << q)], u[q]);
    }
#pragma unroll
    for (int s = 0; s < (1 << G); ++s) psi[P ^ gg.xm[s]] = a[s];
  }
}

// adjoint pass of a group (tape order): psi <- U^ psi ; M += conj(lam) psi ; lam <- U^ lam over the cosets
// inside the support AFTER the group.  Every lane of the team runs the same trip count (the cross-matrix
// reduction shuffles), lanes beyond the support contribute zeros.
template <int G, int T, int TW, int WPT>
__device__ __forceinline__ void c_adj_group(double2 *psi, double2 *lam, const uint32_t *o, int dj, const double2 *Us,
                                            int tid, int lane, int warp_in_team, uint32_t tmask, double *ring,
                                            double *mout_op) {
  CGeom<G> gg;
  gg.load(o);
  int npar[G];
#pragma unroll
  for (int q = 0; q < G; ++q) npar[q] = qas_meta_npar(o[q * QAS_OP_WORDS + 4]);
  double M[G][8];
#pragma unroll
  for (int q = 0; q < G; ++q)
#pragma unroll
    for (int e = 0; e < 8; ++e) M[q][e] = 0.0;
  const uint32_t rc = (G == 1) ? o[2] : 0u;
  const uint32_t ncos = 1u << (dj - G);
  for (uint32_t c0 = 0; c0 < ncos; c0 += T) {
    const uint32_t c = c0 + tid;
    if (c >= ncos) continue;
    const uint32_t P = gg.rep(c);
    if (G == 1 && rc && !(__popc(P & rc) & 1)) continue;
    double2 a[1 << G], l[1 << G];
#pragma unroll
    for (int s = 0; s < (1 << G); ++s) {
      const uint32_t x = P ^ gg.xm[s];
      a[s] = psi[x];
      l[s] = lam[x];
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
      const uint32_t x = P ^ gg.xm[s];
      psi[x] = a[s];
      lam[x] = l[s];
    }
  }
  __syncwarp(tmask);
#pragma unroll
  for (int q = 0; q < G; ++q)
    if (npar[q]) reduce_store_m<TW, false>(M[q], lane, tmask, WPT == 1 ? mout_op + (size_t)q * 8
                                                                        : ring + ((size_t)q * WPT + warp_in_team) * 8, false);
}

// Per-circuit slice record of the compressed H phase, built once per circuit by one lane per slice:
//   x = x' (compressed X mask, < 2^12) | CS_FLAT | CS_IMAG | slice index << 16
//   y, z = the (<= 4) basis masks zeta_q of the slice's Z span pulled back to compressed coordinates:
//          parity(i(t) & zeta_q) = parity(t & zc_q) with bit k of zc_q = parity(B_k & zeta_q)   (i(t) = xor_k t_k B_k)
//   w = eight nibbles: the sign index of the stride offsets a * T, a < 8  (the index is linear in t)
#define CS_FLAT 0x4000u
#define CS_IMAG 0x8000u
#define CS_PHA 16          // chead[CS_PHA + a] = physical index of the stride offset a * T

template <int T>
__device__ __forceinline__ void c_prepare_slices(const EvalCParams &P, const uint32_t *gci, uint4 *sl, uint32_t *chead, int nsl,
                                                 int d, int tid, uint32_t ent0, uint32_t hb, uint32_t tmask, int lane0) {
  const int n = P.n;
  // (warp-uniform trip count: the basis vectors travel by shuffle from the lanes that loaded them)
  for (int s0 = 0; s0 < nsl; s0 += T) {
    const int s = s0 + tid;
    const bool on = s < nsl;
    const uint32_t ent = s0 == 0 ? ent0 : (on ? gci[1 + n + s] : 0u);
    const uint32_t si = on ? ent >> 16 : 0u;
    const uint4 info = __ldg(P.sinfo + si);
    uint4 r = make_uint4((ent & 0xfffu) | (si << 16) | ((int)si >= P.S_real ? CS_IMAG : 0u), 0u, 0u, 0u);
    const bool small = info.x <= 4u;
    const uint32_t z0 = info.y & 0xffffu, z1 = info.y >> 16, z2 = info.z & 0xffffu, z3 = info.z >> 16;
    uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
    for (int k = 0; k < d; ++k) {
      const uint32_t bk = __shfl_sync(tmask, hb, lane0 + 1 + k);
      c0 |= (uint32_t)(__popc(bk & z0) & 1) << k;
      c1 |= (uint32_t)(__popc(bk & z1) & 1) << k;
      c2 |= (uint32_t)(__popc(bk & z2) & 1) << k;
      c3 |= (uint32_t)(__popc(bk & z3) & 1) << k;
    }
    // nibble a of D = sign index of a * T = xor of e_j over the set bits j of a, e_j = index of the single bit T << j
    constexpr int LT = (T == 8) ? 3 : (T == 16) ? 4 : (T == 32) ? 5 : (T == 64) ? 6 : (T == 128) ? 7 : 8;
    uint32_t e[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      e[j] = ((c0 >> (LT + j)) & 1u) | (((c1 >> (LT + j)) & 1u) << 1) | (((c2 >> (LT + j)) & 1u) << 2) | (((c3 >> (LT + j)) & 1u) << 3);
    const uint32_t D = ((e[0] * 0x11111111u) & 0xf0f0f0f0u) ^ ((e[1] * 0x11111111u) & 0xff00ff00u) ^ ((e[2] * 0x11111111u) & 0xffff0000u);
    if (small) {
      r.y = c0 | (c1 << 16);
      r.z = c2 | (c3 << 16);
      r.w = D;
    } else {
      r.x |= CS_FLAT;
    }
    if (on) sl[s] = r;
  }
  {
    const uint32_t x = (uint32_t)((tid & 7) * T);
    uint32_t ph = 0u;
    for (int k = 0; k < d; ++k) {
      const uint32_t bk = __shfl_sync(tmask, hb, lane0 + 1 + k);
      ph ^= ((x >> k) & 1u) ? bk : 0u;
    }
    if (tid < 8) chead[CS_PHA + tid] = ph;
  }
}

// lambda = H psi on the support (compressed coordinates), returns this thread's share of Re <psi|lambda>.
// A thread takes HA amplitudes at a time (t = t0 + T a) and walks the in-span slices ONCE for all of them.  The
// sign V_s[i] at the physical index i = xor_k t_k B_k comes
//   * for slices whose Z masks span <= 4 dimensions (all the excitation slices of a chemistry Hamiltonian: 28 of
//     35 for LiH) from a 16-entry table indexed by the parities of t with the pulled-back span basis: the index of
//     t0 costs four popcounts per (slice, thread), the indices of the HA strided amplitudes are that index xor a
//     per-slice nibble (CSlice.w) -- the index is linear in t and t0, a * T have disjoint bits;
//   * for the few slices with many Z masks (the diagonal one, the number-operator dressed single excitations) from
//     the flat 2^n table at the physical index (base index xor the per-circuit stride offsets).
template <int HA, int T, bool GRAD>
__device__ __forceinline__ double c_hphase(const EvalCParams &P, const double2 *psi, double2 *lam, const uint32_t *chead,
                                           const uint4 *sl, int nsl, int d, int tid) {
  double e_part = 0.0;
  const uint32_t dim = 1u << d;
  const int n = P.n;
  for (uint32_t t0 = tid; t0 < dim; t0 += T * HA) {
    uint32_t ph0 = 0u;
    for (int kk = 0; kk < d; ++kk) ph0 ^= ((t0 >> kk) & 1u) ? chead[1 + kk] : 0u;
    double2 acc[HA];
#pragma unroll
    for (int a = 0; a < HA; ++a) acc[a] = make_double2(0.0, 0.0);
    for (int s = 0; s < nsl; ++s) {
      const uint4 r = sl[s];
      const uint32_t xp = r.x & 0xfffu, si = r.x >> 16;
      double v[HA];
      if (!(r.x & CS_FLAT)) {
        const double *lut = P.slut + (size_t)si * 16;
        const uint32_t idx0 = (__popc(t0 & (r.y & 0xffffu)) & 1) | ((__popc(t0 & (r.y >> 16)) & 1) << 1) |
                              ((__popc(t0 & (r.z & 0xffffu)) & 1) << 2) | ((__popc(t0 & (r.z >> 16)) & 1) << 3);
#pragma unroll
        for (int a = 0; a < HA; ++a) v[a] = __ldg(lut + ((idx0 ^ (r.w >> (4 * a))) & 15u));
      } else {
        const double *vrow = P.vflat + ((size_t)si << n);
#pragma unroll
        for (int a = 0; a < HA; ++a) v[a] = (t0 + (uint32_t)(T * a) < dim) ? __ldg(vrow + (ph0 ^ chead[CS_PHA + a])) : 0.0;
      }
      if (!(r.x & CS_IMAG)) {
#pragma unroll
        for (int a = 0; a < HA; ++a) {
          const uint32_t t = t0 + (uint32_t)(T * a);
          if (t < dim) {
            const double2 pa = psi[t ^ xp];
            acc[a].x = fma(v[a], pa.x, acc[a].x); acc[a].y = fma(v[a], pa.y, acc[a].y);
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < HA; ++a) {
          const uint32_t t = t0 + (uint32_t)(T * a);
          if (t < dim) {
            const double2 pa = psi[t ^ xp];
            acc[a].x = fma(-v[a], pa.y, acc[a].x); acc[a].y = fma(v[a], pa.x, acc[a].y);
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < HA; ++a) {
      const uint32_t t = t0 + (uint32_t)(T * a);
      if (t < dim) {
        const double2 ai = psi[t];
        e_part = fma(ai.x, acc[a].x, fma(ai.y, acc[a].y, e_part));
        if constexpr (GRAD) lam[t] = acc[a];
      }
    }
  }
  return e_part;
}

// AG = gates per ADJOINT pass (forward passes always follow the tape's groups).  AG = 1 un-applies the two gates of a
// group in two passes: the adjoint pass of a two-gate group holds 4 + 4 amplitudes and two cross matrices (the kernel's
// register peak); one gate at a time needs half of that, which buys more circuits in flight for more passes.
template <int T, bool GRAD, int BLK, int MINB, int AG = 2>
__global__ void __launch_bounds__(BLK, MINB) evalc_kernel(const EvalCParams P) {
  constexpr int WPT = T >= 32 ? T / 32 : 1;
  constexpr int TW = T >= 32 ? 32 : T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int team = threadIdx.x / T, tid = threadIdx.x % T;
  const int lane = threadIdx.x & 31, warp_in_team = tid >> 5;
  const uint32_t tmask = T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << ((lane / TW) * TW));
  const size_t team_bytes = evalc_team_smem_bytes(P.slot_bits, T, GRAD, P.ops_cap, P.S);
  unsigned char *base = smem_raw + (size_t)team * team_bytes;
  double2 *psi = reinterpret_cast<double2 *>(base), *lam = nullptr;
  base += (size_t)16 << P.slot_bits;
  if constexpr (GRAD) { lam = reinterpret_cast<double2 *>(base); base += (size_t)16 << P.slot_bits; }
  uint32_t *rec = reinterpret_cast<uint32_t *>(base);
  uint32_t *ops = rec + QAS_REC_HDR;
  base += (((size_t)(QAS_REC_HDR + P.ops_cap * QAS_OP_WORDS) * 4) + 15) & ~(size_t)15;
  uint32_t *chead = reinterpret_cast<uint32_t *>(base);          // nsl, B_0 .. | CS_PHA: stride offsets
  base += 32 * 4;
  uint4 *sl = reinterpret_cast<uint4 *>(base);                   // slice records of this circuit
  base += (size_t)P.S * 16;
  double2 *Us = reinterpret_cast<double2 *>(base);
  base += (size_t)P.ops_cap * 64;
  // cross matrices [ops_cap][8] (GRAD): in place of the gate matrices.  The adjoint sweep reads U_j for the last time in
  // the pass of op j and publishes M_j after that pass's sweep (all lanes of the team are past their loads: warp shuffles
  // of the reduction for one-warp teams, the team barrier before the fixed-order sum for multi-warp teams), so the
  // 64 bytes of the op's slot are free -- 16 one-warp circuits of a 2^7 register fit an SM instead of 12
  double *Ms = reinterpret_cast<double *>(Us);
  double *red = nullptr;
  if constexpr (GRAD && WPT > 1) { red = reinterpret_cast<double *>(base); base += (size_t)2 * 2 * WPT * 64; }
  double *ered = reinterpret_cast<double *>(base);
  base += (((size_t)WPT * 8) + 15) & ~(size_t)15;
  int *bslot = reinterpret_cast<int *>(base);
  const int count = *P.count;
  const int n = P.n;

  // Work queue with the next circuit's index in flight: lane 0 draws the queue slot of circuit i + 1 while circuit
  // i is being set up and reads its batch index after the forward sweep, so that neither the L2 atomic nor the
  // order look-up sits on the serial path of a team (four dependent L2 round trips per circuit before).
  int qnext = 0, bnext = -1;
  if (tid == 0) {
    qnext = atomicAdd(P.work_counter, 1);
    bnext = qnext < count ? P.order[qnext] : -1;
  }
  for (;;) {
    if (tid == 0) *bslot = bnext;
    team_sync<T>(team, tmask);
    const int b = *bslot;
    if (b < 0) break;
    if (tid == 0) qnext = atomicAdd(P.work_counter, 1);
    long long clk_prev = 0;
    if (P.phase_clk && tid == 0) clk_prev = clock64();
    // ---- 1. fetch the record, zero the support
    const uint32_t *grec = P.tape + (size_t)b * P.rec_stride;
    const uint32_t *gci = grec + QAS_REC_HDR + (size_t)P.ops_cap * QAS_OP_WORDS;
    // everything a lane needs from the record is requested before the header is consumed: one L2 round trip
    // instead of header -> body -> slice list
    const int recw = QAS_REC_HDR + P.ops_cap * QAS_OP_WORDS;
    uint32_t early[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int w = tid + j * T; early[j] = w < recw ? grec[w] : 0u; }
    const uint32_t ent0 = tid < P.S ? gci[1 + n + tid] : 0u;           // this lane's slice entry
    // nsl, B_0 .. : every warp of a multi-warp team its own copy; sub-warp teams (T = 16) index by their own lanes
    constexpr int LW = T >= 32 ? 32 : T;
    const int tl = lane % LW, lane0 = lane - tl;
    const uint32_t hb = tl <= QAS_CT_MAX_D ? gci[tl] : 0u;
    const int nops = (int)grec[0], status = (int)grec[1], d = (int)grec[2], ngroups = (int)grec[3];
    const int nsl = (int)__shfl_sync(tmask, hb, lane0);
    {
      const int used = QAS_REC_HDR + nops * QAS_OP_WORDS;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int w = tid + j * T; if (w < used) rec[w] = early[j]; }
      for (int w = tid + 4 * T; w < used; w += T) rec[w] = grec[w];
      if (tid <= d) chead[tid] = hb;
      c_prepare_slices<T>(P, gci, sl, chead, nsl, d, tid, ent0, hb, tmask, lane0);
      const uint32_t dim = 1u << d;
      for (uint32_t i = tid; i < dim; i += T) psi[i] = make_double2(i == 0u ? 1.0 : 0.0, 0.0);
    }
    team_sync<T>(team, tmask);
    for (int w = tid; w < nops * 4; w += T) Us[w] = P.U[ops[(w >> 2) * QAS_OP_WORDS + 3]].u[w & 3];
    team_sync<T>(team, tmask);

    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 0, (unsigned long long)(c_ - clk_prev)); clk_prev = c_; }
    // ---- 2. forward sweep (tape is stored last-gate-first: walk the groups from the end)
    {
      int e = nops - 1;
      for (int it = 0; it < ngroups; ++it) {
        const int g = (int)QAS_META_GLAST(ops[(size_t)e * QAS_OP_WORDS + 4]);
        const int j = e - g + 1;
        const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
        const int dj = (int)QAS_META_DJ(o[4]);
        if (qas_meta_kind(o[4]) == QAS_K_U1) {
          if (g == 2) c_fwd_group<2, T>(psi, o, dj, Us + 4 * j, tid);
          else c_fwd_group<1, T>(psi, o, dj, Us + 4 * j, tid);
        } else {
          const double2 d0 = Us[4 * j], d1 = Us[4 * j + 3];
          const uint32_t r = o[1], sup = 1u << dj;
          for (uint32_t i = tid; i < sup; i += T) psi[i] = cmul((__popc(i & r) & 1) ? d1 : d0, psi[i]);
        }
        e = j - 1;
        team_sync<T>(team, tmask);
      }
    }

    if (tid == 0) bnext = qnext < count ? P.order[qnext] : -1;
    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 1, (unsigned long long)(c_ - clk_prev)); clk_prev = c_; }
    // ---- 3. lambda = H psi on the support, E = Re <psi|lambda>
    double e_part;
    {
      const int per = (1 << d) / T;                 // amplitudes per thread
      if (per <= 2) e_part = c_hphase<2, T, GRAD>(P, psi, lam, chead, sl, nsl, d, tid);
      else if (per <= 4) e_part = c_hphase<4, T, GRAD>(P, psi, lam, chead, sl, nsl, d, tid);
      else e_part = c_hphase<8, T, GRAD>(P, psi, lam, chead, sl, nsl, d, tid);
    }
#pragma unroll
    for (int off = TW / 2; off > 0; off >>= 1) e_part += shfl_xor_d(e_part, off, tmask);
    if constexpr (WPT > 1) {
      if (lane == 0) ered[warp_in_team] = e_part;
      team_sync<T>(team, tmask);
      if (tid == 0) {
        double e = 0.0;
#pragma unroll
        for (int w = 0; w < WPT; ++w) e += ered[w];
        e_part = e;
      }
    } else if constexpr (GRAD) {
      team_sync<T>(team, tmask);
    }
    if (tid == 0) P.energy[b] = status ? __longlong_as_double(0x7ff8000000000000ll) : e_part;

    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 2, (unsigned long long)(c_ - clk_prev)); clk_prev = c_; }
    // ---- 4. adjoint sweep (tape order = reverse time order)
    if constexpr (GRAD) {
      int j = 0;
      for (int it = 0; it < ngroups; ++it) {
        const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
        const int g = (int)QAS_META_GFIRST(o[4]);
        const int dj = (int)QAS_META_DJ(o[4]);
        double *ring = WPT > 1 ? red + (size_t)(it & 1) * 2 * WPT * 8 : nullptr;
        double *mop = Ms + (size_t)j * 8;
        if (qas_meta_kind(o[4]) == QAS_K_U1) {
          if (g == 2) {
            if constexpr (AG >= 2) {
              c_adj_group<2, T, TW, WPT>(psi, lam, o, dj, Us + 4 * j, tid, lane, warp_in_team, tmask, ring, mop);
            } else {       // tape order: op j first (the support is still the one after the group), then op j + 1 on its own support
              // (multi-warp teams: the second pass publishes its per-warp partials in the ring slots of gate 1; both gates'
              // partials are summed in fixed warp order after the group, as for a two-gate pass)
              c_adj_group<1, T, TW, WPT>(psi, lam, o, dj, Us + 4 * j, tid, lane, warp_in_team, tmask, ring, mop);
              team_sync<T>(team, tmask);
              c_adj_group<1, T, TW, WPT>(psi, lam, o + QAS_OP_WORDS, (int)QAS_META_DJ(o[QAS_OP_WORDS + 4]), Us + 4 * (j + 1), tid, lane,
                                         warp_in_team, tmask, WPT > 1 ? ring + (size_t)WPT * 8 : ring, mop + 8);
            }
          } else c_adj_group<1, T, TW, WPT>(psi, lam, o, dj, Us + 4 * j, tid, lane, warp_in_team, tmask, ring, mop);
        } else {
          const double2 d0 = Us[4 * j], d1 = Us[4 * j + 3];
          const uint32_t r = o[1], sup = 1u << dj;
          const int npar = qas_meta_npar(o[4]);
          double acc[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = 0.0;
          for (uint32_t i = tid; i < sup; i += T) {
            const bool hi = __popc(i & r) & 1;
            const double2 dd = hi ? d1 : d0;
            const double2 q = cmul_conj(dd, psi[i]);
            const double2 lv = lam[i];
            psi[i] = q;
            if (npar) {
              const double re = fma(lv.x, q.x, lv.y * q.y), im = fma(lv.x, q.y, -lv.y * q.x);
              if (hi) { acc[6] += re; acc[7] += im; } else { acc[0] += re; acc[1] += im; }
            }
            lam[i] = cmul_conj(dd, lv);
          }
          __syncwarp(tmask);
          if (npar) reduce_store_m<TW, false>(acc, lane, tmask, WPT == 1 ? mop : ring + (size_t)warp_in_team * 8, false);
        }
        if constexpr (WPT > 1) {
          team_sync<T>(team, tmask);
          if (tid < 8 * g) {
            const int q = tid >> 3, v = tid & 7;
            if (qas_meta_npar(o[q * QAS_OP_WORDS + 4])) {
              double m = 0.0;
              for (int wp = 0; wp < WPT; ++wp) m += ring[((size_t)q * WPT + wp) * 8 + v];
              mop[(size_t)q * 8 + v] = m;
            }
          }
        } else {
          team_sync<T>(team, tmask);
        }
        j += g;
      }
      // dE/dtheta_j = 2 Re sum_ab (dU_j)_ab M_ab for every parametrised op: one lane per op, all the dU loads of
      // the circuit in flight together (one L2 round trip per circuit), straight into the compact gradient
      for (int oi = tid; oi < nops; oi += T) {
        const uint32_t *o = ops + (size_t)oi * QAS_OP_WORDS;
        const int npar = min(qas_meta_npar(o[4]), P.l);
        if (!npar) continue;
        const double2 *mp = reinterpret_cast<const double2 *>(Ms + (size_t)oi * 8);
        double2 m[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = mp[q];
        double *gout = P.grad + ((size_t)b * P.p + qas_meta_depth(o[4])) * P.l;
        const GateDTab *dt = P.dU + o[3];
        for (int jj = 0; jj < npar; ++jj) {
          double part = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 dq = dt->d[jj][q];
            part += dq.x * m[q].x - dq.y * m[q].y;
          }
          gout[jj] = 2.0 * part;
        }
      }
    }
    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 3, (unsigned long long)(c_ - clk_prev)); atomicAdd(P.phase_clk + 4, 1ull); }
    team_sync<T>(team, tmask);   // state buffers are reused by the next circuit of this team
  }
}

