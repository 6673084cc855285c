// kernels.cuh -- device code of the batched candidate-circuit evaluator (sm_100a).
//
// One TEAM of T threads evaluates one circuit with the whole complex128 statevector (and,
// for gradients, the adjoint vector lambda = H psi) resident in shared memory; a CTA of 256
// threads hosts 256/T teams (sub-warp teams for <= 5 qubits, multi-warp teams with named
// barriers above 6).  Per circuit:
//   1. prologue: tape compiler (tape.h) lowers k -> generalised-gate tape in shared memory
//   2. forward sweep over the tape (2x2 unitaries on index pairs (P, P^m))
//   3. lambda[i] = sum_g V_g[i] psi[i ^ x_g]   (Hamiltonian grouped by X-mask; V_g are sign
//      tables precomputed once per context), E = Re <psi|lambda>
//   4. adjoint sweep: per gate  psi <- U^ psi ; M_ab += conj(lambda_a) psi_b ; lambda <- U^ lambda,
//      dE/dtheta_j = 2 Re sum_ab (dU/dtheta_j)_ab M_ab   -- one 2x2 reduction yields all (<=3)
//      parameter derivatives of the gate.
// What the reference does instead: one PennyLane QNode execution per circuit plus autograd /
// parameter-shift / lightning adjoint (qas/qml_models/hamiltonian_model.py:43-91).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "tape.h"

struct __align__(16) GateTab { double2 u[4]; };        // U00 U01 U10 U11
struct __align__(16) GateDTab { double2 d[3][4]; };    // dU/dtheta_j, same layout

struct WComb { uint32_t w[8]; };

struct EvalParams {
  const int32_t *k;            // [B][p]
  const qas_pool_entry *pool;  // [c]
  const GateTab *U;            // [NCONST + p*c]
  const double *vtab;          // [S][2^(n-RB)][2^RB] sign tables, coset-major (see "sign tables")
  const uint32_t *cdesc;       // [C_real + C_imag] class representatives rho as (swizzled) byte offsets
  const uint32_t *cdesc_u;     // the same representatives in quotient coordinates, qas_ubar(rho)
  int hs_min;                  // support-restricted H application when dbar + hs_min <= n - RB (0: never)
  int C_real, C_imag;          // X-mask classes with real / imaginary sign tables (real first)
  int rb0, rb1, rb2;           // ascending pivots of the register subspace W of the H-apply blocking
  WComb wc16;                  // byte offset of wcomb[j] = xor of the basis vectors of W selected by j
  double *energy;              // [B]
  double2 *gstate;             // global-memory statevector scratch (path 1)
  double *mout;                // [B][ops_cap][8] cross matrices M of the parametrised ops (GRAD)
  const GateDTab *dU;          // [NCONST + p*c]          (thread-per-circuit kernel only)
  double *grad;                // [B][p][l] compact gradients, zero-filled (thread-per-circuit kernel only)
  unsigned long long *phase_clk;   // diagnostic: summed per-team clock64 deltas [prologue, forward, H, adjoint] or null
  int *work_counter;           // global queue head of the team kernels (zeroed before every launch)
  const uint32_t *tape;        // [B][rec_stride] compiled tape records (compile_tapes_kernel / compile_ctapes_kernel)
  size_t rec_stride;           // words per record
  const int *order;            // circuit indices of the work queue (null: 0 .. B-1)
  const int *count;            // ... and how many (null: B)
  uint64_t init_basis;
  int B, p, c, l, n, ops_cap, init_kind;
};

// ------------------------------------------------------------------------------ helpers
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {   // a*b + c
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 cfma_conj(double2 a, double2 b, double2 c) {  // conj(a)*b + c
  return make_double2(fma(a.x, b.x, fma(a.y, b.y, c.x)), fma(a.x, b.y, fma(-a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 cmul_conj(double2 a, double2 b) {            // conj(a)*b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ uint32_t insert_zero(uint32_t t, int b) {
  return ((t >> b) << (b + 1)) | (t & ((1u << b) - 1u));
}
// amplitudes per thread = 2^bits in the fused gate passes and in the H application: as many as
// leave every thread of the team one coset, at most 3 bits
__host__ __device__ constexpr int qas_block_bits(int n, int T) {
  int lt = 0;
  while ((1 << lt) < T) ++lt;
  return n - lt >= 3 ? 3 : (n - lt > 0 ? n - lt : 0);
}
__device__ __forceinline__ double shfl_xor_d(double v, int off, uint32_t mask = 0xffffffffu) {
  return __shfl_xor_sync(mask, v, off);
}

template <int T>
__device__ __forceinline__ void team_sync(int team, uint32_t tmask) {
  if constexpr (T <= 32) {
    __syncwarp(tmask);
  } else if constexpr (T == 256) {
    __syncthreads();          // 256-thread teams are always alone in their CTA
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(T) : "memory");
  }
}

// ------------------------------------------------------------------------------ gate table
// U and dU/dtheta for every (depth i, pool key) from the shared parameter super-set
// params[p][c][l]; one thread per entry.  Matrix conventions = PennyLane's (oracle/gates.py).
__device__ __forceinline__ double2 cexpi(double t) {
  double s, c;
  sincos(t, &s, &c);
  return make_double2(c, s);
}
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double2 cmuli(double2 a) { return make_double2(-a.y, a.x); }   // i*a

// entry idx = depth * c + key of the table; also called by spare CTAs of the c-tape compile kernel (kernels_c.cuh), which
// saves the separate launch in front of the compiler
__device__ __forceinline__ void qas_gate_table_entry(const qas_pool_entry *pool, int c, int p, int l, const double *params,
                                                     GateTab *U, GateDTab *dU, int idx) {
  if (idx >= p * c) return;
  const int key = idx % c;
  const int g = pool[key].gate_id;
  if (qas_gate_nparams(g) == 0) return;
  const double *th = params + (size_t)idx * l;
  GateTab u;
  GateDTab d;
  const double2 z0 = make_double2(0.0, 0.0);
#pragma unroll
  for (int a = 0; a < 4; ++a) { u.u[a] = z0; d.d[0][a] = z0; d.d[1][a] = z0; d.d[2][a] = z0; }
  switch (g) {
    case QAS_G_ROT: case QAS_G_CROT: {
      // Rot(phi,theta,omega) = RZ(omega) RY(theta) RZ(phi)
      const double phi = th[0], theta = th[1], om = th[2];
      double s, co;
      sincos(0.5 * theta, &s, &co);
      const double2 ep = cexpi(0.5 * (phi + om)), em = cexpi(0.5 * (phi - om));
      const double2 epc = make_double2(ep.x, -ep.y), emc = make_double2(em.x, -em.y);
      u.u[0] = cscale(epc, co);  u.u[1] = cscale(em, -s);
      u.u[2] = cscale(emc, s);   u.u[3] = cscale(ep, co);
      // d/dphi
      d.d[0][0] = cscale(cmuli(u.u[0]), -0.5); d.d[0][1] = cscale(cmuli(u.u[1]), 0.5);
      d.d[0][2] = cscale(cmuli(u.u[2]), -0.5); d.d[0][3] = cscale(cmuli(u.u[3]), 0.5);
      // d/dtheta
      d.d[1][0] = cscale(epc, -0.5 * s);  d.d[1][1] = cscale(em, -0.5 * co);
      d.d[1][2] = cscale(emc, 0.5 * co);  d.d[1][3] = cscale(ep, -0.5 * s);
      // d/domega
      d.d[2][0] = cscale(cmuli(u.u[0]), -0.5); d.d[2][1] = cscale(cmuli(u.u[1]), -0.5);
      d.d[2][2] = cscale(cmuli(u.u[2]), 0.5);  d.d[2][3] = cscale(cmuli(u.u[3]), 0.5);
      break;
    }
    case QAS_G_RX: case QAS_G_CRX: {
      double s, co;
      sincos(0.5 * th[0], &s, &co);
      u.u[0] = u.u[3] = make_double2(co, 0.0);
      u.u[1] = u.u[2] = make_double2(0.0, -s);
      d.d[0][0] = d.d[0][3] = make_double2(-0.5 * s, 0.0);
      d.d[0][1] = d.d[0][2] = make_double2(0.0, -0.5 * co);
      break;
    }
    case QAS_G_RY: case QAS_G_CRY: {
      double s, co;
      sincos(0.5 * th[0], &s, &co);
      u.u[0] = u.u[3] = make_double2(co, 0.0);
      u.u[1] = make_double2(-s, 0.0);
      u.u[2] = make_double2(s, 0.0);
      d.d[0][0] = d.d[0][3] = make_double2(-0.5 * s, 0.0);
      d.d[0][1] = make_double2(-0.5 * co, 0.0);
      d.d[0][2] = make_double2(0.5 * co, 0.0);
      break;
    }
    case QAS_G_RZ: case QAS_G_CRZ: case QAS_G_ISINGXX: case QAS_G_ISINGYY: case QAS_G_ISINGZZ: {
      // diag(e^{-it/2}, e^{it/2}); for the Ising gates this is the parity-diagonal core
      const double2 e = cexpi(0.5 * th[0]);
      u.u[0] = make_double2(e.x, -e.y);
      u.u[3] = e;
      d.d[0][0] = cscale(cmuli(u.u[0]), -0.5);
      d.d[0][3] = cscale(cmuli(u.u[3]), 0.5);
      break;
    }
    case QAS_G_PHASESHIFT: case QAS_G_U1: case QAS_G_CPHASE: {
      const double2 e = cexpi(th[0]);
      u.u[0] = make_double2(1.0, 0.0);
      u.u[3] = e;
      d.d[0][3] = cmuli(e);
      break;
    }
    case QAS_G_U2: {
      const double r = 0.70710678118654752440;
      const double2 ed = cexpi(th[1]), ef = cexpi(th[0]), efd = cexpi(th[0] + th[1]);
      u.u[0] = make_double2(r, 0.0);   u.u[1] = cscale(ed, -r);
      u.u[2] = cscale(ef, r);          u.u[3] = cscale(efd, r);
      d.d[0][2] = cmuli(u.u[2]);       d.d[0][3] = cmuli(u.u[3]);
      d.d[1][1] = cmuli(u.u[1]);       d.d[1][3] = cmuli(u.u[3]);
      break;
    }
    case QAS_G_U3: {
      double s, co;
      sincos(0.5 * th[0], &s, &co);
      const double2 ed = cexpi(th[2]), ef = cexpi(th[1]), efd = cexpi(th[1] + th[2]);
      u.u[0] = make_double2(co, 0.0);  u.u[1] = cscale(ed, -s);
      u.u[2] = cscale(ef, s);          u.u[3] = cscale(efd, co);
      d.d[0][0] = make_double2(-0.5 * s, 0.0);  d.d[0][1] = cscale(ed, -0.5 * co);
      d.d[0][2] = cscale(ef, 0.5 * co);         d.d[0][3] = cscale(efd, -0.5 * s);
      d.d[1][2] = cmuli(u.u[2]);                d.d[1][3] = cmuli(u.u[3]);
      d.d[2][1] = cmuli(u.u[1]);                d.d[2][3] = cmuli(u.u[3]);
      break;
    }
    default: return;
  }
  U[QAS_NCONST + idx] = u;
  dU[QAS_NCONST + idx] = d;
}

static __global__ void build_gate_table_kernel(const qas_pool_entry *pool, int c, int p, int l,
                                        const double *params, GateTab *U, GateDTab *dU) {
  qas_gate_table_entry(pool, c, p, l, params, U, dU, blockIdx.x * blockDim.x + threadIdx.x);
}

// ------------------------------------------------------------------------------ sign tables
// H is grouped by X mask ("slice"):  (H psi)[i] = sum_s (i if imag) V_s[i] psi[i ^ x_s],
// V_s[i] = sum_{t in slice s} c_t * (+-1 from i^{nY}) * (-1)^{popc((i ^ x_s) & z_t)}.
// Blocking of the H application: every thread owns the 2^RB amplitudes of one coset
// P ^ span(w_0..w_{RB-1}) of a "register subspace" W chosen per Hamiltonian on the host (basis in
// reduced row-echelon form, pivots rb0<rb1<rb2; wcomb[j] = xor of the basis vectors selected by
// the bits of j).  Slices whose X masks agree modulo W form a class: the thread loads the 2^RB
// partner amplitudes psi[P ^ rho ^ wcomb[j]] ONCE per class (rho = class representative, zero at
// the pivots) and serves every slice x = rho ^ wcomb[jw] of the class from registers (partner of
// own slot j is slot j ^ jw).  For chemistry Hamiltonians W is spanned by excitation masks, which
// collapses e.g. the 35 X masks of LiH-10q into 11 classes of 4.  Every class owns exactly 2^RB
// table rows (jw = 0..2^RB-1; a zero row where the Hamiltonian has no such mask), so the class
// loop is fully static: no per-slice descriptor decoding, partner slots are compile-time register
// names.  Rows are stored coset-major, vtab[row][u][j] = V_row[P(u) ^ wcomb[j]]: as double2 pairs,
// vtab[row][jp][u][2], so that every 128-bit load of a warp covers 512 contiguous bytes.
__host__ __device__ __forceinline__ uint32_t qas_insert_zero(uint32_t t, int b) {
  return ((t >> b) << (b + 1)) | (t & ((1u << b) - 1u));
}
__host__ __device__ __forceinline__ uint32_t qas_deposit_rest(uint32_t u, int RB, int rb0, int rb1, int rb2) {
  if (RB >= 1) u = qas_insert_zero(u, rb0);
  if (RB >= 2) u = qas_insert_zero(u, rb1);
  if (RB >= 3) u = qas_insert_zero(u, rb2);
  return u;
}

static __global__ void build_sign_tables_kernel(int n, int S, int RB, int rb0, int rb1, int rb2, WComb wc,
                                         const uint32_t *sx, const int *sbegin, const uint32_t *tz,
                                         const double *tc, double *vtab) {
  const size_t dim = (size_t)1 << n;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dim * (size_t)S) return;
  const int s = (int)(idx >> n);
  const uint32_t e = (uint32_t)(idx & (dim - 1));
  // element e of a row: [jp][u][2] with j = 2 jp + (e & 1)  (RB = 0: [u])
  const uint32_t nrest = (uint32_t)(dim >> RB);
  const uint32_t u = RB ? ((e >> 1) % nrest) : e;
  const uint32_t jj = RB ? ((((e >> 1) / nrest) << 1) | (e & 1u)) : 0u;
  const uint32_t i = qas_deposit_rest(u, RB, rb0, rb1, rb2) ^ wc.w[jj];
  const uint32_t j = i ^ sx[s];
  double v = 0.0;
  for (int t = sbegin[s]; t < sbegin[s + 1]; ++t) v += (__popc(j & tz[t]) & 1) ? -tc[t] : tc[t];
  vtab[idx] = v;
}

// acc[j] += (i if imag) V[j] * a[j ^ XH]   for the 2^RB amplitudes a thread owns
template <int NB, int XH>
__device__ __forceinline__ void hacc(double2 (&acc)[NB], const double2 (&a)[NB], const double (&v)[NB],
                                     bool imag) {
  if (!imag) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      acc[j].x = fma(v[j], a[(j ^ XH) & (NB - 1)].x, acc[j].x);
      acc[j].y = fma(v[j], a[(j ^ XH) & (NB - 1)].y, acc[j].y);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      acc[j].x = fma(-v[j], a[(j ^ XH) & (NB - 1)].y, acc[j].x);
      acc[j].y = fma(v[j], a[(j ^ XH) & (NB - 1)].x, acc[j].y);
    }
  }
}
// row jw of a class: partner of own slot j is slot j ^ jw (jw is a compile-time value after unrolling)
template <int NB>
__device__ __forceinline__ void hacc_row(double2 (&acc)[NB], const double2 (&a)[NB], const double (&v)[NB],
                                         int jw, bool imag) {
  if constexpr (NB == 1) { hacc<1, 0>(acc, a, v, imag); }
  else if constexpr (NB == 2) { if (jw == 0) hacc<2, 0>(acc, a, v, imag); else hacc<2, 1>(acc, a, v, imag); }
  else if constexpr (NB == 4) {
    switch (jw) { case 0: hacc<4, 0>(acc, a, v, imag); break; case 1: hacc<4, 1>(acc, a, v, imag); break;
                  case 2: hacc<4, 2>(acc, a, v, imag); break; default: hacc<4, 3>(acc, a, v, imag); break; }
  } else {
    switch (jw) { case 0: hacc<8, 0>(acc, a, v, imag); break; case 1: hacc<8, 1>(acc, a, v, imag); break;
                  case 2: hacc<8, 2>(acc, a, v, imag); break; case 3: hacc<8, 3>(acc, a, v, imag); break;
                  case 4: hacc<8, 4>(acc, a, v, imag); break; case 5: hacc<8, 5>(acc, a, v, imag); break;
                  case 6: hacc<8, 6>(acc, a, v, imag); break; default: hacc<8, 7>(acc, a, v, imag); break; }
  }
}
// p -> this thread's first element of the row; pair jp of the row sits nrest double2 further on
template <int NB>
__device__ __forceinline__ void load_vrow(double (&v)[NB], const double *p, uint32_t nrest) {
  if constexpr (NB >= 2) {
#pragma unroll
    for (int j = 0; j < NB; j += 2) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(p) + (size_t)(j >> 1) * nrest);
      v[j] = t.x;
      v[j + 1] = t.y;
    }
  } else {
    v[0] = __ldg(p);
  }
}

// ------------------------------------------------------------------------------ tape compile
// One thread per circuit: runs the tape compiler and the pass planner (tape.h) and writes a record
//   [nops, status, init_index, ngroups | ops[ops_cap][5] | gd[ops_cap][2 + n] | hd[QAS_HD_WORDS]]   (uint32)
// to global memory.  Doing this in its own launch (B-way parallel) keeps the serial GF(2) work
// out of the evaluation teams' critical path; the records stay L2-resident.
#define QAS_REC_HDR 4
__host__ __device__ __forceinline__ size_t qas_rec_words(int ops_cap, int n) {
  return (size_t)QAS_REC_HDR + (size_t)ops_cap * QAS_OP_WORDS + (size_t)ops_cap * QAS_GD_STRIDE(n) + QAS_HD_WORDS;
}
__host__ __device__ __forceinline__ size_t qas_rec_hd_offset(int ops_cap, int n) {
  return (size_t)QAS_REC_HDR + (size_t)ops_cap * QAS_OP_WORDS + (size_t)ops_cap * QAS_GD_STRIDE(n);
}
// ------------------------------------------------------------------------------ evaluation
// Shared-memory statevectors are stored swizzled (qas_swz, tape.h): a pass whose cosets differ
// only in high index bits would otherwise put the lanes of a quarter-warp on the same 16-byte bank
// group.  The planner hands the gate passes slot coordinates, so only the dense loops evaluate it.
template <bool SWZ>
__device__ __forceinline__ uint32_t sw(uint32_t i) {
  if constexpr (SWZ) return qas_swz(i);
  return i;
}
__device__ __forceinline__ void apply_u(double2 &a0, double2 &a1, const double2 (&u)[4]) {
  const double2 b0 = cfma(u[1], a1, cmul(u[0], a0));
  const double2 b1 = cfma(u[3], a1, cmul(u[2], a0));
  a0 = b0;
  a1 = b1;
}
__device__ __forceinline__ void apply_udag(double2 &a0, double2 &a1, const double2 (&u)[4]) {
  const double2 b0 = cfma_conj(u[2], a1, cmul_conj(u[0], a0));
  const double2 b1 = cfma_conj(u[3], a1, cmul_conj(u[1], a0));
  a0 = b0;
  a1 = b1;
}
// M_ab += conj(l_a) q_b  (acc = M00 M01 M10 M11 as re,im pairs)
__device__ __forceinline__ void macc(double (&acc)[8], double2 l0, double2 l1, double2 q0, double2 q1) {
  acc[0] = fma(l0.x, q0.x, fma(l0.y, q0.y, acc[0])); acc[1] = fma(l0.x, q0.y, fma(-l0.y, q0.x, acc[1]));
  acc[2] = fma(l0.x, q1.x, fma(l0.y, q1.y, acc[2])); acc[3] = fma(l0.x, q1.y, fma(-l0.y, q1.x, acc[3]));
  acc[4] = fma(l1.x, q0.x, fma(l1.y, q0.y, acc[4])); acc[5] = fma(l1.x, q0.y, fma(-l1.y, q0.x, acc[5]));
  acc[6] = fma(l1.x, q1.x, fma(l1.y, q1.y, acc[6])); acc[7] = fma(l1.x, q1.y, fma(-l1.y, q1.x, acc[7]));
}

// coset enumeration of one pass group from its planner descriptor gd = {P0, dc, X_0 ..}: P(t) = P0 ^
// xor of the X_i selected by the bits of t; the cosets t < 2^dc are the ones that can be non-zero
// (see qas_plan_groups).  xm[c] = xor of the gate masks selected by the bits of slot c.
// NXS = number of X vectors when known at compile time (they then live in registers), -1 otherwise
// (run-time qubit count: read from the descriptor on every use).  GT = descriptor word type.
template <int G, int NXS, typename GT>
struct GroupGeom {
  uint32_t xm[1 << G];
  uint32_t P0, nsup;
  uint32_t X[NXS > 0 ? NXS : 1];
  const GT *gd;
  int nx;
  __device__ __forceinline__ void load(const uint32_t *o, const GT *gd_, int n) {
    gd = gd_;
    nx = n - G;
    P0 = gd[0];
    nsup = 1u << gd[1];
    if constexpr (NXS > 0) {
#pragma unroll
      for (int i = 0; i < NXS; ++i) X[i] = gd[QAS_GD_HDR + i];
    }
    xm[0] = 0u;
#pragma unroll
    for (int c = 1; c < (1 << G); ++c) {
      const int q = c >= 4 ? 2 : (c >= 2 ? 1 : 0);
      xm[c] = xm[c & ~(1 << q)] ^ o[q * QAS_OP_WORDS];
    }
  }
  __device__ __forceinline__ uint32_t base(uint32_t t) const {
    uint32_t P = P0;
    if constexpr (NXS > 0) {
#pragma unroll
      for (int i = 0; i < NXS; ++i) P ^= ((t >> i) & 1u) ? X[i] : 0u;
    } else if constexpr (NXS < 0) {
      for (int i = 0; i < nx; ++i) P ^= ((t >> i) & 1u) ? (uint32_t)gd[QAS_GD_HDR + i] : 0u;
    }
    return P;
  }
};

// warp-level reduce-scatter of a 2x2 cross matrix (8 doubles) over the TW lanes of a team's warp;
// afterwards lanes with (lane & (TW/8-1)) == 0 hold component v = 4*h1 + 2*h2 + h3.
// ACC: add to the value already at dst when `add` is set (later cosets of a multi-iteration pass;
// only used with the shared-memory ring, where every slot has a single writer lane)
template <int TW, bool ACC>
__device__ __forceinline__ void reduce_store_m(double (&acc)[8], int lane, uint32_t tmask, double *dst, bool add) {
  const bool h1 = lane & (TW / 2), h2 = lane & (TW / 4), h3 = lane & (TW / 8);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double send = h1 ? acc[q] : acc[q + 4], keep = h1 ? acc[q + 4] : acc[q];
    acc[q] = keep + shfl_xor_d(send, TW / 2, tmask);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const double send = h2 ? acc[q] : acc[q + 2], keep = h2 ? acc[q + 2] : acc[q];
    acc[q] = keep + shfl_xor_d(send, TW / 4, tmask);
  }
  {
    const double send = h3 ? acc[0] : acc[1], keep = h3 ? acc[1] : acc[0];
    acc[0] = keep + shfl_xor_d(send, TW / 8, tmask);
  }
#pragma unroll
  for (int off = TW / 16; off > 0; off >>= 1) acc[0] += shfl_xor_d(acc[0], off, tmask);
  if ((lane & (TW / 8 - 1)) == 0) {
    double *d = dst + (h1 ? 4 : 0) + (h2 ? 2 : 0) + (h3 ? 1 : 0);
    if constexpr (ACC) *d = add ? *d + acc[0] : acc[0];
    else *d = acc[0];
  }
}

// shared-memory layout per team (bytes, all 16-aligned):
//   psi [DIM] double2 | lam [DIM] double2 (GRAD) | rec [4 + ops_cap*5] u32
//   | gd [ops_cap][2 + n] u16 (shared-memory path; the global-memory path reads the u32 record)
//   | Us [ops_cap][4] double2 (STAGE_U) | red [2][3][WPT][8] f64 (GRAD, multi-warp teams) | ered [WPT] f64
//   | hx [QAS_HD_WORDS] u32 + tcls [32] u32 (shared-memory path: final-support descriptor, class coset ids)
__host__ __device__ constexpr bool qas_stage_u(int N) { return N == 0 || N >= 7; }
// support-restricted H application: below 8 qubits the whole register is a few warp iterations
#define QAS_HSPARSE_MIN_N 8
__host__ __device__ inline size_t eval_team_smem_bytes(int n, int T, bool grad, int ops_cap,
                                                       bool state_in_smem, bool stage_u) {
  const size_t dim = (size_t)1 << n;
  const int wpt = T >= 32 ? T / 32 : 1;
  size_t b = 0;
  if (state_in_smem) b += dim * 16 * (grad ? 2 : 1);
  b += (((size_t)(QAS_REC_HDR + ops_cap * QAS_OP_WORDS) * 4) + 15) & ~(size_t)15;
  if (state_in_smem) b += (((size_t)ops_cap * QAS_GD_STRIDE(n) * 2) + 15) & ~(size_t)15;
  if (stage_u) b += (size_t)ops_cap * 64;
  if (grad && wpt > 1) b += (size_t)2 * 3 * wpt * 64;
  b += (((size_t)wpt * 8) + 15) & ~(size_t)15;
  b += 16;                                   // the team's work-queue slot
  if (state_in_smem && n >= QAS_HSPARSE_MIN_N) b += (size_t)(QAS_HD_WORDS + 32) * 4;
  return b;
}

// ---- one forward pass: G gates on the 2^G amplitudes of every coset that can be non-zero
template <int G, int NXS, typename GT, int T, bool SWZ, bool STAGE_U>
__device__ __forceinline__ void fwd_group(double2 *psi, const uint32_t *o, const GT *gd, int n, int j,
                                          const double2 *Us, const GateTab *Utab, int tid) {
  GroupGeom<G, NXS, GT> gg;
  gg.load(o, gd, n);
  double2 u[G][4];
#pragma unroll
  for (int q = 0; q < G; ++q) {
    const double2 *g = STAGE_U ? (Us + 4 * (j + q)) : Utab[o[q * QAS_OP_WORDS + 3]].u;
#pragma unroll
    for (int e = 0; e < 4; ++e) u[q][e] = g[e];
  }
  const uint32_t rc = (G == 1) ? o[2] : 0u;
  constexpr int UNR = (8 >> G) > 0 ? (8 >> G) : 1;
#pragma unroll UNR
  for (uint32_t t = tid; t < gg.nsup; t += T) {
    const uint32_t P = gg.base(t);
    if (G == 1 && rc && !(__popc(sw<SWZ>(P) & rc) & 1)) continue;   // rc is a mask over index bits
    double2 a[1 << G];
#pragma unroll
    for (int c = 0; c < (1 << G); ++c) a[c] = psi[P ^ gg.xm[c]];
#pragma unroll
    for (int q = G - 1; q >= 0; --q) {       // time order = descending tape index
#pragma unroll
      for (int c = 0; c < (1 << G); ++c)
        if (!((c >> q) & 1)) apply_u(a[c], a[c | (1 << q)], u[q]);
    }
#pragma unroll
    for (int c = 0; c < (1 << G); ++c) psi[P ^ gg.xm[c]] = a[c];
  }
}

// ---- one adjoint pass: per gate (tape order)  psi <- U^ psi ; M += conj(lam) psi ; lam <- U^ lam
// The 2x2 cross matrix M of a parametrised gate is reduced over the team's warp lanes and
// published: single-warp teams write it straight to global memory (mout_op), multi-warp teams leave
// per-warp partials in the shared-memory ring (summed in fixed warp order after the team barrier).
// 3-gate passes reduce M right after each gate (one live accumulator instead of three: the 8 + 8
// amplitudes already take 64 registers); 1- and 2-gate passes accumulate over the thread's cosets
// first and reduce once.
template <int G, int NXS, typename GT, int T, int TW, int WPT, bool SWZ, bool STAGE_U>
__device__ __forceinline__ void adj_group(double2 *psi, double2 *lam, const uint32_t *o, const GT *gd, int n,
                                          int j, const double2 *Us, const GateTab *Utab, uint32_t DIM, int tid,
                                          int lane, int warp_in_team, uint32_t tmask, double *ring,
                                          double *mout_op) {
  constexpr bool EACH = (G == 3);
  GroupGeom<G, NXS, GT> gg;
  gg.load(o, gd, n);
  int npar[G];
#pragma unroll
  for (int q = 0; q < G; ++q) npar[q] = qas_meta_npar(o[q * QAS_OP_WORDS + 4]);
  double M[EACH ? 1 : G][8];
  if constexpr (!EACH) {
#pragma unroll
    for (int q = 0; q < G; ++q)
#pragma unroll
      for (int e = 0; e < 8; ++e) M[q][e] = 0.0;
  }
  const uint32_t rc = (G == 1) ? o[2] : 0u;
  // Only the first 2^dc cosets are visited.  psi is zero outside them, and lambda is not needed there
  // either: the derivative of gate g pairs lambda_g with dU_g psi_{g-1}, which lives on the support
  // of psi_g, and lambda_{g-1} = U_g^ lambda_g on the (smaller) support of psi_{g-1} only reads
  // lambda_g inside the support of psi_g.  So lambda, like psi, is propagated inside the shrinking
  // support and whatever it holds outside is never read again.  (The bound is rounded up to whole
  // warps / teams so that the shuffles below see uniform trip counts.)
  // FULL: the support is everything (uniform start state, deep circuits) -- compile-time trip count, no
  // inside/outside predicates
  auto sweep = [&](auto full_tag) {
  constexpr bool FULL = decltype(full_tag)::value;
  const uint32_t nper = FULL ? (DIM >> G) : min(DIM >> G, (gg.nsup + (uint32_t)(TW - 1)) & ~(uint32_t)(TW - 1));
  constexpr int UNR = EACH ? 1 : ((4 >> G) > 0 ? (4 >> G) : 1);
#pragma unroll UNR
  for (uint32_t t = tid; t < nper; t += T) {
    const uint32_t P = gg.base(t);
    if (G == 1 && rc && !(__popc(sw<SWZ>(P) & rc) & 1)) continue;   // rc is a mask over index bits
    // ins: this coset is inside the support (the rounding above admits a few that are not).  wins: some
    // lane of this warp is inside (warp-uniform for whole-warp teams; sub-warp teams always reduce)
    const bool ins = FULL || t < gg.nsup;
    const bool wins = FULL || T < 32 || (t & ~31u) < gg.nsup;
    (void)wins;
    double2 a[1 << G], l[1 << G];
#pragma unroll
    for (int c = 0; c < (1 << G); ++c) {
      const uint32_t x = P ^ gg.xm[c];
      a[c] = ins ? psi[x] : make_double2(0.0, 0.0);
      l[c] = lam[x];
    }
#pragma unroll
    for (int q = 0; q < G; ++q) {
      double2 u[4];
      const double2 *g = STAGE_U ? (Us + 4 * (j + q)) : Utab[o[q * QAS_OP_WORDS + 3]].u;
#pragma unroll
      for (int e = 0; e < 4; ++e) u[e] = g[e];
      double (&Mq)[8] = M[EACH ? 0 : q];
      if constexpr (EACH) {
#pragma unroll
        for (int e = 0; e < 8; ++e) Mq[e] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < (1 << G); ++c) {
        if ((c >> q) & 1) continue;
        const int c1 = c | (1 << q);
        if (ins) {
          apply_udag(a[c], a[c1], u);
          if (npar[q]) macc(Mq, l[c], l[c1], a[c], a[c1]);
        }
        apply_udag(l[c], l[c1], u);
      }
      if constexpr (EACH) {
        if (npar[q] && wins) {
          if constexpr (WPT == 1) reduce_store_m<TW, false>(Mq, lane, tmask, mout_op + (size_t)q * 8, false);
          else reduce_store_m<TW, true>(Mq, lane, tmask, ring + ((size_t)q * WPT + warp_in_team) * 8, t != (uint32_t)tid);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < (1 << G); ++c) {
      const uint32_t x = P ^ gg.xm[c];
      if (ins) psi[x] = a[c];
      lam[x] = l[c];
    }
  }
  };
  // the second copy only pays where the whole register is one or two iterations (n < 8; measured:
  // QAOA-7q 0.299 -> 0.277 ms, but H2O-8q 1.05 -> 1.24 ms with both copies in the kernel)
  constexpr bool DUP = NXS > 0 && NXS + G < 8;
  if constexpr (DUP) { if (gg.nsup >= (DIM >> G)) sweep(std::true_type{}); else sweep(std::false_type{}); }
  else sweep(std::false_type{});
  if constexpr (!EACH) {
    const bool wins0 = T < 32 || ((uint32_t)tid & ~31u) < gg.nsup;
    if (wins0) {
#pragma unroll
      for (int q = 0; q < G; ++q)
        if (npar[q]) reduce_store_m<TW, false>(M[q], lane, tmask, WPT == 1 ? mout_op + (size_t)q * 8
                                                                              : ring + ((size_t)q * WPT + warp_in_team) * 8, false);
    }
  }
}

// N > 0: compile-time qubit count, state in shared memory.  N == 0: run-time qubit count,
// state in global memory (one team per CTA, L2/HBM resident) -- same code path otherwise.
// GMAX = gates per fused pass, RB = log2(amplitudes per thread) in the H application, PD = depth of
// the register ring that prefetches sign-table rows (L2 latency is the exposed cost there)
template <int N, int T, bool GRAD, int GMAX, int RB, int PD, int MINB, int BLK = 256>
__global__ void __launch_bounds__(BLK, MINB) eval_kernel(const EvalParams P) {
  constexpr int TEAMS = BLK / T;
  constexpr int WPT = T >= 32 ? T / 32 : 1;      // warps per team
  constexpr int TW = T >= 32 ? 32 : T;           // lanes of one warp that belong to the team
  constexpr bool STAGE_U = qas_stage_u(N);
  constexpr bool SWZ = N != 0;
  const int n = N ? N : P.n;
  const uint32_t DIM = 1u << n;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int team = threadIdx.x / T, tid = threadIdx.x % T;
  const int lane = threadIdx.x & 31, warp_in_team = tid >> 5;
  // sub-warp teams synchronise and shuffle with their own lane mask, so the teams that share a
  // warp run independent tapes without waiting on each other
  const uint32_t tmask = T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << ((lane / TW) * TW));
  const size_t team_bytes = eval_team_smem_bytes(n, T, GRAD, P.ops_cap, N != 0, STAGE_U);
  unsigned char *base = smem_raw + (size_t)team * team_bytes;
  double2 *psi, *lam = nullptr;
  if constexpr (N != 0) {
    psi = reinterpret_cast<double2 *>(base);
    base += (size_t)DIM * 16;
    if constexpr (GRAD) { lam = reinterpret_cast<double2 *>(base); base += (size_t)DIM * 16; }
  } else {
    psi = P.gstate + (size_t)blockIdx.x * DIM * (GRAD ? 2 : 1);
    if constexpr (GRAD) lam = psi + DIM;
  }
  const int rec_words = QAS_REC_HDR + P.ops_cap * QAS_OP_WORDS;       // staged part of a tape record
  const size_t grec_words = P.rec_stride;                             // its full length in global memory
  const int gstride = QAS_GD_STRIDE(n);
  uint32_t *rec = reinterpret_cast<uint32_t *>(base);
  uint32_t *ops = rec + QAS_REC_HDR;
  base += (((size_t)rec_words * 4) + 15) & ~(size_t)15;
  // group descriptors: 16-bit copies in shared memory (N != 0) or the 32-bit record in global memory
  using GT = typename std::conditional<N != 0, uint16_t, uint32_t>::type;
  constexpr int NXS3 = N != 0 ? N - 3 : -1, NXS2 = N != 0 ? N - 2 : -1, NXS1 = N != 0 ? N - 1 : -1;
  uint16_t *gd16 = reinterpret_cast<uint16_t *>(base);
  if constexpr (N != 0) base += (((size_t)P.ops_cap * gstride * 2) + 15) & ~(size_t)15;
  double2 *Us = reinterpret_cast<double2 *>(base);
  if constexpr (STAGE_U) base += (size_t)P.ops_cap * 64;
  double *red = nullptr;
  if constexpr (GRAD && WPT > 1) { red = reinterpret_cast<double *>(base); base += (size_t)2 * 3 * WPT * 64; }
  double *ered = reinterpret_cast<double *>(base);
  base += (((size_t)WPT * 8) + 15) & ~(size_t)15;
  int *bslot = reinterpret_cast<int *>(base);
  base += 16;
  uint32_t *hx = reinterpret_cast<uint32_t *>(base);      // N != 0 only
  uint32_t *tcls = hx + QAS_HD_WORDS;
  // slice descriptors, shared by the CTA's teams
  uint32_t *sdesc = reinterpret_cast<uint32_t *>(smem_raw + (size_t)TEAMS * team_bytes);
  for (int s = threadIdx.x; s < P.C_real + P.C_imag; s += BLK) sdesc[s] = P.cdesc[s];
  __syncthreads();

  // Every team pulls circuits from one global queue: circuit lengths differ (2-14 gates on the LiH
  // batch), and with a fixed circuit -> CTA assignment the short circuit's team would idle until its
  // CTA mates finish.  The grid is persistent (CTAs per SM x SMs).
  for (;;) {
    if (tid == 0) *bslot = atomicAdd(P.work_counter, 1);
    team_sync<T>(team, tmask);
    const int qslot = *bslot;
    if (qslot >= (P.count ? *P.count : P.B)) break;
    const int b = P.order ? P.order[qslot] : qslot;
    constexpr bool active = true;

    long long clk_prev = 0;
    if (P.phase_clk && tid == 0) clk_prev = clock64();
    // ---- 1. fetch the compiled tape record, zero the state / the gradient row
    {
      const uint32_t *grec = P.tape + (size_t)b * grec_words;
      const int used = (int)(QAS_REC_HDR + grec[0] * QAS_OP_WORDS);
      for (int w = tid; w < QAS_REC_HDR || w < used; w += T) rec[w] = grec[w];
      if constexpr (N != 0) {
        const uint32_t *ggd = grec + QAS_REC_HDR + (size_t)P.ops_cap * QAS_OP_WORDS;
        const int gused = (int)grec[3] * gstride;
        for (int w = tid; w < gused; w += T) gd16[w] = (uint16_t)ggd[w];
        if constexpr (N >= QAS_HSPARSE_MIN_N) {
          const uint32_t *ghd = grec + qas_rec_hd_offset(P.ops_cap, n);
          for (int w = tid; w < QAS_HD_WORDS; w += T) hx[w] = ghd[w];
        }
      }
      const double amp0 = (P.init_kind == QAS_INIT_PLUS) ? exp2(-0.5 * (double)n) : 0.0;
      for (uint32_t i = tid; i < DIM; i += T) psi[i] = make_double2(amp0, 0.0);
    }
    team_sync<T>(team, tmask);
    const int nops = (int)rec[0];
    const int status = (int)rec[1];
    const int ngroups = (int)rec[3];
    if (tid == 0 && P.init_kind != QAS_INIT_PLUS) psi[sw<SWZ>(rec[2])] = make_double2(1.0, 0.0);
    if constexpr (STAGE_U) {   // gate matrices of this circuit -> shared memory (one round trip)
      for (int w = tid; w < nops * 4; w += T) Us[w] = P.U[ops[(w >> 2) * QAS_OP_WORDS + 3]].u[w & 3];
    }
    const int nloop = ngroups;
    const GT *gdb;
    if constexpr (N != 0) gdb = gd16;
    else gdb = P.tape + (size_t)b * grec_words + QAS_REC_HDR + (size_t)P.ops_cap * QAS_OP_WORDS;
    team_sync<T>(team, tmask);

    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 0, (unsigned long long)(c_ - clk_prev)); clk_prev = c_; }
    // ---- 2. forward sweep (tape is stored last-gate-first: walk the groups from the end)
    {
      int e = nops - 1;
      for (int it = 0; it < nloop; ++it) {
        if (e >= 0) {
          const int g = (int)QAS_META_GLAST(ops[(size_t)e * QAS_OP_WORDS + 4]);
          const int j = e - g + 1;
          const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
          const GT *gd = gdb + (size_t)(nloop - 1 - it) * gstride;
          if (qas_meta_kind(o[4]) == QAS_K_U1) {
            if (GMAX >= 3 && g == 3) { if constexpr (GMAX >= 3) fwd_group<3, NXS3, GT, T, SWZ, STAGE_U>(psi, o, gd, n, j, Us, P.U, tid); }
            else if (GMAX >= 2 && g == 2) { if constexpr (GMAX >= 2) fwd_group<2, NXS2, GT, T, SWZ, STAGE_U>(psi, o, gd, n, j, Us, P.U, tid); }
            else fwd_group<1, NXS1, GT, T, SWZ, STAGE_U>(psi, o, gd, n, j, Us, P.U, tid);
          } else {
            const double2 *gm = STAGE_U ? (Us + 4 * j) : P.U[o[3]].u;
            const double2 d0 = gm[0], d1 = gm[3];
            const uint32_t r = o[1];
#pragma unroll 4
            for (uint32_t i = tid; i < DIM; i += T) {
              const uint32_t x = sw<SWZ>(i);
              psi[x] = cmul((__popc(i & r) & 1) ? d1 : d0, psi[x]);
            }
          }
          e = j - 1;
        }
        team_sync<T>(team, tmask);
      }
    }

    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 1, (unsigned long long)(c_ - clk_prev)); clk_prev = c_; }
    // ---- 3. lambda = H psi, E = Re <psi|lambda>   (register-blocked over X-mask classes)
    double e_part = 0.0;
    {
      constexpr int NB = 1 << RB;
      const uint32_t nrest = DIM >> RB;
      const char *psib = reinterpret_cast<const char *>(psi);
      const int C = P.C_real + P.C_imag;
      bool sparse = false;
      constexpr bool HSPARSE = N >= QAS_HSPARSE_MIN_N;
      if constexpr (HSPARSE) sparse = P.hs_min > 0 && C <= 32 && (int)hx[1] + P.hs_min <= N - RB;
      if (sparse) {
        // Support-restricted application (tape.h, "Support of the FINAL state"): the cosets are walked in
        // the basis adapted to Dbar.  Coset u(t) only hears from the classes whose Dbar-coset id equals
        // t >> dbar (every other partner coset holds zeros), and lambda = H psi is only needed ON the
        // support (energy = <psi|lambda>; the adjoint sweep never leaves it, see adj_group): the cosets
        // t < 2^dbar, from the classes with id 0.
        if constexpr (HSPARSE) {
          constexpr int NU = N - RB;
          const int dbar = (int)hx[1], nchk = NU - dbar;
          for (int cls = tid; cls < C; cls += T) {
            const uint32_t ru = P.cdesc_u[cls];
            uint32_t tau = 0;
            for (int i = 0; i < nchk; ++i) tau |= (uint32_t)(__popc(ru & hx[2 + NU + i]) & 1) << i;
            tcls[cls] = tau;
          }
          team_sync<T>(team, tmask);
          for (uint32_t t = tid; t < nrest; t += T) {
            uint32_t u = hx[0];
#pragma unroll
            for (int i = 0; i < NU; ++i) u ^= ((t >> i) & 1u) ? hx[2 + i] : 0u;
            const uint32_t sb16 = sw<SWZ>(qas_deposit_rest(u, RB, P.rb0, P.rb1, P.rb2)) << 4;
            double2 acc[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = make_double2(0.0, 0.0);
            if (t >> dbar) {         // outside the support: lambda is never read there (adj_group); keep it defined
              if constexpr (GRAD) {
#pragma unroll
                for (int j = 0; j < NB; ++j)
                  *reinterpret_cast<double2 *>(reinterpret_cast<char *>(lam) + (sb16 ^ P.wc16.w[j])) = acc[j];
              }
              continue;
            }
            const double *vp = P.vtab + (size_t)u * (NB >= 2 ? 2 : 1);
            for (int cls = 0; cls < C; ++cls) {
              if (tcls[cls] != 0u) continue;
              const uint32_t r16 = sb16 ^ sdesc[cls];
              double v[NB][NB];
#pragma unroll
              for (int jw = 0; jw < NB; ++jw) load_vrow<NB>(v[jw], vp + ((size_t)(cls * NB + jw) << n), nrest);
              double2 a[NB];
#pragma unroll
              for (int j = 0; j < NB; ++j) a[j] = *reinterpret_cast<const double2 *>(psib + (r16 ^ P.wc16.w[j]));
              const bool imag = cls >= P.C_real;
#pragma unroll
              for (int jw = 0; jw < NB; ++jw) hacc_row<NB>(acc, a, v[jw], jw, imag);
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const uint32_t off = sb16 ^ P.wc16.w[j];
              const double2 ai = *reinterpret_cast<const double2 *>(psib + off);
              e_part = fma(ai.x, acc[j].x, fma(ai.y, acc[j].y, e_part));
              if constexpr (GRAD) *reinterpret_cast<double2 *>(reinterpret_cast<char *>(lam) + off) = acc[j];
            }
          }
        }
      } else {
      for (uint32_t u = tid; u < nrest; u += T) {
        const uint32_t sb16 = sw<SWZ>(qas_deposit_rest(u, RB, P.rb0, P.rb1, P.rb2)) << 4;
        double2 acc[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = make_double2(0.0, 0.0);
        // sign-table rows are prefetched PD rows ahead through a register ring (L2 latency); the
        // table ends with zero rows, so the prefetch never needs a bounds test
        const double *vp = P.vtab + (size_t)u * (NB >= 2 ? 2 : 1);
        double ring[PD][NB];
#pragma unroll
        for (int d = 0; d < PD; ++d) load_vrow<NB>(ring[d], vp + ((size_t)d << n), nrest);
        for (int cls = 0; cls < C; ++cls) {
          const uint32_t r16 = sb16 ^ sdesc[cls];
          double2 a[NB];
#pragma unroll
          for (int j = 0; j < NB; ++j) a[j] = *reinterpret_cast<const double2 *>(psib + (r16 ^ P.wc16.w[j]));
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
          const double2 ai = *reinterpret_cast<const double2 *>(psib + off);
          e_part = fma(ai.x, acc[j].x, fma(ai.y, acc[j].y, e_part));
          if constexpr (GRAD) *reinterpret_cast<double2 *>(reinterpret_cast<char *>(lam) + off) = acc[j];
        }
      }
      }
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
      team_sync<T>(team, tmask);      // lambda visible to the whole team
    }
    if (tid == 0 && active) P.energy[b] = status ? __longlong_as_double(0x7ff8000000000000ll) : e_part;

    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 2, (unsigned long long)(c_ - clk_prev)); clk_prev = c_; }
    // ---- 4. adjoint sweep (tape order = reverse time order)
    if constexpr (GRAD) {
      int j = 0;
      for (int it = 0; it < nloop; ++it) {
        if (j < nops) {
          const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
          const int g = (int)QAS_META_GFIRST(o[4]);
          double *ring = WPT > 1 ? red + (size_t)(it & 1) * 3 * WPT * 8 : nullptr;
          double *mop = P.mout + ((size_t)b * P.ops_cap + j) * 8;
          const GT *gd = gdb + (size_t)it * gstride;
          // warps of the team that own cosets inside the support (they publish partial M's)
          int nact = WPT;
          if (qas_meta_kind(o[4]) == QAS_K_U1) {
            nact = (int)min((uint32_t)WPT, ((1u << gd[1]) + 31u) >> 5);
            if (GMAX >= 3 && g == 3) { if constexpr (GMAX >= 3) adj_group<3, NXS3, GT, T, TW, WPT, SWZ, STAGE_U>(psi, lam, o, gd, n, j, Us, P.U, DIM, tid, lane, warp_in_team, tmask, ring, mop); }
            else if (GMAX >= 2 && g == 2) { if constexpr (GMAX >= 2) adj_group<2, NXS2, GT, T, TW, WPT, SWZ, STAGE_U>(psi, lam, o, gd, n, j, Us, P.U, DIM, tid, lane, warp_in_team, tmask, ring, mop); }
            else adj_group<1, NXS1, GT, T, TW, WPT, SWZ, STAGE_U>(psi, lam, o, gd, n, j, Us, P.U, DIM, tid, lane, warp_in_team, tmask, ring, mop);
          } else {
            const double2 *gm = STAGE_U ? (Us + 4 * j) : P.U[o[3]].u;
            const double2 d0 = gm[0], d1 = gm[3];
            const uint32_t r = o[1];
            const int npar = qas_meta_npar(o[4]);
            double acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.0;
#pragma unroll 2
            for (uint32_t i = tid; i < DIM; i += T) {
              const bool hi = __popc(i & r) & 1;
              const uint32_t x = sw<SWZ>(i);
              const double2 d = hi ? d1 : d0;
              const double2 q = cmul_conj(d, psi[x]);
              const double2 lv = lam[x];
              psi[x] = q;
              if (npar) {
                const double re = fma(lv.x, q.x, lv.y * q.y), im = fma(lv.x, q.y, -lv.y * q.x);
                if (hi) { acc[6] += re; acc[7] += im; } else { acc[0] += re; acc[1] += im; }
              }
              lam[x] = cmul_conj(d, lv);
            }
            if (npar) reduce_store_m<TW, false>(acc, lane, tmask, WPT == 1 ? mop : ring + (size_t)warp_in_team * 8, false);
          }
          if constexpr (WPT > 1) {
            team_sync<T>(team, tmask);
            // fixed-order sum of the warps' partials -> M of each parametrised op of the group
            if (tid < 8 * g) {
              const int q = tid >> 3, v = tid & 7;
              if (qas_meta_npar(o[q * QAS_OP_WORDS + 4])) {
                double m = 0.0;
                for (int wp = 0; wp < nact; ++wp) m += ring[((size_t)q * WPT + wp) * 8 + v];
                mop[(size_t)q * 8 + v] = m;
              }
            }
          }
          j += g;
        }
        if constexpr (WPT == 1) team_sync<T>(team, tmask);
      }
    }
    if (P.phase_clk && tid == 0) { const long long c_ = clock64(); atomicAdd(P.phase_clk + 3, (unsigned long long)(c_ - clk_prev)); }
    team_sync<T>(team, tmask);   // state buffers are reused by the next circuit of this team
  }
}

// ------------------------------------------------------------------------------ small registers
// n <= 5: ONE THREAD per circuit, and the whole evaluation in one kernel -- tape compile, forward
// sweep, H application, adjoint sweep and the parameter derivatives.  The statevectors and the
// packed tape are thread-private columns of shared memory (element i of thread t at [i * BLK + t]:
// every lane of a warp touches the same row, so all accesses are conflict-free and nothing is ever
// shared between threads -- no barriers, no shuffles, no tape round trip through L2/HBM).  Sign-table
// values and class descriptors are the same for every lane (broadcast loads); the 2x2 cross matrix
// of a gate is a thread-local sum, contracted with dU/dtheta on the spot.
// The sub-warp-team path this replaces paid a 1.8 KB tape record per circuit (three more kernels:
// compile, memset, finalize) and per-pass __syncwarp latency.
struct QasPackedSink {            // op -> {m | r<<5 | rc<<10 | kind<<15 | npar<<16 | depth<<18, tab}
  uint2 *ops;                     // thread-private column, stride BLK
  int stride, cap, n, overflow;
  __device__ __forceinline__ void emit(uint32_t m, uint32_t r, uint32_t rc, uint32_t tab, uint32_t meta) {
    if (n >= cap) { overflow = 1; return; }
    ops[(size_t)n * stride] = make_uint2(m | (r << 5) | (rc << 10) | (qas_meta_kind(meta) << 15) |
                                         ((uint32_t)qas_meta_npar(meta) << 16) | ((uint32_t)qas_meta_depth(meta) << 18), tab);
    n++;
  }
};
// shared memory: psi | lam (GRAD) | packed ops [ops_cap][BLK] | frame [BLK][11] u32 | pool [c]
__host__ __device__ inline size_t eval_small_smem_bytes(int n, int blk, bool grad, int ops_cap, int c) {
  return ((size_t)16 << n) * blk * (grad ? 2 : 1) + (size_t)ops_cap * 8 * blk + (size_t)blk * 11 * 4 +
         (((size_t)c * sizeof(qas_pool_entry)) + 15) / 16 * 16;
}

template <int N, bool GRAD, int BLK>
__global__ void __launch_bounds__(BLK) eval_small_kernel(const EvalParams P) {
  constexpr uint32_t DIM = 1u << N, HALF = DIM >> 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *psi = reinterpret_cast<double2 *>(smem_raw) + threadIdx.x;
  double2 *lam = psi + (size_t)DIM * BLK;           // only present when GRAD
  unsigned char *sp = smem_raw + ((size_t)16 << N) * BLK * (GRAD ? 2 : 1);
  uint2 *ops = reinterpret_cast<uint2 *>(sp) + threadIdx.x;
  sp += (size_t)P.ops_cap * 8 * BLK;
  uint32_t *frame = reinterpret_cast<uint32_t *>(sp) + (size_t)threadIdx.x * 11;   // odd stride: conflict-free
  sp += (size_t)BLK * 11 * 4;
  qas_pool_entry *spool = reinterpret_cast<qas_pool_entry *>(sp);
  for (int x = threadIdx.x; x < P.c; x += BLK) spool[x] = P.pool[x];
  __syncthreads();
  const int b = blockIdx.x * BLK + threadIdx.x;
  if (b >= P.B) return;
  // ---- tape compile (CNOT-folded frame, tape.h), ops packed into this thread's column
  int nops, status;
  uint32_t init_index;
  {
    uint32_t *col = frame, *row = frame + 5;
    QasPackedSink sink{ops, BLK, P.ops_cap, 0, 0};
    status = qas_compile_circuit(P.k + (size_t)b * P.p, P.p, spool, P.c, N, col, row, sink);
    nops = status ? 0 : sink.n;
    init_index = (P.init_kind == QAS_INIT_BASIS && !status) ? qas_frame_map_basis(col, N, P.init_basis) : 0u;
  }
  {
    const double amp0 = (P.init_kind == QAS_INIT_PLUS) ? exp2(-0.5 * (double)N) : 0.0;
#pragma unroll
    for (uint32_t i = 0; i < DIM; ++i) psi[(size_t)i * BLK] = make_double2(amp0, 0.0);
    if (P.init_kind != QAS_INIT_PLUS) psi[(size_t)init_index * BLK] = make_double2(1.0, 0.0);
  }
  // ---- forward sweep (ops are stored last-gate-first)
  for (int j = nops - 1; j >= 0; --j) {
    const uint2 o = ops[(size_t)j * BLK];
    const uint32_t m = o.x & 31u, r = (o.x >> 5) & 31u, rc = (o.x >> 10) & 31u;
    const GateTab &gt = P.U[o.y];
    if (((o.x >> 15) & 1u) == QAS_K_U1) {
      double2 u[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) u[e] = gt.u[e];
      const int pb = 31 - __clz(r);
#pragma unroll
      for (uint32_t t = 0; t < HALF; ++t) {
        uint32_t p0 = insert_zero(t, pb);
        p0 |= (uint32_t)(__popc(p0 & r) & 1) << pb;
        if (rc && !(__popc(p0 & rc) & 1)) continue;
        const uint32_t p1 = p0 ^ m;
        double2 a0 = psi[(size_t)p0 * BLK], a1 = psi[(size_t)p1 * BLK];
        apply_u(a0, a1, u);
        psi[(size_t)p0 * BLK] = a0;
        psi[(size_t)p1 * BLK] = a1;
      }
    } else {
      const double2 d0 = gt.u[0], d1 = gt.u[3];
#pragma unroll
      for (uint32_t i = 0; i < DIM; ++i)
        psi[(size_t)i * BLK] = cmul((__popc(i & r) & 1) ? d1 : d0, psi[(size_t)i * BLK]);
    }
  }
  // ---- lambda = H psi (class by class, same tables as the team kernel), E = Re <psi|lambda>
  double e_acc = 0.0;
  {
    constexpr int RB = N >= 4 ? 1 : 0;              // = the host's choice for these qubit counts
    constexpr int NB = 1 << RB;
    constexpr uint32_t nrest = DIM >> RB;
    const int C = P.C_real + P.C_imag;
    for (uint32_t u = 0; u < nrest; ++u) {
      const uint32_t base_i = qas_deposit_rest(u, RB, P.rb0, P.rb1, P.rb2);
      double2 acc[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) acc[q] = make_double2(0.0, 0.0);
      const double *vp = P.vtab + (size_t)u * (NB >= 2 ? 2 : 1);
      for (int cls = 0; cls < C; ++cls) {
        const uint32_t rho = P.cdesc[cls] >> 4;
        const bool imag = cls >= P.C_real;
        double2 a[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) a[q] = psi[(size_t)(base_i ^ rho ^ (P.wc16.w[q] >> 4)) * BLK];
#pragma unroll
        for (int jw = 0; jw < NB; ++jw) {
          double v[NB];
          load_vrow<NB>(v, vp + ((size_t)(cls * NB + jw) << N), nrest);
          hacc_row<NB>(acc, a, v, jw, imag);
        }
      }
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        const uint32_t i = base_i ^ (P.wc16.w[q] >> 4);
        const double2 ai = psi[(size_t)i * BLK];
        e_acc = fma(ai.x, acc[q].x, fma(ai.y, acc[q].y, e_acc));
        if constexpr (GRAD) lam[(size_t)i * BLK] = acc[q];
      }
    }
  }
  P.energy[b] = status ? __longlong_as_double(0x7ff8000000000000ll) : e_acc;
  // ---- adjoint sweep; derivatives written straight into grad[b][depth][j] (zero-filled beforehand)
  if constexpr (GRAD) {
    for (int j = 0; j < nops; ++j) {
      const uint2 o = ops[(size_t)j * BLK];
      const uint32_t m = o.x & 31u, r = (o.x >> 5) & 31u, rc = (o.x >> 10) & 31u;
      const int npar = (int)((o.x >> 16) & 3u);
      const GateTab &gt = P.U[o.y];
      double M[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) M[e] = 0.0;
      if (((o.x >> 15) & 1u) == QAS_K_U1) {
        double2 u[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) u[e] = gt.u[e];
        const int pb = 31 - __clz(r);
#pragma unroll
        for (uint32_t t = 0; t < HALF; ++t) {
          uint32_t p0 = insert_zero(t, pb);
          p0 |= (uint32_t)(__popc(p0 & r) & 1) << pb;
          if (rc && !(__popc(p0 & rc) & 1)) continue;
          const uint32_t p1 = p0 ^ m;
          double2 a0 = psi[(size_t)p0 * BLK], a1 = psi[(size_t)p1 * BLK];
          double2 l0 = lam[(size_t)p0 * BLK], l1 = lam[(size_t)p1 * BLK];
          apply_udag(a0, a1, u);
          if (npar) macc(M, l0, l1, a0, a1);
          apply_udag(l0, l1, u);
          psi[(size_t)p0 * BLK] = a0; psi[(size_t)p1 * BLK] = a1;
          lam[(size_t)p0 * BLK] = l0; lam[(size_t)p1 * BLK] = l1;
        }
      } else {
        const double2 d0 = gt.u[0], d1 = gt.u[3];
#pragma unroll
        for (uint32_t i = 0; i < DIM; ++i) {
          const bool hi = __popc(i & r) & 1;
          const double2 d = hi ? d1 : d0;
          const double2 q = cmul_conj(d, psi[(size_t)i * BLK]);
          const double2 lv = lam[(size_t)i * BLK];
          psi[(size_t)i * BLK] = q;
          const double re = fma(lv.x, q.x, lv.y * q.y), im = fma(lv.x, q.y, -lv.y * q.x);
          if (hi) { M[6] += re; M[7] += im; } else { M[0] += re; M[1] += im; }
          lam[(size_t)i * BLK] = cmul_conj(d, lv);
        }
      }
      if (npar) {
        const GateDTab &dt = P.dU[o.y];
        double *g = P.grad + ((size_t)b * P.p + (o.x >> 18)) * P.l;
        for (int jj = 0; jj < npar; ++jj) {
          double part = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 dq = dt.d[jj][q];
            part += dq.x * M[2 * q] - dq.y * M[2 * q + 1];
          }
          g[jj] = 2.0 * part;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------ gradient finalize
// dE/dtheta_j = 2 Re sum_ab (dU_j)_ab M_ab for every parametrised op of every circuit: one thread
// per (circuit, tape op), all of its <= 3 parameter slots.  Kept out of the evaluation kernel so that no team
// waits on the dU table (L2 latency) between passes.  grad must be zero-filled beforehand.
// order / count: the circuits the classic kernel evaluated (the full bin of a binned call), or null for all B.
static __global__ void grad_finalize_kernel(const uint32_t *tape, size_t rec_stride, const double *mout, const GateDTab *dU, int B,
                                     int p, int l, int ops_cap, double *grad, const int *order, const int *count) {
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= (size_t)B * ops_cap) return;
  const int qi = (int)(x / ops_cap), oi = (int)(x % ops_cap);
  if (count && qi >= *count) return;
  const int b = order ? order[qi] : qi;
  const uint32_t *rec = tape + (size_t)b * rec_stride;
  if (oi >= (int)rec[0] || rec[1]) return;
  const uint32_t *o = rec + QAS_REC_HDR + (size_t)oi * QAS_OP_WORDS;
  const uint32_t mt = o[4];
  const int npar = min(qas_meta_npar(mt), l);
  if (!npar) return;
  const double2 *mp = reinterpret_cast<const double2 *>(mout + ((size_t)b * ops_cap + oi) * 8);
  double2 m[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) m[q] = mp[q];
  double *g = grad + ((size_t)b * p + qas_meta_depth(mt)) * l;
  for (int jj = 0; jj < npar; ++jj) {
    const double2 *d = dU[o[3]].d[jj];
    double part = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) part += d[q].x * m[q].x - d[q].y * m[q].y;
    g[jj] = 2.0 * part;
  }
}

// ------------------------------------------------------------------------------ batch reduce
// Deterministic restatement of  stack -> nan_to_num -> mean/sum over the batch
// (reference qas/mcts/mcts.py:345-347).  Stage 1: one CTA per chunk of circuits and depth
// tile; thread (i,j) owns acc[i][*][j] and walks the chunk in order.  Stage 2: fixed-order
// sum of the chunk partials.
__device__ __forceinline__ double nan_to_num_d(double x) {
  if (x != x) return 0.0;
  if (isinf(x)) return x > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
  return x;
}

// Sparse mode (energy and npar_key given): grad was NOT zero-filled; entry (b, i, j) is read iff the kernels wrote it, i.e.
// j < npar_key[k[b][i]] and the circuit's energy is not NaN (qas_eval_batch, "sparse_reduce").  Half of the keys of a
// Rot + CNOT pool carry no parameters, so this also halves the gradient bytes the reduction reads.
static __global__ void grad_chunk_reduce_kernel(const int32_t *k, const double *grad, int B, int p, int c,
                                         int l, int chunk, int depth_tile, int ngroups, double *partial,
                                         const double *energy, const uint8_t *npar_key) {
  extern __shared__ double acc_all[];             // [ngroups][depth_tile][c][l] | npar per key [c] u8
  const int ch = blockIdx.x, i0 = blockIdx.y * depth_tile;
  const int nd = min(depth_tile, p - i0);
  const int gthreads = blockDim.x / ngroups, grp = threadIdx.x / gthreads, gt = threadIdx.x % gthreads;
  const size_t asz = (size_t)depth_tile * c * l;
  uint8_t *snp = reinterpret_cast<uint8_t *>(acc_all + asz * ngroups);
  const bool sparse = energy != nullptr && npar_key != nullptr;
  for (size_t x = threadIdx.x; x < asz * ngroups; x += blockDim.x) acc_all[x] = 0.0;
  for (int x = threadIdx.x; x < c; x += blockDim.x) snp[x] = sparse ? npar_key[x] : (uint8_t)l;
  __syncthreads();
  double *acc = acc_all + (size_t)grp * asz;
  const int b_lo = min(B, (ch * ngroups + grp) * chunk), b_hi = min(B, b_lo + chunk);
  for (int x = gt; x < nd * l; x += gthreads) {
    const int di = x / l, j = x % l;
    double *a = acc + (size_t)di * c * l + j;
    // eight circuits' loads in flight per step; the accumulation order stays b ascending within the group
    int b = b_lo;
    for (; b + 8 <= b_hi; b += 8) {
      int key[8];
      double gval[8];
      double en[8];
      // keys, energies and gradient entries are requested together (one round trip per step): an entry the kernels did
      // not write is loaded as whatever the buffer holds and dropped below
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        key[q] = k[(size_t)(b + q) * p + i0 + di];
        en[q] = sparse ? energy[b + q] : 0.0;
        gval[q] = grad[((size_t)(b + q) * p + i0 + di) * l + j];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const bool on = key[q] >= 0 && key[q] < c && j < (int)snp[key[q]];
        if (!on || en[q] != en[q]) key[q] = -1;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (key[q] >= 0) a[key[q] * l] += nan_to_num_d(gval[q]);
    }
    for (; b < b_hi; ++b) {
      const int key = k[(size_t)b * p + i0 + di];
      if (key < 0 || key >= c || j >= (int)snp[key]) continue;
      if (sparse) { const double en = energy[b]; if (en != en) continue; }
      const double gval = grad[((size_t)b * p + i0 + di) * l + j];
      a[key * l] += nan_to_num_d(gval);
    }
  }
  __syncthreads();
  double *out = partial + ((size_t)ch * p + i0) * c * l;
  for (int x = threadIdx.x; x < nd * c * l; x += blockDim.x) {
    double s = acc_all[x];
    for (int g = 1; g < ngroups; ++g) s += acc_all[(size_t)g * asz + x];      // fixed group order
    out[x] = s;
  }
}

// one warp per output element: lane q sums chunks q, q+32, ... in order, then a fixed shuffle tree
// combines the 32 lanes -- the summation order depends only on (nchunks), never on scheduling
static __global__ void grad_final_reduce_kernel(const double *partial, int nchunks, size_t stride, size_t total,
                                                double scale, double *out) {
  const size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (x >= total) return;
  double s = 0.0;
  for (int ch = lane; ch < nchunks; ch += 32) s += partial[(size_t)ch * stride + x];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[x] = s * scale;
}

// ------------------------------------------------------------------------------ optimiser step
// One Adam step with the rule of qml.AdamOptimizer (the optimiser the reference passes to search /
// circuitModelTuning, qas/mcts/mcts.py:226,351-352,387,407-408), on the device so that a tuning
// loop never moves the parameter tensor over PCIe:
//   g <- nan_to_num(grad) + noise_factor * noise        (mcts.py:346,349-350 / :404-406)
//   m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g g ;  x <- x - step * m / (sqrt(v) + eps)
// step = stepsize * sqrt(1 - b2^t) / (1 - b1^t) is computed on the host in double precision.
static __global__ void adam_step_kernel(double *x, const double *grad, double *m, double *v, const double *noise,
                                 size_t n, double step, double b1, double b2, double eps, double noise_factor) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double g = nan_to_num_d(grad[i]);
  if (noise) g += noise_factor * noise[i];
  const double mi = b1 * m[i] + (1.0 - b1) * g;
  const double vi = b2 * v[i] + (1.0 - b2) * g * g;
  m[i] = mi;
  v[i] = vi;
  x[i] = x[i] - step * mi / (sqrt(vi) + eps);
}

// ------------------------------------------------------------------------------ FP64 peak
// 8 independent DFMA chains per thread; 2 flop per DFMA.
static __global__ void dfma_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
