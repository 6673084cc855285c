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
#include "tape.h"

struct __align__(16) GateTab { double2 u[4]; };        // U00 U01 U10 U11
struct __align__(16) GateDTab { double2 d[3][4]; };    // dU/dtheta_j, same layout

struct EvalParams {
  const int32_t *k;            // [B][p]
  const qas_pool_entry *pool;  // [c]
  const GateTab *U;            // [NCONST + p*c]
  const GateDTab *dU;          // [NCONST + p*c]
  const double *vtab;          // [S][2^RB][2^(n-RB)] sign tables, register-bit-major (see 3.3)
  const uint32_t *sdesc;       // [S] slice descriptors: x_rest | xh<<20 | imag<<23 | newclass<<24
  int rb0, rb1, rb2;           // ascending positions of the register bits of the H-apply blocking
  double *energy;              // [B]
  double *grad;                // [B][p][l] compact, may be null when !GRAD
  double2 *gstate;             // global-memory statevector scratch (path 1)
  const uint32_t *tape;        // [B][4 + ops_cap*5] compiled tape records (compile_tapes_kernel)
  uint64_t init_basis;
  int B, p, c, l, n, S, ops_cap, init_kind;
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
#ifndef QAS_RB_MAX
#define QAS_RB_MAX 2     // at most 2^2 amplitudes per thread in the register-blocked H application
#endif
__host__ __device__ constexpr int qas_register_bits(int n, int T) {
  int lt = 0;
  while ((1 << lt) < T) ++lt;
  return n - lt >= QAS_RB_MAX ? QAS_RB_MAX : (n - lt > 0 ? n - lt : 0);
}
__device__ __forceinline__ double shfl_xor_d(double v, int off) {
  return __shfl_xor_sync(0xffffffffu, v, off);
}

template <int T>
__device__ __forceinline__ void team_sync(int team) {
  if constexpr (T <= 32) {
    __syncwarp();
  } else if constexpr (T == 256) {
    __syncthreads();
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

__global__ void build_gate_table_kernel(const qas_pool_entry *pool, int c, int p, int l,
                                        const double *params, GateTab *U, GateDTab *dU) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
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

// ------------------------------------------------------------------------------ sign tables
// H is grouped by X mask ("slice"):  (H psi)[i] = sum_s (i if imag) V_s[i] psi[i ^ x_s],
// V_s[i] = sum_{t in slice s} c_t * (+-1 from i^{nY}) * (-1)^{popc((i ^ x_s) & z_t)}.
// Blocking of the H application: every thread owns 2^RB amplitudes that differ only in RB
// "register bits" (positions rb0<rb1<rb2, chosen per Hamiltonian on the host so that as many X
// masks as possible differ only inside them).  Slices whose X masks agree outside the register
// bits form a class: the thread loads the 2^RB partner amplitudes ONCE per class and serves
// every slice of the class from registers (partner j ^ xh).  The tables are stored
// register-bit-major, vtab[s][j][u] with i = deposit(u) | spread(j), so that for fixed (s, j)
// consecutive threads read consecutive doubles.
#define QAS_SD_XREST 0xfffffu
#define QAS_SD_XH(d) (((d) >> 20) & 7u)
#define QAS_SD_IMAG (1u << 23)
#define QAS_SD_NEWCLASS (1u << 24)

__host__ __device__ __forceinline__ uint32_t qas_insert_zero(uint32_t t, int b) {
  return ((t >> b) << (b + 1)) | (t & ((1u << b) - 1u));
}
__host__ __device__ __forceinline__ uint32_t qas_deposit_rest(uint32_t u, int RB, int rb0, int rb1, int rb2) {
  if (RB >= 1) u = qas_insert_zero(u, rb0);
  if (RB >= 2) u = qas_insert_zero(u, rb1);
  if (RB >= 3) u = qas_insert_zero(u, rb2);
  return u;
}
__host__ __device__ __forceinline__ uint32_t qas_spread_reg(uint32_t j, int rb0, int rb1, int rb2) {
  return ((j & 1u) << rb0) | (((j >> 1) & 1u) << rb1) | (((j >> 2) & 1u) << rb2);
}

__global__ void build_sign_tables_kernel(int n, int S, int RB, int rb0, int rb1, int rb2,
                                         const uint32_t *sx, const int *sbegin, const uint32_t *tz,
                                         const double *tc, double *vtab) {
  const size_t dim = (size_t)1 << n;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dim * (size_t)S) return;
  const int s = (int)(idx >> n);
  const uint32_t e = (uint32_t)(idx & (dim - 1));
  const uint32_t u = e & ((1u << (n - RB)) - 1u), jj = e >> (n - RB);
  const uint32_t i = qas_deposit_rest(u, RB, rb0, rb1, rb2) | qas_spread_reg(jj, rb0, rb1, rb2);
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
template <int NB>
__device__ __forceinline__ void hacc_dispatch(double2 (&acc)[NB], const double2 (&a)[NB],
                                              const double (&v)[NB], uint32_t xh, bool imag) {
  switch (xh) {
    case 0: hacc<NB, 0>(acc, a, v, imag); break;
    case 1: if constexpr (NB > 1) hacc<NB, 1>(acc, a, v, imag); break;
    case 2: if constexpr (NB > 2) hacc<NB, 2>(acc, a, v, imag); break;
    case 3: if constexpr (NB > 2) hacc<NB, 3>(acc, a, v, imag); break;
    case 4: if constexpr (NB > 4) hacc<NB, 4>(acc, a, v, imag); break;
    case 5: if constexpr (NB > 4) hacc<NB, 5>(acc, a, v, imag); break;
    case 6: if constexpr (NB > 4) hacc<NB, 6>(acc, a, v, imag); break;
    default: if constexpr (NB > 4) hacc<NB, 7>(acc, a, v, imag); break;
  }
}

// ------------------------------------------------------------------------------ tape compile
// One thread per circuit: runs the tape compiler (tape.h) and writes a record
//   [nops, status, init_index, n1q | ops[ops_cap][5]]   (uint32)
// to global memory.  Doing this in its own launch (B-way parallel) keeps the serial GF(2)
// frame walk out of the evaluation teams' critical path; the records stay L2-resident.
#define QAS_REC_HDR 4
__global__ void compile_tapes_kernel(const int32_t *k, const qas_pool_entry *pool, int B, int p, int c,
                                     int n, int ops_cap, int init_kind, uint64_t init_basis,
                                     uint32_t *tape) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint32_t col[32], row[32];
  uint32_t *rec = tape + (size_t)b * (QAS_REC_HDR + (size_t)ops_cap * QAS_OP_WORDS);
  QasTapeSink sink{rec + QAS_REC_HDR, ops_cap, 0, 0};
  const int st = qas_compile_circuit(k + (size_t)b * p, p, pool, c, n, col, row, sink);
  rec[0] = st ? 0u : (uint32_t)sink.n;
  rec[1] = (uint32_t)st;
  rec[2] = (init_kind == QAS_INIT_BASIS && !st) ? qas_frame_map_basis(col, n, init_basis) : 0u;
  rec[3] = 0u;
}

// ------------------------------------------------------------------------------ evaluation
// shared-memory layout per team (bytes, all 16-aligned):
//   psi [DIM] double2 | lam [DIM] double2 (GRAD) | rec [4 + ops_cap*5] u32
//   | Us [ops_cap][4] double2 (STAGE_U) | red [2][WPT][8] f64 | ered [WPT] f64
__host__ __device__ constexpr bool qas_stage_u(int N) { return N == 0 || N >= 9; }
__host__ __device__ inline size_t eval_team_smem_bytes(int n, int T, bool grad, int ops_cap,
                                                       bool state_in_smem, bool stage_u) {
  const size_t dim = (size_t)1 << n;
  const int wpt = T >= 32 ? T / 32 : 1;
  size_t b = 0;
  if (state_in_smem) b += dim * 16 * (grad ? 2 : 1);
  b += (((size_t)(QAS_REC_HDR + ops_cap * QAS_OP_WORDS) * 4) + 15) & ~(size_t)15;
  if (stage_u) b += (size_t)ops_cap * 64;
  b += (size_t)2 * wpt * 8 * 8;
  b += (((size_t)wpt * 8) + 15) & ~(size_t)15;
  return b;
}

// N > 0: compile-time qubit count, state in shared memory.  N == 0: run-time qubit count,
// state in global memory (one team per CTA, L2/HBM resident) -- same code path otherwise.
#ifndef QAS_MIN_CTAS
#define QAS_MIN_CTAS 3
#endif
template <int N, int T, bool GRAD>
__global__ void __launch_bounds__(256, QAS_MIN_CTAS) eval_kernel(const EvalParams P) {
  constexpr int TEAMS = 256 / T;
  constexpr int WPT = T >= 32 ? T / 32 : 1;      // warps per team
  constexpr int TW = T >= 32 ? 32 : T;           // lanes of one warp that belong to the team
  constexpr int RB = qas_register_bits(N == 0 ? 20 : N, T);   // register bits of the H blocking
  constexpr bool STAGE_U = qas_stage_u(N);
  const int n = N ? N : P.n;
  const uint32_t DIM = 1u << n, HALF = DIM >> 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int team = threadIdx.x / T, tid = threadIdx.x % T;
  const int lane = threadIdx.x & 31, warp_in_team = tid >> 5;
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
  const int rec_words = QAS_REC_HDR + P.ops_cap * QAS_OP_WORDS;
  uint32_t *rec = reinterpret_cast<uint32_t *>(base);
  uint32_t *ops = rec + QAS_REC_HDR;
  base += (((size_t)rec_words * 4) + 15) & ~(size_t)15;
  double2 *Us = reinterpret_cast<double2 *>(base);
  if constexpr (STAGE_U) base += (size_t)P.ops_cap * 64;
  double *red = reinterpret_cast<double *>(base);
  base += (size_t)2 * WPT * 8 * 8;
  double *ered = reinterpret_cast<double *>(base);
  // slice descriptors, shared by the CTA's teams
  uint32_t *sdesc = reinterpret_cast<uint32_t *>(smem_raw + (size_t)TEAMS * team_bytes);
  for (int s = threadIdx.x; s < P.S; s += 256) sdesc[s] = P.sdesc[s];
  __syncthreads();

  for (int b0 = blockIdx.x * TEAMS; b0 < P.B; b0 += gridDim.x * TEAMS) {
    const int b = b0 + team;
    const bool active = b < P.B;
    if constexpr (T > 32) { if (!active) continue; }   // whole warps: nobody else waits on them

    // ---- 1. fetch the compiled tape record, zero the state / the gradient row
    {
      const uint32_t *grec = P.tape + (size_t)(active ? b : 0) * rec_words;
      for (int w = tid; w < rec_words; w += T) rec[w] = active ? grec[w] : 0u;
      const double amp0 = (P.init_kind == QAS_INIT_PLUS) ? exp2(-0.5 * (double)n) : 0.0;
      for (uint32_t i = tid; i < DIM; i += T) psi[i] = make_double2(amp0, 0.0);
      if (GRAD && active)
        for (int i = tid; i < P.p * P.l; i += T) P.grad[(size_t)b * P.p * P.l + i] = 0.0;
    }
    team_sync<T>(team);
    const int nops = (int)rec[0];
    const int status = (int)rec[1];
    if (tid == 0 && P.init_kind != QAS_INIT_PLUS) psi[rec[2]] = make_double2(1.0, 0.0);
    if constexpr (STAGE_U) {   // gate matrices of this circuit -> shared memory (one round trip)
      for (int w = tid; w < nops * 4; w += T) Us[w] = P.U[ops[(w >> 2) * QAS_OP_WORDS + 3]].u[w & 3];
    }
    int nloop = nops;
    if constexpr (T < 32) nloop = __reduce_max_sync(0xffffffffu, nops);
    team_sync<T>(team);

    // ---- 2. forward sweep (tape is stored last-gate-first)
    for (int it = 0; it < nloop; ++it) {
      const int j = nops - 1 - it;
      if (j >= 0) {
        const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
        const uint32_t m = o[0], r = o[1], rc = o[2];
        const double2 *g = STAGE_U ? (Us + 4 * j) : P.U[o[3]].u;
        if (qas_meta_kind(o[4]) == QAS_K_U1) {
          const double2 u00 = g[0], u01 = g[1], u10 = g[2], u11 = g[3];
          const int pb = 31 - __clz(r);
#pragma unroll 4
          for (uint32_t t = tid; t < HALF; t += T) {
            uint32_t p0 = insert_zero(t, pb);
            p0 |= (uint32_t)(__popc(p0 & r) & 1) << pb;
            if (rc && !(__popc(p0 & rc) & 1)) continue;
            const uint32_t p1 = p0 ^ m;
            const double2 a0 = psi[p0], a1 = psi[p1];
            psi[p0] = cfma(u01, a1, cmul(u00, a0));
            psi[p1] = cfma(u11, a1, cmul(u10, a0));
          }
        } else {
          const double2 d0 = g[0], d1 = g[3];
#pragma unroll 4
          for (uint32_t i = tid; i < DIM; i += T)
            psi[i] = cmul((__popc(i & r) & 1) ? d1 : d0, psi[i]);
        }
      }
      team_sync<T>(team);
    }

    // ---- 3. lambda = H psi, E = Re <psi|lambda>   (register-blocked over X-mask classes)
    double e_part = 0.0;
    {
      constexpr int NB = 1 << RB;
      const uint32_t nrest = DIM >> RB;
      uint32_t roff[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) roff[j] = qas_spread_reg((uint32_t)j, P.rb0, P.rb1, P.rb2);
      for (uint32_t u = tid; u < nrest; u += T) {
        const uint32_t base_i = qas_deposit_rest(u, RB, P.rb0, P.rb1, P.rb2);
        double2 acc[NB], a[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) { acc[j] = make_double2(0.0, 0.0); a[j] = acc[j]; }
        // sign-table values are prefetched one slice ahead (L2 latency ~300 cycles)
        const double *vp = P.vtab + u;
        double vn[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) vn[j] = __ldg(vp + (size_t)j * nrest);
        for (int s = 0; s < P.S; ++s) {
          const uint32_t d = sdesc[s];
          double v[NB];
#pragma unroll
          for (int j = 0; j < NB; ++j) v[j] = vn[j];
          if (s + 1 < P.S) {
            const double *vq = vp + ((size_t)(s + 1) << n);
#pragma unroll
            for (int j = 0; j < NB; ++j) vn[j] = __ldg(vq + (size_t)j * nrest);
          }
          if (d & QAS_SD_NEWCLASS) {
            const uint32_t pb = base_i ^ (d & QAS_SD_XREST);
#pragma unroll
            for (int j = 0; j < NB; ++j) a[j] = psi[pb | roff[j]];
          }
          hacc_dispatch<NB>(acc, a, v, QAS_SD_XH(d), (d & QAS_SD_IMAG) != 0);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const uint32_t i = base_i | roff[j];
          const double2 ai = psi[i];
          e_part = fma(ai.x, acc[j].x, fma(ai.y, acc[j].y, e_part));
          if constexpr (GRAD) lam[i] = acc[j];
        }
      }
    }
#pragma unroll
    for (int off = TW / 2; off > 0; off >>= 1) e_part += shfl_xor_d(e_part, off);
    if constexpr (WPT > 1) {
      if (lane == 0) ered[warp_in_team] = e_part;
      team_sync<T>(team);
      if (tid == 0) {
        double e = 0.0;
#pragma unroll
        for (int w = 0; w < WPT; ++w) e += ered[w];
        e_part = e;
      }
    } else if constexpr (GRAD) {
      team_sync<T>(team);      // lambda visible to the whole team
    }
    if (tid == 0 && active) P.energy[b] = status ? __longlong_as_double(0x7ff8000000000000ll) : e_part;

    // ---- 4. adjoint sweep (tape order = reverse time order)
    if constexpr (GRAD) {
      for (int it = 0; it < nloop; ++it) {
        const bool live = it < nops;
        const uint32_t *o = ops + (size_t)(live ? it : 0) * QAS_OP_WORDS;
        const uint32_t m = o[0], r = o[1], rc = o[2], tab = o[3], mt = o[4];
        const int npar = live ? qas_meta_npar(mt) : 0;
        // lanes of the team's first warp prefetch dU/dtheta entries: lane l holds (dU_j)_{l%4} with
        // j = l/4 (teams of >= 16 lanes) or j = round (8-lane teams); the load overlaps the pair loop
        constexpr int ROUNDS = T >= 16 ? 1 : 3;
        double2 dpre[ROUNDS];
#pragma unroll
        for (int rr = 0; rr < ROUNDS; ++rr) {
          const int jj = T >= 16 ? (tid >> 2) : rr;
          const bool valid = T >= 16 ? (tid < 4 * npar) : (tid < 4 && rr < npar);
          dpre[rr] = valid ? P.dU[tab].d[jj][tid & 3] : make_double2(0.0, 0.0);
        }
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.0;
        if (live) {
          const double2 *g = STAGE_U ? (Us + 4 * it) : P.U[tab].u;
          if (qas_meta_kind(mt) == QAS_K_U1) {
            const double2 u00 = g[0], u01 = g[1], u10 = g[2], u11 = g[3];
            const int pb = 31 - __clz(r);
#pragma unroll 2
            for (uint32_t t = tid; t < HALF; t += T) {
              uint32_t p0 = insert_zero(t, pb);
              p0 |= (uint32_t)(__popc(p0 & r) & 1) << pb;
              if (rc && !(__popc(p0 & rc) & 1)) continue;
              const uint32_t p1 = p0 ^ m;
              const double2 a0 = psi[p0], a1 = psi[p1];
              const double2 q0 = cfma_conj(u10, a1, cmul_conj(u00, a0));   // U^dagger psi
              const double2 q1 = cfma_conj(u11, a1, cmul_conj(u01, a0));
              psi[p0] = q0;
              psi[p1] = q1;
              const double2 l0 = lam[p0], l1 = lam[p1];
              if (npar) {   // M_ab += conj(l_a) q_b
                acc[0] = fma(l0.x, q0.x, fma(l0.y, q0.y, acc[0])); acc[1] = fma(l0.x, q0.y, fma(-l0.y, q0.x, acc[1]));
                acc[2] = fma(l0.x, q1.x, fma(l0.y, q1.y, acc[2])); acc[3] = fma(l0.x, q1.y, fma(-l0.y, q1.x, acc[3]));
                acc[4] = fma(l1.x, q0.x, fma(l1.y, q0.y, acc[4])); acc[5] = fma(l1.x, q0.y, fma(-l1.y, q0.x, acc[5]));
                acc[6] = fma(l1.x, q1.x, fma(l1.y, q1.y, acc[6])); acc[7] = fma(l1.x, q1.y, fma(-l1.y, q1.x, acc[7]));
              }
              lam[p0] = cfma_conj(u10, l1, cmul_conj(u00, l0));
              lam[p1] = cfma_conj(u11, l1, cmul_conj(u01, l0));
            }
          } else {
            const double2 d0 = g[0], d1 = g[3];
#pragma unroll 2
            for (uint32_t i = tid; i < DIM; i += T) {
              const bool hi = __popc(i & r) & 1;
              const double2 d = hi ? d1 : d0;
              const double2 q = cmul_conj(d, psi[i]);
              const double2 lv = lam[i];
              psi[i] = q;
              if (npar) {
                const double re = fma(lv.x, q.x, lv.y * q.y), im = fma(lv.x, q.y, -lv.y * q.x);
                if (hi) { acc[6] += re; acc[7] += im; } else { acc[0] += re; acc[1] += im; }
              }
              lam[i] = cmul_conj(d, lv);
            }
          }
        }
        // team reduction of the 2x2 cross matrix M (8 doubles): reduce-scatter butterfly
        bool any_par = npar > 0;
        if constexpr (T < 32) any_par = __any_sync(0xffffffffu, any_par);
        if (any_par) {
          const bool h1 = lane & (TW / 2), h2 = lane & (TW / 4), h3 = lane & (TW / 8);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double send = h1 ? acc[q] : acc[q + 4], keep = h1 ? acc[q + 4] : acc[q];
            acc[q] = keep + shfl_xor_d(send, TW / 2);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const double send = h2 ? acc[q] : acc[q + 2], keep = h2 ? acc[q + 2] : acc[q];
            acc[q] = keep + shfl_xor_d(send, TW / 4);
          }
          {
            const double send = h3 ? acc[0] : acc[1], keep = h3 ? acc[1] : acc[0];
            acc[0] = keep + shfl_xor_d(send, TW / 8);
          }
#pragma unroll
          for (int off = TW / 16; off > 0; off >>= 1) acc[0] += shfl_xor_d(acc[0], off);
          if ((lane & (TW / 8 - 1)) == 0 && npar) {
            const int v = (h1 ? 4 : 0) + (h2 ? 2 : 0) + (h3 ? 1 : 0);
            red[((it & 1) * WPT + warp_in_team) * 8 + v] = acc[0];
          }
        }
        team_sync<T>(team);
        // gradient of this gate: lane l = 4j+q contributes Re[(dU_j)_q * M_q]; 4-lane shuffle sum.
        // (for T < 32 every lane of the warp takes part in the shuffles; only lanes tid < 12 count)
        bool fin = npar > 0;
        if constexpr (T < 32) fin = __any_sync(0xffffffffu, fin); else fin = fin && (tid < 32);
        if (fin) {
#pragma unroll
          for (int rr = 0; rr < ROUNDS; ++rr) {
            const int jj = T >= 16 ? (tid >> 2) : rr;
            const bool valid = T >= 16 ? (tid < 4 * npar) : (tid < 4 && rr < npar);
            double part = 0.0;
            if (valid) {
              const int q = tid & 3;
              double mre = 0.0, mim = 0.0;
#pragma unroll
              for (int w = 0; w < WPT; ++w) {
                mre += red[((it & 1) * WPT + w) * 8 + 2 * q];
                mim += red[((it & 1) * WPT + w) * 8 + 2 * q + 1];
              }
              part = dpre[rr].x * mre - dpre[rr].y * mim;
            }
            part += shfl_xor_d(part, 1);
            part += shfl_xor_d(part, 2);
            if (valid && (tid & 3) == 0 && active && !status)
              P.grad[((size_t)b * P.p + qas_meta_depth(mt)) * P.l + jj] = 2.0 * part;
          }
        }
      }
    }
    team_sync<T>(team);   // state buffers are reused by the next circuit of this team
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

__global__ void grad_chunk_reduce_kernel(const int32_t *k, const double *grad, int B, int p, int c,
                                         int l, int chunk, int depth_tile, double *partial) {
  extern __shared__ double acc[];                 // [depth_tile][c][l]
  const int ch = blockIdx.x, i0 = blockIdx.y * depth_tile;
  const int nd = min(depth_tile, p - i0);
  for (int x = threadIdx.x; x < nd * c * l; x += blockDim.x) acc[x] = 0.0;
  __syncthreads();
  const int b_lo = ch * chunk, b_hi = min(B, b_lo + chunk);
  for (int x = threadIdx.x; x < nd * l; x += blockDim.x) {
    const int di = x / l, j = x % l;
    double *a = acc + (size_t)di * c * l + j;
    for (int b = b_lo; b < b_hi; ++b) {
      const int key = k[(size_t)b * p + i0 + di];
      const double gval = grad[((size_t)b * p + i0 + di) * l + j];
      if (key >= 0 && key < c) a[key * l] += nan_to_num_d(gval);
    }
  }
  __syncthreads();
  double *out = partial + ((size_t)ch * p + i0) * c * l;
  for (int x = threadIdx.x; x < nd * c * l; x += blockDim.x) out[x] = acc[x];
}

__global__ void grad_final_reduce_kernel(const double *partial, int nchunks, size_t total,
                                         double scale, double *out) {
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= total) return;
  double s = 0.0;
  for (int ch = 0; ch < nchunks; ++ch) s += partial[(size_t)ch * total + x];
  out[x] = s * scale;
}

// ------------------------------------------------------------------------------ FP64 peak
// 8 independent DFMA chains per thread; 2 flop per DFMA.
__global__ void dfma_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
