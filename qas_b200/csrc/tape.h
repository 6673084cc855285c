// tape.h -- the tape compiler: lowers one structure list k (pool keys) to the generalised-gate
// tape the statevector kernels execute.  Compiled for host AND device: the evaluation kernel
// runs it in its prologue (one compile per circuit, no tape round-trip through HBM) and
// qas_compile_tape() exposes the identical code on the host for tests.
//
// Restates backboneCirc (reference qas/qml_models/hamiltonian_model.py:22-41): walk k, skip
// PlaceHolder, emit one operator per pool key.  What is new here:
//
//  * CNOT / SWAP folding.  A CNOT is the basis permutation |i> -> |C i> with C linear over
//    GF(2).  Instead of moving amplitudes, the compiler keeps the frame S_j = C_N ... C_{j+1}
//    (product of all CNOT maps AFTER position j) and stores the vector in the END-of-circuit
//    frame: phi_j[S_j i] = psi_j[i].  CNOTs then cost nothing; a 1-qubit gate on logical
//    qubit q becomes a "generalised" gate that pairs physical indices (P, P ^ m) with
//    m = S e_q (column q of S) and reads the logical bit as parity(P & r), r = row q of S^-1.
//    The frame at the end is the identity, so the Hamiltonian is applied in plain logical
//    coordinates, and basis / uniform initial states stay basis / uniform.
//  * every gate of the pool vocabulary lowers to two primitive kinds:
//      K_U1    2x2 unitary on the pair (P, P^m), optionally controlled by parity(P & rc)
//      K_PDIAG diagonal phase d[parity(P & r)]   (IsingZZ; IsingXX/YY after basis change)
//
// The walk is backwards in time, so ops are emitted last-gate-first; the kernels run the
// forward pass from the end of the tape and the adjoint sweep from its start.
#pragma once
#include <stdint.h>
#include "../../include/qas_b200.h"

#if defined(__CUDACC__)
#define QAS_HD __host__ __device__ __forceinline__
#else
#define QAS_HD inline
#endif

#define QAS_MAX_QUBITS 20
#define QAS_OP_WORDS 5      // {m, r, rc, tab, meta}
#define QAS_K_U1 0u
#define QAS_K_PDIAG 1u

// constant 2x2 table entries (device table rows [0, QAS_NCONST))
enum {
  QAS_C_H = 0, QAS_C_X, QAS_C_Y, QAS_C_Z, QAS_C_S, QAS_C_SDG, QAS_C_T, QAS_C_TDG, QAS_C_SX,
  QAS_C_SH, QAS_C_HSDG, QAS_C_ZZ_MPI4, QAS_NCONST
};

// meta word: depth (bits 0-15) | kind (16-19) | nparams (20-21)
QAS_HD uint32_t qas_meta(int depth, uint32_t kind, int npar) {
  return (uint32_t)depth | (kind << 16) | ((uint32_t)npar << 20);
}
QAS_HD int qas_meta_depth(uint32_t m) { return (int)(m & 0xffffu); }
QAS_HD uint32_t qas_meta_kind(uint32_t m) { return (m >> 16) & 0xfu; }
QAS_HD int qas_meta_npar(uint32_t m) { return (int)((m >> 20) & 0x3u); }

// number of parameters of a gate id (SUPPORTED_OPS_DICT[name].num_params upstream)
QAS_HD int qas_gate_nparams(int g) {
  switch (g) {
    case QAS_G_ROT: case QAS_G_U3: case QAS_G_CROT: return 3;
    case QAS_G_U2: return 2;
    case QAS_G_RX: case QAS_G_RY: case QAS_G_RZ: case QAS_G_PHASESHIFT: case QAS_G_U1:
    case QAS_G_CRX: case QAS_G_CRY: case QAS_G_CRZ: case QAS_G_CPHASE:
    case QAS_G_ISINGXX: case QAS_G_ISINGYY: case QAS_G_ISINGZZ: return 1;
    default: return 0;
  }
}
QAS_HD int qas_gate_nwires(int g) { return g >= QAS_G_CNOT ? 2 : 1; }
// worst-case number of tape ops one pool key can expand to
QAS_HD int qas_gate_expansion(int g) {
  switch (g) {
    case QAS_G_PLACEHOLDER: case QAS_G_CNOT: case QAS_G_SWAP: return 0;
    case QAS_G_ISINGXX: case QAS_G_ISINGYY: return 5;
    case QAS_G_ISWAP: return 3;
    case QAS_G_SISWAP: return 10;
    default: return 1;
  }
}

// Where compiled ops go.  Any type with fields n / overflow and emit(m, r, rc, tab, meta) works:
// QasTapeSink writes 5-word records; the thread-per-circuit kernel packs ops into shared memory.
struct QasTapeSink {
  uint32_t *ops;   // [cap][QAS_OP_WORDS]
  int cap;
  int n;           // emitted so far
  int overflow;
  QAS_HD void emit(uint32_t m, uint32_t r, uint32_t rc, uint32_t tab, uint32_t meta) {
    if (n >= cap) { overflow = 1; return; }
    uint32_t *o = ops + (size_t)n * QAS_OP_WORDS;
    o[0] = m; o[1] = r; o[2] = rc; o[3] = tab; o[4] = meta;
    n++;
  }
};

template <class Sink>
QAS_HD void qas_emit(Sink &s, uint32_t m, uint32_t r, uint32_t rc, uint32_t tab, uint32_t meta) {
  s.emit(m, r, rc, tab, meta);
}

// Compile one circuit.  col/row: scratch arrays of n words each (frame S and S^-1).
// Returns 0, or 1 if some key is out of range (the reference asserts min(k)>=0, max(k)<c at
// utils.py:12-15), or 2 on tape overflow.
template <class Sink>
QAS_HD int qas_compile_circuit(const int32_t *k, int p, const qas_pool_entry *pool, int c, int n,
                               uint32_t *col, uint32_t *row, Sink &s) {
  for (int q = 0; q < n; ++q) col[q] = row[q] = 1u << (n - 1 - q);
  s.n = 0;
  s.overflow = 0;
  for (int i = p - 1; i >= 0; --i) {
    const int key = k[i];
    if (key < 0 || key >= c) return 1;
    const int g = pool[key].gate_id;
    const int a = pool[key].wire0, b = pool[key].wire1;
    const uint32_t ptab = (uint32_t)QAS_NCONST + (uint32_t)i * (uint32_t)c + (uint32_t)key;
    const int np = qas_gate_nparams(g);
    if (g <= QAS_G_CNOT) {
      // PlaceHolder, every one-qubit gate and CNOT -- the whole vocabulary of the near-neighbour-CNOT pools the search
      // scripts use -- on ONE predicated path: the threads of a warp compile different circuits, and the three-way
      // branch per depth step was the main divergence of the compile kernels.  Same emits / frame updates as the
      // cases below (a one-qubit gate or a PlaceHolder leaves the frame as it is: bb == a, xor with zero).
      const bool cx = g == QAS_G_CNOT;
      const int bb = cx ? b : a;
      const uint32_t ca = col[a], ra = row[a], cb = col[bb], rb = row[bb];
      if (g != QAS_G_PLACEHOLDER && !cx)
        qas_emit(s, ca, ra, 0u, g < QAS_G_ROT ? (uint32_t)(g - QAS_G_HADAMARD) : ptab, qas_meta(i, QAS_K_U1, np));
      col[a] = ca ^ (cx ? cb : 0u);
      row[bb] = rb ^ (cx ? ra : 0u);
      continue;
    }
    switch (g) {
      case QAS_G_PLACEHOLDER: break;
      case QAS_G_HADAMARD: case QAS_G_PAULIX: case QAS_G_PAULIY: case QAS_G_PAULIZ:
      case QAS_G_S: case QAS_G_SDG: case QAS_G_T: case QAS_G_TDG: case QAS_G_SX:
        // constant table rows are laid out in the same order as the gate ids 1..9
        qas_emit(s, col[a], row[a], 0u, (uint32_t)(g - QAS_G_HADAMARD), qas_meta(i, QAS_K_U1, 0));
        break;
      case QAS_G_ROT: case QAS_G_RX: case QAS_G_RY: case QAS_G_RZ: case QAS_G_PHASESHIFT:
      case QAS_G_U1: case QAS_G_U2: case QAS_G_U3:
        qas_emit(s, col[a], row[a], 0u, ptab, qas_meta(i, QAS_K_U1, np));
        break;
      case QAS_G_CNOT:      // fold: S <- S o C,  S^-1 <- C S^-1   (control a, target b)
        col[a] ^= col[b];
        row[b] ^= row[a];
        break;
      case QAS_G_SWAP: {
        uint32_t t = col[a]; col[a] = col[b]; col[b] = t;
        t = row[a]; row[a] = row[b]; row[b] = t;
        break;
      }
      case QAS_G_CZ:
        qas_emit(s, col[b], row[b], row[a], QAS_C_Z, qas_meta(i, QAS_K_U1, 0));
        break;
      case QAS_G_CY:
        qas_emit(s, col[b], row[b], row[a], QAS_C_Y, qas_meta(i, QAS_K_U1, 0));
        break;
      case QAS_G_CRX: case QAS_G_CRY: case QAS_G_CRZ: case QAS_G_CROT: case QAS_G_CPHASE:
        qas_emit(s, col[b], row[b], row[a], ptab, qas_meta(i, QAS_K_U1, np));
        break;
      case QAS_G_ISINGZZ:
        qas_emit(s, 0u, row[a] ^ row[b], 0u, ptab, qas_meta(i, QAS_K_PDIAG, 1));
        break;
      case QAS_G_ISINGXX:   // (H x H) ZZ(theta) (H x H); emitted last-first
        qas_emit(s, col[a], row[a], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, 0u, row[a] ^ row[b], 0u, ptab, qas_meta(i, QAS_K_PDIAG, 1));
        qas_emit(s, col[a], row[a], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        break;
      case QAS_G_ISINGYY:   // (SH x SH) ZZ(theta) (HS^ x HS^)
        qas_emit(s, col[a], row[a], 0u, QAS_C_SH, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_SH, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, 0u, row[a] ^ row[b], 0u, ptab, qas_meta(i, QAS_K_PDIAG, 1));
        qas_emit(s, col[a], row[a], 0u, QAS_C_HSDG, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_HSDG, qas_meta(i, QAS_K_U1, 0));
        break;
      case QAS_G_ISWAP: {   // (S x S) CZ SWAP  (SWAP acts first, so it is folded last)
        qas_emit(s, col[a], row[a], 0u, QAS_C_S, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_S, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], row[a], QAS_C_Z, qas_meta(i, QAS_K_U1, 0));
        uint32_t t = col[a]; col[a] = col[b]; col[b] = t;
        t = row[a]; row[a] = row[b]; row[b] = t;
        break;
      }
      case QAS_G_SISWAP:    // exp(i pi/8 (XX+YY)) = IsingXX(-pi/4) IsingYY(-pi/4)
        qas_emit(s, col[a], row[a], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, 0u, row[a] ^ row[b], 0u, QAS_C_ZZ_MPI4, qas_meta(i, QAS_K_PDIAG, 0));
        qas_emit(s, col[a], row[a], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_H, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[a], row[a], 0u, QAS_C_SH, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_SH, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, 0u, row[a] ^ row[b], 0u, QAS_C_ZZ_MPI4, qas_meta(i, QAS_K_PDIAG, 0));
        qas_emit(s, col[a], row[a], 0u, QAS_C_HSDG, qas_meta(i, QAS_K_U1, 0));
        qas_emit(s, col[b], row[b], 0u, QAS_C_HSDG, qas_meta(i, QAS_K_U1, 0));
        break;
      default: return 1;
    }
  }
  return s.overflow ? 2 : 0;
}

// ---- pass planner ----------------------------------------------------------------------------
// (1) Grouping.  Consecutive uncontrolled K_U1 ops (tape indices j .. j+g-1, g <= gmax <= 3) share
// ONE pass over the state when every thread can own the 2^g amplitudes {P ^ span(m_0..m_{g-1})}
// and apply all g gates in registers.  Condition: no cross terms, parity(m_a & r_b) = 0 for a != b
// (with parity(m_a & r_a) = 1, which always holds, the m's and the r's are linearly independent and
// the register index bit c_a of slot P ^ (c_0 m_0 ^ c_1 m_1 ^ c_2 m_2) equals the logical bit of
// gate a whenever r_a.P = 0 for all a).  Greedy in tape order; recorded in the meta words:
// bits 24-25 = g on the FIRST op of a group (adjoint sweep walks up), bits 26-27 = g on the LAST
// op (forward sweep walks down).
//
// (2) Support tracking.  Starting from a basis state |b0>, the state after some gates is non-zero
// only on the affine subspace b0 ^ D, D = span of the xor masks of the gates applied so far.  The
// planner walks the groups in TIME order, keeps D, and gives every group a descriptor
//     gd = { P0, dc, X_0 .. X_{n-g-1} }
// (in slot coordinates when `swizzled`, like the ops' xor masks after planning)
// such that the coset representatives of the group are  P(t) = P0 ^ xor_{i: bit i of t} X_i,
// t = 0 .. 2^(n-g)-1, every P(t) lies in the common kernel of the group's r's (so slot index bits
// are logical bits), and the cosets that can hold non-zero amplitudes AFTER the group are exactly
// t < 2^dc.  The forward sweep visits only those; the adjoint sweep visits all cosets for lambda
// but touches psi and the cross matrices only for t < 2^dc.  For shallow circuits from a basis
// state (the QAS search regime) this removes most of the forward work and ~2/3 of the adjoint's.
// The X_i are put in lowest-set-bit echelon form of their (swizzled) shared-memory slot index and
// sorted by pivot, so that the three lane bits of a quarter-warp get independent low address bits
// whenever the subspace allows it (bank conflicts).  Uniform initial state: D is everything.
// The pivot tables are PACKED into a few 32-bit registers (QasPackTab: 8-bit entries up to 8 qubits, 16-bit up to 16) and
// every reduction is a data-dependent loop over the set bits it actually meets -- a handful of iterations -- instead of
// a fully unrolled sweep over all NQ pivots: a third of the executed instructions for an 8-qubit H2O tape and a code
// footprint that fits the instruction cache (the unrolled tables made compile_tapes_kernel 63 k SASS instructions with
// `no_instruction` as its top stall; tables in local memory, the other way to roll the loops, were slower still).  The
// results are bit-identical to the unrolled formulation (same insertion order, same reductions).
#define QAS_PUNROLL _Pragma("unroll")     // register-resident tables of the c-tape compiler (qas_compress_packed)
#define QAS_META_GFIRST(m) (((m) >> 24) & 3u)
#define QAS_META_GLAST(m) (((m) >> 26) & 3u)
#define QAS_GD_HDR 2        // P0, dc
#define QAS_GD_STRIDE(n) (QAS_GD_HDR + (n))

QAS_HD int qas_parity32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v) & 1;
#else
  return __builtin_popcount(v) & 1;
#endif
}
QAS_HD int qas_msb32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz(v);
#else
  return 31 - __builtin_clz(v);
#endif
}
QAS_HD int qas_lsb32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs(v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
// slot index of amplitude i in the swizzled shared-memory statevector: every 3-bit group of the
// index above bit 2 is xor-folded onto bits 0..2, so index bit p also moves bank-group bit p mod 3
// (xor-linear and an involution: the fold only reads bits >= 3, which it leaves unchanged)
QAS_HD uint32_t qas_swz(uint32_t i) {
  uint32_t f = i >> 3;
  f ^= f >> 3;
  f ^= f >> 6;
  f ^= f >> 12;
  return i ^ (f & 7u);
}

// Subspaces of GF(2)^n are kept as pivot-indexed tables, entry p = the basis vector whose pivot is bit p (0 = none),
// packed EB bits per entry into 32-bit words that stay in registers: a run-time index costs a select chain over the
// W words and one shift, never a local-memory access (one thread plans one circuit).
template <int NQ>
struct QasPackTab {
  static constexpr int EB = NQ <= 8 ? 8 : (NQ <= 16 ? 16 : 32);
  static constexpr int PER = 32 / EB;
  static constexpr int W = (NQ + PER - 1) / PER;
  static constexpr uint32_t EM = EB == 32 ? 0xffffffffu : ((1u << (EB & 31)) - 1u);
  uint32_t w[W];
  QAS_HD void clear() {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < W; ++i) w[i] = 0u;
  }
  QAS_HD uint32_t get(int p) const {
    const int wi = p / PER;
    uint32_t x = w[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < W; ++i) x = (wi == i) ? w[i] : x;
    return (x >> ((p % PER) * EB)) & EM;
  }
  QAS_HD void toggle(int p, uint32_t v) {          // entry p ^= v  (v = the whole vector for an empty entry)
    const int wi = p / PER;
    const uint32_t s = v << ((p % PER) * EB);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < W; ++i) w[i] ^= (wi == i) ? s : 0u;
  }
  QAS_HD int rank(int n) const {
    int r = 0;
    for (int p = 0; p < n; ++p) r += get(p) ? 1 : 0;
    return r;
  }
};
// highest-set-bit pivots (rank / membership)
template <int NQ>
QAS_HD bool qas_tab_add_msb(QasPackTab<NQ> &t, uint32_t x) {
  bool placed = false;
  while (x) {
    const int p = qas_msb32(x);
    const uint32_t v = t.get(p);
    if (!v) { t.toggle(p, x); placed = true; break; }
    x ^= v;
  }
  return placed;
}
// lowest-set-bit pivots: x is reduced against `all` (the union of the tables built so far: their pivots are disjoint)
// and, if something is left, stored in `dst` and in `all`.  Reducing at pivot p only changes bits above p.
template <int NQ>
QAS_HD bool qas_tab_add_lsb(QasPackTab<NQ> &all, QasPackTab<NQ> &dst, bool dst_is_all, uint32_t x) {
  bool placed = false;
  while (x) {
    const int p = qas_lsb32(x);
    const uint32_t v = all.get(p);
    if (!v) {
      all.toggle(p, x);
      if (!dst_is_all) dst.toggle(p, x);
      placed = true;
      break;
    }
    x ^= v;
  }
  return placed;
}

// (3) Support of the FINAL state, for the H application.  The evaluation kernel blocks lambda = H psi
// over cosets of a register subspace W (kernels.cuh, "sign tables"): coset u (n - RB bits, the index
// with W's pivot bits removed) receives from X-mask class rho only if psi is non-zero somewhere on
// coset u ^ rho, i.e. iff  u ^ rho ^ ubar(b0)  lies in Dbar = image of D in the quotient by W.  The
// planner hands the kernel a basis of the quotient space adapted to Dbar,
//     hd = { ubar(b0), dbar, X_0 .. X_{nu-1}, chk_0 .. chk_{nu-dbar-1} }          (nu = n - RB)
// X_0..X_{dbar-1} = reduced echelon basis of Dbar (highest-bit pivots, ascending: X_i only touches
// bits <= its pivot, so the low counter bits -- the lanes -- stay on neighbouring table entries),
// X_dbar.. = the unit vectors of the non-pivot positions.  Enumerating u(t) = ubar(b0) ^ xor of the
// X_i selected by t, the Dbar-coset of u is t >> dbar, and the Dbar-coset of a class representative
// is read off its parities with the check masks chk_i.  dbar = nu means "dense: no restriction".
#define QAS_HD_WORDS 32
struct QasHInfo {
  int RB, enable;
  int rb[3];           // ascending pivots of W
  uint32_t wv[3];      // basis of W in reduced echelon form (wv[q] has pivot rb[q])
};
QAS_HD uint32_t qas_ubar(const QasHInfo &hi, uint32_t x) {
  // compile-time trip counts: the descriptor is a kernel argument and must not be indexed at run time (local memory)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 3; ++q) if (q < hi.RB && ((x >> hi.rb[q]) & 1u)) x ^= hi.wv[q];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 2; q >= 0; --q) if (q < hi.RB) x = ((x >> (hi.rb[q] + 1)) << hi.rb[q]) | (x & ((1u << hi.rb[q]) - 1u));
  return x;
}

template <int NQ>
QAS_HD void qas_plan_supports(const uint32_t *ops, int nops, int ngroups, int n, int init_kind,
                              uint32_t init_index, uint32_t *gd, int swizzled,
                              const QasHInfo *hi = nullptr, uint32_t *hd = nullptr) {
  QasPackTab<NQ> D;
  D.clear();
  int drank = 0;
  if (init_kind == QAS_INIT_PLUS) {
    for (int p = 0; p < n; ++p) D.toggle(p, 1u << p);
    drank = n;
  }
  const uint32_t b0 = (init_kind == QAS_INIT_BASIS) ? init_index : 0u;
  const int stride = QAS_GD_STRIDE(n);
  int e = nops - 1;
  for (int gi = ngroups - 1; gi >= 0; --gi) {          // groups in time order
    const int g = (int)QAS_META_GLAST(ops[(size_t)e * QAS_OP_WORDS + 4]);
    const int j = e - g + 1;
    const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
    uint32_t *d = gd + (size_t)gi * stride;
    e = j - 1;
    if (qas_meta_kind(o[4]) != QAS_K_U1) {       // parity-diagonal: dense, support unchanged
      d[0] = 0u; d[1] = (uint32_t)n;
      for (int i = 0; i < n; ++i) d[QAS_GD_HDR + i] = 1u << i;
      continue;
    }
    uint32_t m0 = o[0], r0 = o[1], m1 = 0u, r1 = 0u, m2 = 0u, r2 = 0u;
    if (g > 1) { m1 = o[QAS_OP_WORDS]; r1 = o[QAS_OP_WORDS + 1]; }
    if (g > 2) { m2 = o[2 * QAS_OP_WORDS]; r2 = o[2 * QAS_OP_WORDS + 1]; }
    // projection onto the common kernel of the r's along span(m's), then to slot coordinates
    auto proj = [&](uint32_t x) {
      uint32_t y = x;
      if (qas_parity32(r0 & x)) y ^= m0;
      if (qas_parity32(r1 & x)) y ^= m1;
      if (qas_parity32(r2 & x)) y ^= m2;
      return swizzled ? qas_swz(y) : y;
    };
    auto proj_unit = [&](int p) {                 // proj(1 << p): the parities are single bits of the r's
      uint32_t y = 1u << p;
      if ((r0 >> p) & 1u) y ^= m0;
      if ((r1 >> p) & 1u) y ^= m1;
      if ((r2 >> p) & 1u) y ^= m2;
      return swizzled ? qas_swz(y) : y;
    };
    QasPackTab<NQ> Sall, Sin;                     // Sall = Sin + the completion vectors (disjoint pivots)
    Sall.clear();
    int dc = 0, tot = 0;
    if (drank == n) {     // the state is dense already: every coset is inside, one echelon pass
      for (int p = 0; p < n; ++p) dc += qas_tab_add_lsb<NQ>(Sall, Sall, true, proj_unit(p)) ? 1 : 0;
      tot = dc;
      d[0] = proj(b0);
      d[1] = (uint32_t)dc;
      int cnt = 0;
      for (int p = 0; p < n; ++p) { const uint32_t v = Sall.get(p); if (v) d[QAS_GD_HDR + cnt++] = v; }
    } else {
      for (int p = 0; p < n; ++p) {
        const uint32_t v = D.get(p);
        if (v) dc += qas_tab_add_lsb<NQ>(Sall, Sall, true, proj(v)) ? 1 : 0;
      }
      Sin = Sall;
      tot = dc;
      for (int p = 0; p < n && tot < n - g; ++p) tot += qas_tab_add_lsb<NQ>(Sall, Sall, true, proj_unit(p)) ? 1 : 0;
      d[0] = proj(b0);
      d[1] = (uint32_t)dc;
      int cnt = 0;
      for (int p = 0; p < n; ++p) { const uint32_t v = Sin.get(p); if (v) d[QAS_GD_HDR + cnt++] = v; }
      for (int p = 0; p < n; ++p) { const uint32_t v = Sall.get(p); if (v && !Sin.get(p)) d[QAS_GD_HDR + cnt++] = v; }
      drank += qas_tab_add_msb<NQ>(D, m0) ? 1 : 0;
      if (g > 1) drank += qas_tab_add_msb<NQ>(D, m1) ? 1 : 0;
      if (g > 2) drank += qas_tab_add_msb<NQ>(D, m2) ? 1 : 0;
    }
  }
  if (!hd) return;
  const int nu = n - (hi ? hi->RB : 0);
  hd[0] = 0u;
  hd[1] = (uint32_t)nu;                           // dense unless shown otherwise
  if (!hi || !hi->enable || nu > NQ || 2 + 2 * nu > QAS_HD_WORDS) return;
  if (drank >= n) return;                         // the final state is dense
  QasPackTab<NQ> E;
  E.clear();
  int dbar = 0;
  for (int p = 0; p < n; ++p) {
    const uint32_t v = D.get(p);
    if (v) dbar += qas_tab_add_msb<NQ>(E, qas_ubar(*hi, v)) ? 1 : 0;
  }
  if (dbar >= nu) return;
  // reduced echelon form: clear every pivot bit from the other basis vectors (ascending pivots, so the
  // vector xor-ed in is already free of lower pivot bits)
  for (int p = 0; p < nu; ++p) {
    const uint32_t v = E.get(p);
    if (!v) continue;
    for (int p2 = p + 1; p2 < nu; ++p2)
      if ((E.get(p2) >> p) & 1u) E.toggle(p2, v);
  }
  hd[0] = qas_ubar(*hi, b0);
  hd[1] = (uint32_t)dbar;
  int cnt = 0;
  for (int p = 0; p < nu; ++p) { const uint32_t v = E.get(p); if (v) hd[2 + cnt++] = v; }
  for (int p = 0; p < nu; ++p) if (!E.get(p)) hd[2 + cnt++] = 1u << p;
  int kc = 0;
  for (int f = 0; f < nu; ++f)
    if (!E.get(f)) {
      uint32_t chk = 1u << f;
      for (int p = 0; p < nu; ++p) if ((E.get(p) >> f) & 1u) chk |= 1u << p;
      hd[2 + nu + kc++] = chk;
    }
}

// Returns the number of groups.  gd receives one descriptor of QAS_GD_STRIDE(n) words per group,
// in TAPE order of the groups (may be null: grouping only).  swizzled: the statevector lives in
// shared memory; descriptors and the ops' xor masks are then written in slot coordinates.
// grouping only: marks the meta words, returns the number of groups
QAS_HD int qas_mark_groups(uint32_t *ops, int nops, int gmax) {
  int ngroups = 0;
  for (int j = 0; j < nops;) {
    uint32_t *a = ops + (size_t)j * QAS_OP_WORDS;
    int g = 1;
    if (qas_meta_kind(a[4]) == QAS_K_U1 && a[2] == 0u) {
      while (g < gmax && j + g < nops) {
        const uint32_t *b = ops + (size_t)(j + g) * QAS_OP_WORDS;
        if (qas_meta_kind(b[4]) != QAS_K_U1 || b[2] != 0u) break;
        bool ok = true;
        for (int q = 0; q < g; ++q) {
          const uint32_t *e = ops + (size_t)(j + q) * QAS_OP_WORDS;
          if (qas_parity32(e[0] & b[1]) || qas_parity32(b[0] & e[1])) ok = false;
        }
        if (!ok) break;
        ++g;
      }
    }
    a[4] |= (uint32_t)g << 24;
    ops[(size_t)(j + g - 1) * QAS_OP_WORDS + 4] |= (uint32_t)g << 26;
    j += g;
    ++ngroups;
  }
  return ngroups;
}
// support descriptors of an already grouped tape (+ the ops' xor masks to slot coordinates when swizzled)
QAS_HD void qas_plan_marked(uint32_t *ops, int nops, int ngroups, int n, int init_kind, uint32_t init_index,
                            uint32_t *gd, int swizzled, const QasHInfo *hi = nullptr, uint32_t *hd = nullptr) {
  if (n <= 5) qas_plan_supports<5>(ops, nops, ngroups, n, init_kind, init_index, gd, swizzled, hi, hd);
  else if (n <= 8) qas_plan_supports<8>(ops, nops, ngroups, n, init_kind, init_index, gd, swizzled, hi, hd);
  else if (n <= 10) qas_plan_supports<10>(ops, nops, ngroups, n, init_kind, init_index, gd, swizzled, hi, hd);
  else if (n <= 12) qas_plan_supports<12>(ops, nops, ngroups, n, init_kind, init_index, gd, swizzled, hi, hd);
  else qas_plan_supports<QAS_MAX_QUBITS>(ops, nops, ngroups, n, init_kind, init_index, gd, swizzled, hi, hd);
  if (swizzled)
    for (int j = 0; j < nops; ++j) {
      uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
      if (qas_meta_kind(o[4]) == QAS_K_U1) o[0] = qas_swz(o[0]);
    }
}
QAS_HD int qas_plan_groups(uint32_t *ops, int nops, int gmax, int n, int init_kind, uint32_t init_index,
                           uint32_t *gd, int swizzled, const QasHInfo *hi = nullptr, uint32_t *hd = nullptr) {
  const int ngroups = qas_mark_groups(ops, nops, gmax);
  if (gd) qas_plan_marked(ops, nops, ngroups, n, init_kind, init_index, gd, swizzled, hi, hd);
  return ngroups;
}

// physical index of a logical basis state in the end-of-circuit frame: S_0 * basis
QAS_HD uint32_t qas_frame_map_basis(const uint32_t *col, int n, uint64_t basis) {
  uint32_t out = 0;
  for (int q = 0; q < n; ++q)
    if ((basis >> (n - 1 - q)) & 1ull) out ^= col[q];
  return out;
}

// ---- support-compressed tapes ("c-tapes") -----------------------------------------------------
// From the all-zero state a circuit only populates D = span of the xor masks of its 2x2 ops, and
// the support after the first j ops (time order) is the span D_j of the first j masks.  The c-tape
// expresses the circuit in a TIME-ADAPTED basis of D: B_k = mask of the k-th op (time order) that
// enlarges the span.  With i(t) = xor_k t_k B_k the support after j ops is simply {t < 2^{d_j}} --
// a contiguous prefix of a 2^d-amplitude register -- and
//   m  -> m' = coordinates of m (the op that creates B_k gets the unit vector e_k),
//   r  -> r',  r'_k = parity(B_k & r)   (logical-bit parity mask; likewise rc -> rc'),
//   X mask x of the Hamiltonian -> x' = coordinates of x, or "outside D": such a slice connects the
//   support to amplitudes that are exactly zero and drops out of the energy and of every derivative.
// The evaluation kernel (evalc_kernel, kernels_c.cuh) then needs no per-pass descriptors: the cosets
// of a pass are enumerated by inserting zeros at the pivots of the pass's masks inside the prefix
// 2^{d_j}, and the representative is projected onto the common kernel of the r's at run time.
// A controlled op whose control parity mask vanishes on D never fires (control bit constant 0 on the
// support) and is dropped from the tape; its parameters keep their zero gradient.
// Group sizes (meta bits 24-27) must have been set on the ORIGINAL tape (qas_plan_groups with
// gd == nullptr): the no-cross-term test needs complete parity masks.
// meta bits 28-31 of every op = d_j after that op.
#define QAS_CT_MAX_D 12
#define QAS_META_DJ(m) ((m) >> 28)
#define QAS_CT_FULL (-1)          // d == n: nothing to compress, the classic path evaluates it
#define QAS_CT_TOO_BIG (-2)       // d > QAS_CT_MAX_D
// cinfo block of a c-tape record, right after the ops area: head [nsl, B_0 .. B_{n-1}] | list[nsl]
// list entry = x' | slice index << 16
QAS_HD size_t qas_cinfo_words(int n, int S) { return (size_t)1 + (size_t)n + (size_t)S; }

// scratch: ev[n], ec[n] (echelon vectors by highest-bit pivot and their coordinates), mt[nops]
// (per-op m' | d_j << 16).  On success ops / *nops / *ngroups are rewritten in place, chead (1 + n
// words) and list (<= S words) are filled; on QAS_CT_FULL / QAS_CT_TOO_BIG the ops are untouched.
QAS_HD int qas_compress_tape(uint32_t *ops, int *nops_io, int *ngroups_io, int n, const uint32_t *xs, int S,
                             uint32_t *ev, uint32_t *ec, uint32_t *mt, uint32_t *chead, uint32_t *list, int max_d) {
  const int nops = *nops_io;
  uint32_t *basis = chead + 1;
  for (int p = 0; p < n; ++p) ev[p] = ec[p] = 0u;
  int d = 0;
  for (int j = nops - 1; j >= 0; --j) {            // time order
    const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
    uint32_t c = 0u;
    if (qas_meta_kind(o[4]) == QAS_K_U1) {
      uint32_t x = o[0];
      while (x) {
        const int p = qas_msb32(x);
        if (ev[p]) { x ^= ev[p]; c ^= ec[p]; }
        else {
          if (d >= max_d || d >= n - 1) return d >= n - 1 ? QAS_CT_FULL : QAS_CT_TOO_BIG;
          ev[p] = x; ec[p] = c | (1u << d); basis[d] = o[0]; c = 1u << d; ++d;
          break;
        }
      }
    }
    mt[j] = c | ((uint32_t)d << 16);
  }
  // translate in place, dropping controlled ops whose control is constant on the support
  int w = 0, ndrop_groups = 0;
  for (int j = 0; j < nops; ++j) {
    const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
    const uint32_t r = o[1], rc = o[2];
    uint32_t rp = 0u, rcp = 0u;
    for (int k = 0; k < d; ++k) {
      rp |= (uint32_t)qas_parity32(basis[k] & r) << k;
      rcp |= (uint32_t)qas_parity32(basis[k] & rc) << k;
    }
    if (rc && !rcp) { ++ndrop_groups; continue; }  // controlled ops are always one-op groups
    const uint32_t tab = o[3], meta = (o[4] & 0x0fffffffu) | ((mt[j] >> 16) << 28);
    uint32_t *q = ops + (size_t)w * QAS_OP_WORDS;
    q[0] = mt[j] & 0xffffu; q[1] = rp; q[2] = rcp; q[3] = tab; q[4] = meta;
    ++w;
  }
  *nops_io = w;
  *ngroups_io -= ndrop_groups;
  int nsl = 0;
  for (int s = 0; s < S; ++s) {
    uint32_t x = xs[s], c = 0u;
    while (x) {
      const int p = qas_msb32(x);
      if (!ev[p]) break;
      x ^= ev[p]; c ^= ec[p];
    }
    if (!x) list[nsl++] = c | ((uint32_t)s << 16);
  }
  chead[0] = (uint32_t)nsl;
  for (int k = d; k < n; ++k) basis[k] = 0u;
  return d;
}

// ---- c-tape compiler, device formulation ------------------------------------------------------------
// Same result as qas_compress_tape, organised for one GPU thread per circuit without memory-latency chains:
//   * ops live PACKED in shared memory, 3 words each (n <= 16, table index < 65536):
//       w0 = m | r << 16,  w1 = rc | tab << 16,  w2 = meta
//   * the echelon basis is kept in REDUCED form (every pivot bit occurs in its own vector only), so reducing a
//     mask is one pass of independent select-xors over a register table of compile-time length NQ -- no
//     data-dependent loop, no dynamic indexing -- and so are the dual masks r'_k = parity(B_k & r).
// On success the translated 5-word ops are written to the record (gops); on QAS_CT_FULL / QAS_CT_TOO_BIG nothing is
// written and the packed ops are left half-translated: the caller compiles the circuit again and hands the grouped
// original tape to the classic path (qas_unpack_ops).
#define QAS_PK_WORDS 3
struct QasPk3Sink {            // tape-compiler sink: packed ops in shared memory
  uint32_t *pk;
  int cap, n, overflow;
  QAS_HD void emit(uint32_t m, uint32_t r, uint32_t rc, uint32_t tab, uint32_t meta) {
    if (n >= cap) { overflow = 1; return; }
    uint32_t *q = pk + (size_t)n * QAS_PK_WORDS;
    q[0] = m | (r << 16); q[1] = rc | (tab << 16); q[2] = meta;
    n++;
  }
};
// packed ops -> 5-word ops of a classic record (circuits the c-tape compiler leaves to the classic path)
QAS_HD void qas_unpack_ops(const uint32_t *pk, int nops, uint32_t *gops) {
  for (int j = 0; j < nops; ++j) {
    const uint32_t *q = pk + (size_t)j * QAS_PK_WORDS;
    uint32_t *o = gops + (size_t)j * QAS_OP_WORDS;
    o[0] = q[0] & 0xffffu; o[1] = q[0] >> 16; o[2] = q[1] & 0xffffu; o[3] = q[1] >> 16; o[4] = q[2];
  }
}
// greedy grouping of qas_mark_groups on packed ops
QAS_HD int qas_mark_groups_packed(uint32_t *pk, int nops, int gmax) {
  int ngroups = 0;
  for (int j = 0; j < nops;) {
    uint32_t *a = pk + (size_t)j * QAS_PK_WORDS;
    int g = 1;
    if (qas_meta_kind(a[2]) == QAS_K_U1 && (a[1] & 0xffffu) == 0u) {
      while (g < gmax && j + g < nops) {
        const uint32_t *b = pk + (size_t)(j + g) * QAS_PK_WORDS;
        if (qas_meta_kind(b[2]) != QAS_K_U1 || (b[1] & 0xffffu) != 0u) break;
        bool ok = true;
        for (int q = 0; q < g; ++q) {
          const uint32_t *e = pk + (size_t)(j + q) * QAS_PK_WORDS;
          if (qas_parity32((e[0] & 0xffffu) & (b[0] >> 16)) || qas_parity32((b[0] & 0xffffu) & (e[0] >> 16))) ok = false;
        }
        if (!ok) break;
        ++g;
      }
    }
    a[2] |= (uint32_t)g << 24;
    pk[(size_t)(j + g - 1) * QAS_PK_WORDS + 2] |= (uint32_t)g << 26;
    j += g;
    ++ngroups;
  }
  return ngroups;
}

template <int NQ>
QAS_HD int qas_compress_packed(uint32_t *pk, uint32_t *gops, int *nops_io, int *ngroups_io, int n, const uint32_t *xs, int S,
                               uint32_t *chead, uint32_t *list, int max_d) {
  const int nops = *nops_io;
  uint32_t ev[NQ], ec[NQ], bs[NQ];
QAS_PUNROLL
  for (int p = 0; p < NQ; ++p) ev[p] = ec[p] = bs[p] = 0u;
  int d = 0, fail = 0;
  // (no return inside the loops: on the GPU the threads of a warp leave them at different iterations, and an
  // early exit would keep them from re-converging for the rest of the function)
  for (int j = nops - 1; j >= 0; --j) {            // time order
    uint32_t *q = pk + (size_t)j * QAS_PK_WORDS;
    uint32_t c = 0u;
    if (qas_meta_kind(q[2]) == QAS_K_U1 && !fail) {
      const uint32_t m = q[0] & 0xffffu;
      uint32_t x = m;
QAS_PUNROLL
      for (int p = 0; p < NQ; ++p) {               // one-shot reduction against the reduced echelon basis
        const bool hit = ((m >> p) & 1u) && ev[p];
        x ^= hit ? ev[p] : 0u;
        c ^= hit ? ec[p] : 0u;
      }
      if (x && (d >= max_d || d >= n - 1)) {       // would reach the whole register / exceed the largest slot
        fail = d >= n - 1 ? QAS_CT_FULL : QAS_CT_TOO_BIG;
      } else if (x) {                              // enlarges the span: B_d = m, coordinates e_d
        const int pv = qas_msb32(x);
        const uint32_t cv = c | (1u << d);
QAS_PUNROLL
        for (int p = 0; p < NQ; ++p) {
          const bool has = (ev[p] >> pv) & 1u;     // keep the basis reduced: clear the new pivot everywhere else
          ev[p] ^= has ? x : 0u;
          ec[p] ^= has ? cv : 0u;
          if (p == pv) { ev[p] = x; ec[p] = cv; }
          if (p == d) bs[p] = m;
        }
        c = 1u << d;
        ++d;
      }
    }
    q[0] = (q[0] & 0xffff0000u) | c;               // m' in place of m
    q[2] = (q[2] & 0x0fffffffu) | ((uint32_t)d << 28);
  }
  if (fail) return fail;
  // translate, dropping controlled ops whose control is constant on the support
  int w = 0, ndrop = 0;
  for (int j = 0; j < nops; ++j) {
    const uint32_t *q = pk + (size_t)j * QAS_PK_WORDS;
    const uint32_t r = q[0] >> 16, rc = q[1] & 0xffffu;
    uint32_t rp = 0u, rcp = 0u;
QAS_PUNROLL
    for (int k = 0; k < NQ; ++k) {
      rp |= (uint32_t)qas_parity32(bs[k] & r) << k;
      rcp |= (uint32_t)qas_parity32(bs[k] & rc) << k;
    }
    if (rc && !rcp) { ++ndrop; continue; }
    uint32_t *o = gops + (size_t)w * QAS_OP_WORDS;
    o[0] = q[0] & 0xffffu; o[1] = rp; o[2] = rcp; o[3] = q[1] >> 16; o[4] = q[2];
    ++w;
  }
  *nops_io = w;
  *ngroups_io -= ndrop;
  int nsl = 0;
  for (int s = 0; s < S; ++s) {
    const uint32_t x0 = xs[s];
    uint32_t x = x0, c = 0u;
QAS_PUNROLL
    for (int p = 0; p < NQ; ++p) {
      const bool hit = ((x0 >> p) & 1u) && ev[p];
      x ^= hit ? ev[p] : 0u;
      c ^= hit ? ec[p] : 0u;
    }
    if (!x) list[nsl++] = c | ((uint32_t)s << 16);
  }
  chead[0] = (uint32_t)nsl;
  for (int k = 0; k < n; ++k) {
    uint32_t b = 0u;
QAS_PUNROLL
    for (int p = 0; p < NQ; ++p) b = (p == k) ? bs[p] : b;
    chead[1 + k] = b;
  }
  return d;
}

// ---- tile passes (kernels_t.cuh: full-register circuits on 13..16 qubits) ---------------------------------
// A tile pass is a run of consecutive tape ops (whole pass groups) whose xor masks, together with the `clow`
// lowest index bits, span at most TB dimensions.  The pass gets a frame A = [col_0 .. col_{n-1}]:
//   col_0 .. col_{clow-1}   the unit vectors of the low bits (a tile is made of contiguous 2^clow-amplitude chunks),
//   col_clow .. col_{TB-1}  the reduced masks of the pass in the order they enlarge the span (low bits cleared),
//                           completed by unit vectors of unused positions,
//   col_TB .. col_{n-1}     unit vectors of the positions left over: they enumerate the tiles.
// Pass coordinate x (n bits: tile coordinate t | tile number tau << TB) <-> physical index A x.  piv[p] = index of
// the column whose highest set bit is p (every column has a distinct one, so reducing a mask from the top yields
// its coordinates).  Greedy from tape index `start` in direction dir (+1: adjoint sweep, groups marked by GFIRST;
// -1: forward sweep, GLAST); returns the number of ops taken.
#define QAS_TP_MAXQ 16
QAS_HD int qas_plan_tile_pass(const uint32_t *ops, int nops, int start, int dir, int n, int TB, int clow, int max_ops,
                              uint32_t *col, int *piv) {
  const uint32_t lowmask = (1u << clow) - 1u;
  for (int q = 0; q < n; ++q) piv[q] = -1;
  for (int q = 0; q < clow; ++q) { col[q] = 1u << q; piv[q] = q; }
  int d = clow, cnt = 0, j = start;
  while (j >= 0 && j < nops) {
    const uint32_t meta = ops[(size_t)j * QAS_OP_WORDS + 4];
    int g = (int)(dir > 0 ? QAS_META_GFIRST(meta) : QAS_META_GLAST(meta));
    if (g < 1) g = 1;
    if (cnt + g > max_ops) break;
    int added[3], na = 0;
    bool fits = true;
    for (int q = 0; q < g && fits; ++q) {
      const uint32_t *o = ops + (size_t)(j + dir * q) * QAS_OP_WORDS;
      if (qas_meta_kind(o[4]) != QAS_K_U1) continue;
      uint32_t x = o[0] & ~lowmask;
      int pnew = -1;
      for (int p = n - 1; p >= clow; --p)
        if ((x >> p) & 1u) {
          if (piv[p] >= 0) x ^= col[piv[p]];
          else { pnew = p; break; }
        }
      if (pnew >= 0) {
        if (d >= TB) { fits = false; break; }
        col[d] = x; piv[pnew] = d; added[na++] = pnew; ++d;
      }
    }
    if (!fits) {
      for (int a = 0; a < na; ++a) piv[added[a]] = -1;
      d -= na;
      break;
    }
    cnt += g;
    j += dir * g;
  }
  for (int p = clow; p < n && d < TB; ++p) if (piv[p] < 0) { col[d] = 1u << p; piv[p] = d; ++d; }
  for (int p = clow; p < n; ++p) if (piv[p] < 0) { col[d] = 1u << p; piv[p] = d; ++d; }
  return cnt;
}

// Coset representatives of a pass group inside a tile: K = {x in GF(2)^tb : parity(x & r_q) = 0 for every gate q of the
// group} meets every coset of span(m_0 .. m_{g-1}) exactly once (parity(m_a & r_b) = delta_ab), and on K the slot index
// bits of x ^ (c_0 m_0 ^ c_1 m_1) are the gates' logical bits.  kb receives a basis of K in lowest-set-bit echelon form,
// ascending pivots: coset c of the group is represented by the xor of the kb[i] selected by the bits of c, so the
// eight lanes of a quarter-warp (c bits 0..2) land on eight different 16-byte bank groups whenever K has vectors with
// lowest set bits 0, 1 and 2 -- the pivot set of an echelon basis is the smallest possible, i.e. this is the best
// any enumeration of K can do.  Returns the dimension (tb - g).
QAS_HD int qas_tile_group_basis(int g, const uint32_t *m, const uint32_t *r, int tb, uint32_t *kb) {
  uint32_t pv[QAS_TP_MAXQ];
  for (int p = 0; p < tb; ++p) pv[p] = 0u;
  for (int i = 0; i < tb; ++i) {
    uint32_t x = 1u << i;
    for (int q = 0; q < g; ++q) if ((r[q] >> i) & 1u) x ^= m[q];      // projection of e_i onto K along span(m)
    while (x) {
      const int p = qas_lsb32(x);
      if (pv[p]) x ^= pv[p];
      else { pv[p] = x; break; }
    }
  }
  int cnt = 0;
  for (int p = 0; p < tb; ++p) if (pv[p]) kb[cnt++] = pv[p];
  return cnt;
}
