// api.cu -- host side of libqas_b200.so: context, launch logic and the extern "C" ABI declared
// in include/qas_b200.h.  No torch types, no exceptions across the boundary.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"

static thread_local std::string g_create_error;

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
  int *d_counter = nullptr;        // work-queue head of the team kernels
  int blk = 256;                   // threads per CTA of the team kernel
  GateTab *d_U = nullptr;
  GateDTab *d_dU = nullptr;
  int tab_p = 0;
  int resident_p = 0;              // depth of the parameter super-set left on the device by qas_set_params (0: none)
  int32_t *d_k = nullptr;  size_t cap_k = 0;
  double *d_params = nullptr; size_t cap_params = 0;
  double *d_energy = nullptr; size_t cap_energy = 0;
  double *d_grad = nullptr; size_t cap_grad = 0;
  double *d_partial = nullptr; size_t cap_partial = 0;
  double *d_dense = nullptr; size_t cap_dense = 0;
  double2 *d_gstate = nullptr; size_t cap_gstate = 0;
  uint32_t *d_tape = nullptr; size_t cap_tape = 0;
  uint32_t *h_tape = nullptr; size_t cap_htape = 0;   // pinned staging for host-compiled tapes (tiny batches)
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

static int fail(const qas_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}

template <typename Tp>
static cudaError_t ensure(Tp **ptr, size_t *cap, size_t need_elems) {
  if (*cap >= need_elems && *ptr) return cudaSuccess;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *cap = 0;
  size_t want = std::max<size_t>(need_elems, 16);
  cudaError_t e = cudaMalloc((void **)ptr, want * sizeof(Tp));
  if (e == cudaSuccess) *cap = want;
  return e;
}

// ---------------------------------------------------------------------------------- info
extern "C" int qas_abi_version(void) { return QAS_ABI_VERSION; }
extern "C" const char *qas_build_info(void) {
  return "qas_b200 sm_100a; smem-resident complex128 statevector, CNOT-folded tapes, "
         "X-mask sign tables, 2x2 cross-matrix adjoint; built " __DATE__;
}
extern "C" const char *qas_last_error(const qas_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

// ---------------------------------------------------------------------------------- pool checks
static int validate_pool(const qas_pool_entry *pool, int c, int n, int l, std::string &why,
                         int *max_exp) {
  *max_exp = 1;
  for (int i = 0; i < c; ++i) {
    const int g = pool[i].gate_id;
    if (g < 0 || g >= QAS_G_COUNT) { why = "pool entry " + std::to_string(i) + ": unknown gate id"; return 1; }
    const int nw = qas_gate_nwires(g);
    if (pool[i].wire0 < 0 || pool[i].wire0 >= n) { why = "pool entry " + std::to_string(i) + ": wire0 out of range"; return 1; }
    if (nw == 2 && (pool[i].wire1 < 0 || pool[i].wire1 >= n || pool[i].wire1 == pool[i].wire0)) {
      why = "pool entry " + std::to_string(i) + ": wire1 invalid"; return 1;
    }
    if (qas_gate_nparams(g) > l) { why = "pool entry " + std::to_string(i) + ": gate needs more parameter slots than l"; return 1; }
    *max_exp = std::max(*max_exp, qas_gate_expansion(g));
  }
  return 0;
}

static void fill_const_table(std::vector<GateTab> &t) {
  auto set = [&](int id, double2 a, double2 b, double2 c_, double2 d) {
    t[id].u[0] = a; t[id].u[1] = b; t[id].u[2] = c_; t[id].u[3] = d;
  };
  const double r = 0.70710678118654752440;
  const double2 Z0 = make_double2(0, 0), ONE = make_double2(1, 0), I = make_double2(0, 1),
                MI = make_double2(0, -1), MONE = make_double2(-1, 0);
  set(QAS_C_H, make_double2(r, 0), make_double2(r, 0), make_double2(r, 0), make_double2(-r, 0));
  set(QAS_C_X, Z0, ONE, ONE, Z0);
  set(QAS_C_Y, Z0, MI, I, Z0);
  set(QAS_C_Z, ONE, Z0, Z0, MONE);
  set(QAS_C_S, ONE, Z0, Z0, I);
  set(QAS_C_SDG, ONE, Z0, Z0, MI);
  set(QAS_C_T, ONE, Z0, Z0, make_double2(r, r));
  set(QAS_C_TDG, ONE, Z0, Z0, make_double2(r, -r));
  set(QAS_C_SX, make_double2(0.5, 0.5), make_double2(0.5, -0.5), make_double2(0.5, -0.5), make_double2(0.5, 0.5));
  // S*H = [[r, r], [i r, -i r]] ;  H*S^dagger = [[r, -i r], [r, i r]]
  set(QAS_C_SH, make_double2(r, 0), make_double2(r, 0), make_double2(0, r), make_double2(0, -r));
  set(QAS_C_HSDG, make_double2(r, 0), make_double2(0, -r), make_double2(r, 0), make_double2(0, r));
  // parity-diagonal core of exp(+i pi/8 PP) = Ising(-pi/4): diag(e^{+i pi/8}, e^{-i pi/8})
  const double cs = std::cos(M_PI / 8), sn = std::sin(M_PI / 8);
  set(QAS_C_ZZ_MPI4, make_double2(cs, sn), Z0, Z0, make_double2(cs, -sn));
}

// threads per team and threads per CTA by qubit count (n <= 5 normally run the thread-per-circuit
// kernel).  n = 10, 11: two-warp / four-warp teams in 128-thread CTAs -- six / three circuits in flight
// per SM instead of four / two, and less waiting inside a team for the warp that owns the in-support
// cosets of the adjoint sweep (LiH-10q: 2.46 -> 2.23 ms against 128-thread teams in 256-thread CTAs).
constexpr int team_threads(int n) {
  return n <= 4 ? 8 : n == 5 ? 16 : n <= 8 ? 32 : n <= 10 ? 64 : n == 11 ? 128 : 256;
}
constexpr int team_blk(int n) { return (n == 10 || n == 11) ? 128 : 256; }
constexpr int qas_gmax(int n, int T) { return qas_block_bits(n, T) >= 2 ? 2 : 1; }
// H application: 4 amplitudes per thread and a 2-deep prefetch ring where the team has the
// amplitudes for it (measured best on LiH-10q / H2O-8q among (3,1) (3,2) (2,2) (2,4))
constexpr int qas_hrb(int n, int T) { return qas_block_bits(n, T) >= 2 ? 2 : qas_block_bits(n, T); }
constexpr int qas_hpd(int n, int T) { return qas_hrb(n, T) >= 1 ? 2 : 1; }   // PD must divide 2^RB
static int team_threads_for(int n, int shift) {
  const int t = team_threads(n);
  // QAS_B200_TEAM_SHIFT=1: the doubled team in a 256-thread CTA (compiled for n = 8..11, see launch_eval)
  if (shift > 0 && n >= 8 && n <= 11) return t * 2;
  if (shift < 0 && n == 10) return 32;      // QAS_B200_TEAM_SHIFT=-1: one warp per circuit, no team barriers
  return t;
}

// ---- register subspace W of the H application -------------------------------------------------
// Reduced row-echelon basis (pivot = highest set bit) of a set of vectors over GF(2).
struct Gf2Basis {
  uint32_t v[3];
  int pv[3];
  int rank = 0;
  bool add(uint32_t x) {
    for (int i = 0; i < rank; ++i) if ((x >> pv[i]) & 1u) x ^= v[i];
    if (!x || rank == 3) return false;
    const int p = 31 - __builtin_clz(x);
    for (int i = 0; i < rank; ++i) if ((v[i] >> p) & 1u) v[i] ^= x;
    v[rank] = x; pv[rank] = p; ++rank;
    return true;
  }
  uint32_t reduce(uint32_t x) const {
    for (int i = 0; i < rank; ++i) if ((x >> pv[i]) & 1u) x ^= v[i];
    return x;
  }
};
// cost of a candidate W: number of X-mask classes (cosets of W hit by the masks), ties broken
// towards pivots >= 3 (keeps the 8 low index values of consecutive lanes distinct)
static size_t w_cost(const Gf2Basis &b, const std::vector<uint32_t> &xs, std::vector<uint32_t> &tmp) {
  tmp.clear();
  for (uint32_t x : xs) tmp.push_back(b.reduce(x));
  std::sort(tmp.begin(), tmp.end());
  size_t ncls = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
  int low = 0;
  for (int i = 0; i < b.rank; ++i) low += b.pv[i] < 3;
  return ncls * 8 + (size_t)low;
}
static Gf2Basis choose_register_subspace(const std::vector<uint32_t> &xs, int n, int RB) {
  Gf2Basis best;
  if (RB == 0) return best;
  std::vector<uint32_t> cand;
  for (uint32_t x : xs) if (x) cand.push_back(x);
  for (int b = 0; b < n; ++b) cand.push_back(1u << b);
  std::sort(cand.begin(), cand.end());
  cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
  const size_t nc = cand.size();
  std::vector<uint32_t> tmp;
  size_t best_cost = (size_t)-1;
  double combos = 1.0;
  for (int i = 0; i < RB; ++i) combos *= (double)(nc - i) / (i + 1);
  if (combos <= 2.0e5) {      // exhaustive
    std::vector<int> idx(RB);
    for (int i = 0; i < RB; ++i) idx[i] = i;
    while (true) {
      Gf2Basis b;
      bool ok = true;
      for (int i = 0; i < RB && ok; ++i) ok = b.add(cand[idx[i]]);
      if (ok) {
        const size_t cst = w_cost(b, xs, tmp);
        if (cst < best_cost) { best_cost = cst; best = b; }
      }
      int i = RB - 1;
      while (i >= 0 && idx[i] == (int)nc - RB + i) --i;
      if (i < 0) break;
      ++idx[i];
      for (int q = i + 1; q < RB; ++q) idx[q] = idx[q - 1] + 1;
    }
  } else {                    // greedy, one basis vector at a time
    Gf2Basis cur;
    for (int step = 0; step < RB; ++step) {
      Gf2Basis step_best = cur;
      size_t step_cost = (size_t)-1;
      for (uint32_t w : cand) {
        Gf2Basis b = cur;
        if (!b.add(w)) continue;
        const size_t cst = w_cost(b, xs, tmp);
        if (cst < step_cost) { step_cost = cst; step_best = b; }
      }
      cur = step_best;
    }
    best = cur;
  }
  // fallback (cannot happen for n >= RB): unit vectors
  for (int b = n - 1; best.rank < RB && b >= 0; --b) best.add(1u << b);
  return best;
}

// ---------------------------------------------------------------------------------- create
extern "C" int qas_ctx_create(qas_ctx **out, int device, int n_qubits, int init_kind,
                              uint64_t init_basis, const qas_pauli_term *terms, int num_terms,
                              const qas_pool_entry *pool, int c, int l) {
  qas_ctx *ctx = nullptr;
  if (!out) return fail(nullptr, QAS_ERR_INVALID, "out is null");
  *out = nullptr;
  if (n_qubits < 2 || n_qubits > 20)
    return fail(nullptr, QAS_ERR_UNSUPPORTED, "batched evaluator supports 2..20 qubits (use the large-state API above that)");
  if (init_kind < 0 || init_kind > 2) return fail(nullptr, QAS_ERR_INVALID, "bad init_kind");
  if (init_kind == QAS_INIT_BASIS && (init_basis >> n_qubits)) return fail(nullptr, QAS_ERR_INVALID, "init_basis out of range");
  if (!terms || num_terms <= 0 || !pool || c <= 0 || l <= 0) return fail(nullptr, QAS_ERR_INVALID, "empty Hamiltonian or pool");
  std::string why;
  int max_exp = 1;
  if (validate_pool(pool, c, n_qubits, l, why, &max_exp)) return fail(nullptr, QAS_ERR_INVALID, why);
  const uint64_t full = (1ull << n_qubits) - 1;
  for (int t = 0; t < num_terms; ++t)
    if ((terms[t].xmask | terms[t].zmask) & ~full) return fail(nullptr, QAS_ERR_INVALID, "Pauli term acts outside the register");

  cudaError_t e0 = cudaSetDevice(device);
  if (e0 != cudaSuccess) return fail(nullptr, QAS_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e0));

  qas_ctx *cx = new qas_ctx();
  ctx = cx;
  cx->device = device; cx->n = n_qubits; cx->init_kind = init_kind; cx->init_basis = init_basis;
  cx->c = c; cx->l = l; cx->max_expansion = max_exp;
  cx->pool.assign(pool, pool + c);

  // threads per team, amplitudes per thread of the fused gate passes and of the H application
  if (const char *ts = getenv("QAS_B200_TEAM_SHIFT")) cx->team_shift = atoi(ts);
  cx->stage_times = getenv("QAS_B200_STAGE_TIMES") && atoi(getenv("QAS_B200_STAGE_TIMES")) != 0;
  const bool want_phase = getenv("QAS_B200_PHASE_CLOCKS") && atoi(getenv("QAS_B200_PHASE_CLOCKS")) != 0;
  cx->T = team_threads_for(n_qubits, cx->team_shift);
  cx->blk = cx->team_shift > 0 ? 256 : team_blk(n_qubits);
  if (cx->team_shift < 0 && n_qubits == 10) cx->blk = 64;
  cx->GMAX = qas_gmax(n_qubits <= 13 ? n_qubits : 20, cx->T);
  cx->RB = qas_hrb(n_qubits <= 13 ? n_qubits : 20, cx->T);
  cx->PD = qas_hpd(n_qubits <= 13 ? n_qubits : 20, cx->T);
  if (const char *sk = getenv("QAS_B200_SMALL")) cx->small_kernel = atoi(sk) != 0;
  if (n_qubits <= 5 && cx->small_kernel) { cx->GMAX = 1; cx->RB = n_qubits >= 4 ? 1 : 0; cx->PD = cx->RB >= 1 ? 2 : 1; }

  // group the Hamiltonian by X mask: slice key = (xmask, is_imag); term sign from i^{nY}
  std::map<std::pair<uint64_t, int>, std::vector<std::pair<uint32_t, double>>> slices;
  for (int t = 0; t < num_terms; ++t) {
    const int ny = __builtin_popcountll(terms[t].xmask & terms[t].zmask);
    const double sgn = (ny & 2) ? -1.0 : 1.0;
    slices[{terms[t].xmask, ny & 1}].push_back({(uint32_t)terms[t].zmask, sgn * terms[t].coeff});
  }
  // register subspace W (RB-dimensional) minimising the number of X-mask classes
  Gf2Basis W;
  {
    std::vector<uint32_t> xs;
    for (auto &kv : slices) xs.push_back((uint32_t)kv.first.first);
    std::sort(xs.begin(), xs.end());
    xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
    W = choose_register_subspace(xs, n_qubits, cx->RB);
    // basis sorted by ascending pivot: rb0 < rb1 < rb2
    for (int i = 0; i < W.rank; ++i)
      for (int q = i + 1; q < W.rank; ++q)
        if (W.pv[q] < W.pv[i]) { std::swap(W.pv[q], W.pv[i]); std::swap(W.v[q], W.v[i]); }
    for (int q = 0; q < 3; ++q) cx->rb[q] = q < W.rank ? W.pv[q] : 0;
    for (int j = 0; j < 8; ++j) {
      uint32_t w = 0;
      for (int q = 0; q < W.rank; ++q) if ((j >> q) & 1) w ^= W.v[q];
      cx->wc.w[j] = w;
    }
    std::vector<uint32_t> tmp;
    cx->num_classes = (int)(w_cost(W, xs, tmp) / 8);
    cx->hinfo.RB = W.rank;
    for (int q = 0; q < 3; ++q) { cx->hinfo.rb[q] = q < W.rank ? W.pv[q] : 0; cx->hinfo.wv[q] = q < W.rank ? W.v[q] : 0u; }
    if (const char *hs = getenv("QAS_B200_HSPARSE")) cx->hs_min = std::max(0, atoi(hs));
    cx->hinfo.enable = cx->hs_min > 0 && init_kind != QAS_INIT_PLUS && n_qubits >= 8 && n_qubits <= 13;
  }
  // Classes = cosets of W hit by the X masks; real (even nY) classes first, then imaginary ones.
  // Every class owns 2^RB table rows, one per jw (x = rho ^ wcomb[jw]); rows without a slice are
  // zero rows, so the kernel's class loop is fully static (no per-slice descriptor decoding).
  const int NBh = 1 << cx->RB;
  std::vector<uint32_t> sx, cdesc_swz, cdesc_plain, cdesc_u, tz;
  std::vector<int> sbegin;
  std::vector<double> tc;
  auto swz = [](uint32_t i) { return qas_swz(i); };
  for (int imag = 0; imag < 2; ++imag) {
    std::vector<uint32_t> reps;
    for (auto &kv : slices) if (kv.first.second == imag) reps.push_back(W.reduce((uint32_t)kv.first.first));
    std::sort(reps.begin(), reps.end());
    reps.erase(std::unique(reps.begin(), reps.end()), reps.end());
    for (uint32_t rho : reps) {
      cdesc_swz.push_back(swz(rho) << 4);
      cdesc_plain.push_back(rho << 4);
      cdesc_u.push_back(qas_ubar(cx->hinfo, rho));
      for (int jw = 0; jw < NBh; ++jw) {
        const uint32_t x = rho ^ cx->wc.w[jw];
        sx.push_back(x);
        sbegin.push_back((int)tz.size());
        auto it = slices.find({(uint64_t)x, imag});
        if (it != slices.end())
          for (auto &zc : it->second) { tz.push_back(zc.first); tc.push_back(zc.second); }
      }
    }
    (imag ? cx->C_imag : cx->C_real) = (int)reps.size();
  }
  for (int padrow = 0; padrow < 8; ++padrow) { sx.push_back(0u); sbegin.push_back((int)tz.size()); }   // prefetch overrun
  sbegin.push_back((int)tz.size());
  if (tz.empty()) { tz.push_back(0u); tc.push_back(0.0); }
  cx->S = (int)sx.size();
  const size_t dim = (size_t)1 << n_qubits;
  if ((double)cx->S * (double)dim * 8.0 > 4.0e9) {
    delete cx;
    return fail(nullptr, QAS_ERR_UNSUPPORTED, "sign tables would exceed 4 GB; use the large-state API");
  }

  int rc = QAS_OK;
  auto guard = [&](cudaError_t e, const char *what) {
    if (e != cudaSuccess && rc == QAS_OK) {
      g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
      rc = QAS_ERR_CUDA;
    }
  };
  cudaDeviceProp prop;
  guard(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  cx->num_sms = prop.multiProcessorCount;
  guard(cudaStreamCreateWithFlags(&cx->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  guard(cudaEventCreate(&cx->ev0), "cudaEventCreate");
  guard(cudaEventCreate(&cx->ev1), "cudaEventCreate");
  guard(cudaEventCreate(&cx->ek0), "cudaEventCreate");
  guard(cudaEventCreate(&cx->ek1), "cudaEventCreate");
  guard(cudaMalloc(&cx->d_pool, sizeof(qas_pool_entry) * c), "cudaMalloc pool");
  guard(cudaMalloc(&cx->d_counter, sizeof(int)), "cudaMalloc work counter");
  if (want_phase) {
    guard(cudaMalloc(&cx->d_phase, 8 * sizeof(unsigned long long)), "cudaMalloc phase clocks");
    guard(cudaMemset(cx->d_phase, 0, 8 * sizeof(unsigned long long)), "cudaMemset phase clocks");
  }
  guard(cudaMalloc(&cx->d_vtab, sizeof(double) * dim * cx->S), "cudaMalloc sign tables");
  guard(cudaMalloc(&cx->d_sdesc, 4 * std::max<size_t>(cdesc_swz.size(), 1)), "cudaMalloc cdesc");
  guard(cudaMalloc(&cx->d_cdesc_plain, 4 * std::max<size_t>(cdesc_plain.size(), 1)), "cudaMalloc cdesc");
  guard(cudaMalloc(&cx->d_cdesc_u, 4 * std::max<size_t>(cdesc_u.size(), 1)), "cudaMalloc cdesc");
  uint32_t *d_tz = nullptr, *d_sx = nullptr; double *d_tc = nullptr; int *d_sbegin = nullptr;
  guard(cudaMalloc(&d_sx, 4 * cx->S), "cudaMalloc sx");
  guard(cudaMalloc(&d_tz, 4 * tz.size()), "cudaMalloc tz");
  guard(cudaMalloc(&d_tc, 8 * tc.size()), "cudaMalloc tc");
  guard(cudaMalloc(&d_sbegin, 4 * sbegin.size()), "cudaMalloc sbegin");
  if (rc == QAS_OK) {
    guard(cudaMemcpy(cx->d_pool, pool, sizeof(qas_pool_entry) * c, cudaMemcpyHostToDevice), "copy pool");
    guard(cudaMemcpy(d_sx, sx.data(), 4 * sx.size(), cudaMemcpyHostToDevice), "copy sx");
    guard(cudaMemcpy(cx->d_sdesc, cdesc_swz.data(), 4 * cdesc_swz.size(), cudaMemcpyHostToDevice), "copy cdesc");
    guard(cudaMemcpy(cx->d_cdesc_plain, cdesc_plain.data(), 4 * cdesc_plain.size(), cudaMemcpyHostToDevice), "copy cdesc");
    guard(cudaMemcpy(cx->d_cdesc_u, cdesc_u.data(), 4 * cdesc_u.size(), cudaMemcpyHostToDevice), "copy cdesc");
    guard(cudaMemcpy(d_tz, tz.data(), 4 * tz.size(), cudaMemcpyHostToDevice), "copy tz");
    guard(cudaMemcpy(d_tc, tc.data(), 8 * tc.size(), cudaMemcpyHostToDevice), "copy tc");
    guard(cudaMemcpy(d_sbegin, sbegin.data(), 4 * sbegin.size(), cudaMemcpyHostToDevice), "copy sbegin");
    const size_t total = dim * cx->S;
    build_sign_tables_kernel<<<(unsigned)((total + 255) / 256), 256, 0, cx->stream>>>(
        n_qubits, cx->S, cx->RB, cx->rb[0], cx->rb[1], cx->rb[2], cx->wc, d_sx, d_sbegin, d_tz, d_tc, cx->d_vtab);
    guard(cudaGetLastError(), "build_sign_tables launch");
    guard(cudaStreamSynchronize(cx->stream), "build_sign_tables");
  }
  cudaFree(d_tz); cudaFree(d_tc); cudaFree(d_sbegin); cudaFree(d_sx);
  if (rc != QAS_OK) { qas_ctx_destroy(cx); return rc; }
  *out = cx;
  return QAS_OK;
}

extern "C" int qas_ctx_destroy(qas_ctx *ctx) {
  if (!ctx) return QAS_OK;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->d_counter); cudaFree(ctx->d_pool); cudaFree(ctx->d_vtab); cudaFree(ctx->d_sdesc); cudaFree(ctx->d_cdesc_plain); cudaFree(ctx->d_cdesc_u);
  cudaFree(ctx->d_U); cudaFree(ctx->d_dU); cudaFree(ctx->d_k); cudaFree(ctx->d_params);
  cudaFree(ctx->d_energy); cudaFree(ctx->d_grad); cudaFree(ctx->d_partial); cudaFree(ctx->d_dense);
  cudaFree(ctx->d_gstate); cudaFree(ctx->d_mout); cudaFree(ctx->d_tape); cudaFree(ctx->d_phase);
  if (ctx->h_tape) cudaFreeHost(ctx->h_tape);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ek0) cudaEventDestroy(ctx->ek0);
  if (ctx->ek1) cudaEventDestroy(ctx->ek1);
  for (int i = 0; i < 8; ++i) if (ctx->sev[i]) cudaEventDestroy(ctx->sev[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return QAS_OK;
}

extern "C" int qas_debug_phase_clocks(qas_ctx *ctx, uint64_t *out4, int reset) {
  if (!ctx || !out4) return QAS_ERR_INVALID;
  if (!ctx->d_phase) return fail(ctx, QAS_ERR_UNSUPPORTED, "create the context with QAS_B200_PHASE_CLOCKS=1");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  unsigned long long h[8];
  CK(cudaMemcpy(h, ctx->d_phase, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) out4[i] = h[i];
  if (reset) CK(cudaMemset(ctx->d_phase, 0, sizeof h));
  return QAS_OK;
}

extern "C" int qas_get_stats(const qas_ctx *cctx, qas_stats *out) {
  if (!cctx || !out) return QAS_ERR_INVALID;
  qas_ctx *ctx = const_cast<qas_ctx *>(cctx);
  if (ctx->ek_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->ek1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ctx->ek0, ctx->ek1) == cudaSuccess)
      ctx->stats.last_eval_kernel_ms = ms;
    ctx->ek_pending = false;
  }
  *out = ctx->stats;
  return QAS_OK;
}

// ---------------------------------------------------------------------------------- launch
static size_t sdesc_bytes(int S) { return (((size_t)S * 4) + 15) & ~(size_t)15; }

// CTAs per SM the register allocation is bounded for: the 3-gate adjoint pass holds 8 + 8
// amplitudes and a 2x2 cross matrix in registers (128-register budget)
constexpr int qas_minb(int n, int T, int blk) { return blk == 128 ? 3 : (qas_block_bits(n, T) >= 2 ? 2 : 3); }

template <int N, int T, bool GRAD, int RBH, int PD, int GMAX_ = 0, int MINB_ = 0, int BLK = 256>
static cudaError_t launch_smem(qas_ctx *ctx, const EvalParams &P, cudaStream_t st) {
  constexpr int TEAMS = BLK / T;
  const size_t smem = eval_team_smem_bytes(N, T, GRAD, P.ops_cap, true, qas_stage_u(N)) * TEAMS + sdesc_bytes(P.C_real + P.C_imag);
  auto kern = eval_kernel<N, T, GRAD, (GMAX_ ? GMAX_ : qas_gmax(N, T)), RBH, PD, (MINB_ ? MINB_ : qas_minb(N, T, BLK)), BLK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLK, smem);
  const int grid = std::min((P.B + TEAMS - 1) / TEAMS, std::max(1, occ) * ctx->num_sms);   // persistent CTAs, queue-fed
  e = cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  ctx->stats.teams_per_cta = TEAMS;
  ctx->stats.threads_per_team = T;
  ctx->stats.ctas_per_sm = occ;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 0;
  kern<<<grid, BLK, smem, st>>>(P);
  return cudaGetLastError();
}

template <bool GRAD>
static cudaError_t launch_gmem(qas_ctx *ctx, EvalParams &P, cudaStream_t st) {
  const size_t smem = eval_team_smem_bytes(P.n, 256, GRAD, P.ops_cap, false, true) + sdesc_bytes(P.C_real + P.C_imag);
  auto kern = eval_kernel<0, 256, GRAD, 2, 2, 2, 2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
  occ = std::max(1, std::min(occ, 4));
  const size_t per_cta = ((size_t)1 << P.n) * (GRAD ? 2 : 1);
  int grid = std::min(P.B, ctx->num_sms * occ);
  while (grid > 1 && per_cta * grid * 16 > ((size_t)8 << 30)) grid /= 2;
  e = ensure(&ctx->d_gstate, &ctx->cap_gstate, per_cta * grid);
  if (e != cudaSuccess) return e;
  P.gstate = ctx->d_gstate;
  e = cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  ctx->stats.teams_per_cta = 1;
  ctx->stats.threads_per_team = 256;
  ctx->stats.ctas_per_sm = occ;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 1;
  kern<<<grid, 256, smem, st>>>(P);
  return cudaGetLastError();
}

// (N, T) instantiations: the default team size of each qubit count plus tuning alternatives
// n <= 5: thread-per-circuit kernel (eval_small_kernel); QAS_B200_SMALL=0 keeps the team kernel
static bool use_small_kernel(const qas_ctx *ctx) { return ctx->n <= 5 && ctx->small_kernel; }
// ... and the packed tape must fit next to the statevectors (very deep circuits fall back to the team kernel)
static bool small_fits(const qas_ctx *ctx, int ops_cap) {
  return eval_small_smem_bytes(ctx->n, ctx->n <= 4 ? 128 : 64, true, ops_cap, ctx->c) <= 200 * 1024;
}

template <int N, bool GRAD>
static cudaError_t launch_small(qas_ctx *ctx, const EvalParams &P, cudaStream_t st) {
  constexpr int BLK = N <= 4 ? 128 : 64;
  const size_t smem = eval_small_smem_bytes(N, BLK, GRAD, P.ops_cap, P.c);
  auto kern = eval_small_kernel<N, GRAD, BLK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLK, smem);
  ctx->stats.teams_per_cta = BLK;
  ctx->stats.threads_per_team = 1;
  ctx->stats.ctas_per_sm = occ;
  ctx->stats.smem_bytes_per_cta = (int)smem;
  ctx->stats.path = 2;
  kern<<<(P.B + BLK - 1) / BLK, BLK, smem, st>>>(P);
  return cudaGetLastError();
}

template <bool GRAD>
static cudaError_t launch_eval(qas_ctx *ctx, EvalParams &P, cudaStream_t st) {
  if (ctx->small_now) {
    switch (P.n) {
      case 2: return launch_small<2, GRAD>(ctx, P, st);
      case 3: return launch_small<3, GRAD>(ctx, P, st);
      case 4: return launch_small<4, GRAD>(ctx, P, st);
      default: return launch_small<5, GRAD>(ctx, P, st);
    }
  }
  const int nmax_smem = GRAD ? 12 : 13;
  int T = ctx->T, blk = ctx->blk;
  // energy-only launches at n = 10: one warp per circuit (no team barriers; same GMAX / RB / PD, so tapes
  // and sign tables are shared with the gradient launches).  LiH reward-only 72 -> 84 M evals/s; the
  // gradient kernel is slower that way (1.15 -> 1.44 ms: the dense tail of the batch runs on one warp).
  if (!GRAD && P.n == 10 && ctx->team_shift == 0 && T == 64 && blk == 128) { T = 32; blk = 64; }
  if (P.n <= nmax_smem) {
    const size_t smem = eval_team_smem_bytes(P.n, T, GRAD, P.ops_cap, true, qas_stage_u(P.n)) * (blk / T) + sdesc_bytes(P.C_real + P.C_imag);
    if (smem <= 227 * 1024) {
#define QAS_CASE(N_, T_, BLK_) \
  if (P.n == N_ && T == T_ && blk == BLK_) return launch_smem<N_, T_, GRAD, qas_hrb(N_, T_), qas_hpd(N_, T_), 0, 0, BLK_>(ctx, P, st);
      QAS_CASE(2, 8, 256) QAS_CASE(3, 8, 256) QAS_CASE(4, 8, 256) QAS_CASE(5, 16, 256) QAS_CASE(6, 32, 256)
      QAS_CASE(7, 32, 256) QAS_CASE(8, 32, 256) QAS_CASE(9, 64, 256) QAS_CASE(10, 64, 128) QAS_CASE(11, 128, 128)
      QAS_CASE(12, 256, 256)
      QAS_CASE(10, 32, 64)                                                                          // energy-only launches; QAS_B200_TEAM_SHIFT=-1
      QAS_CASE(8, 64, 256) QAS_CASE(9, 128, 256) QAS_CASE(10, 128, 256) QAS_CASE(11, 256, 256)     // QAS_B200_TEAM_SHIFT=1
#undef QAS_CASE
      if constexpr (!GRAD) { if (P.n == 13 && T == 256) return launch_smem<13, 256, false, qas_hrb(13, 256), qas_hpd(13, 256)>(ctx, P, st); }
    }
  }
  // the global-memory instantiation is compiled for 256-thread teams, 3-gate passes, 4-amplitude H rows
  if (ctx->T != 256 || ctx->RB != 2 || ctx->PD != 2 || ctx->GMAX != 2) return cudaErrorInvalidConfiguration;
  return launch_gmem<GRAD>(ctx, P, st);
}

// gate tables sized for depth p (constant rows filled once)
static int ensure_gate_tables(qas_ctx *ctx, int p, cudaStream_t st) {
  if (ctx->tab_p >= p) return QAS_OK;
  const int c = ctx->c;
  cudaFree(ctx->d_U); cudaFree(ctx->d_dU);
  ctx->d_U = nullptr; ctx->d_dU = nullptr; ctx->tab_p = 0; ctx->resident_p = 0;
  const size_t ne = (size_t)QAS_NCONST + (size_t)p * c;
  CK(cudaMalloc(&ctx->d_U, ne * sizeof(GateTab)));
  CK(cudaMalloc(&ctx->d_dU, ne * sizeof(GateDTab)));
  CK(cudaMemsetAsync(ctx->d_dU, 0, ne * sizeof(GateDTab), st));
  std::vector<GateTab> ct(QAS_NCONST);
  fill_const_table(ct);
  CK(cudaMemcpyAsync(ctx->d_U, ct.data(), sizeof(GateTab) * QAS_NCONST, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  ctx->tab_p = p;
  return QAS_OK;
}

extern "C" int qas_set_params(qas_ctx *ctx, int p, const double *params) {
  if (!ctx || p <= 0 || p > 65535 || !params) return fail(ctx, QAS_ERR_INVALID, "bad arguments to qas_set_params");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t npar = (size_t)p * ctx->c * ctx->l;
  { const int rc = ensure_gate_tables(ctx, p, st); if (rc != QAS_OK) return rc; }
  CK(ensure(&ctx->d_params, &ctx->cap_params, npar));
  CK(cudaMemcpyAsync(ctx->d_params, params, npar * 8, cudaMemcpyHostToDevice, st));
  build_gate_table_kernel<<<(p * ctx->c + 127) / 128, 128, 0, st>>>(ctx->d_pool, ctx->c, p, ctx->l, ctx->d_params, ctx->d_U, ctx->d_dU);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  ctx->resident_p = p;
  return QAS_OK;
}

extern "C" int qas_eval_batch(qas_ctx *ctx, int B, int p, const int32_t *k, const double *params,
                              double *energy, double *grad_compact, double *grad_dense,
                              uint32_t flags, void *stream) {
  if (!ctx) return QAS_ERR_INVALID;
  if (B <= 0 || p <= 0 || p > 65535 || !k || (!params && !(flags & QAS_F_PARAMS_RESIDENT)) || !energy)
    return fail(ctx, QAS_ERR_INVALID, "bad arguments to qas_eval_batch");
  const bool dev = flags & QAS_F_DEVICE_PTRS;
  const bool want_grad = grad_compact || grad_dense;
  const int c = ctx->c, l = ctx->l;
  CK(cudaSetDevice(ctx->device));
  // device-pointer calls run on exactly the stream given (NULL = the CUDA default stream, which is
  // what torch.cuda.current_stream() is unless the caller changed it); host-pointer calls default to
  // the context's own stream
  cudaStream_t st = (dev || stream) ? (cudaStream_t)stream : ctx->stream;
  const size_t nk = (size_t)B * p, npar = (size_t)p * c * l, ng = (size_t)B * p * l;

  // one-reward-at-a-time rollouts: no timing events, no structure-list upload (the tapes are compiled
  // on the host and the reward path never reads k on the device)
  const bool tiny = !dev && B <= 16;
  const bool small = use_small_kernel(ctx) && small_fits(ctx, std::max(1, p * ctx->max_expansion));   // n <= 5 kernel compiles from k itself
  const bool resident = (flags & QAS_F_PARAMS_RESIDENT) != 0;
  if (resident && (dev || ctx->resident_p != p))
    return fail(ctx, QAS_ERR_INVALID, "QAS_F_PARAMS_RESIDENT needs host pointers and a prior qas_set_params of the same depth");
  { const int rc = ensure_gate_tables(ctx, p, st); if (rc != QAS_OK) return rc; }
  if (!resident) ctx->resident_p = 0;      // the tables are about to be rebuilt from other parameters

  const int32_t *dk = k;
  const double *dparams = params;
  double *denergy = energy, *dgrad = grad_compact, *ddense = grad_dense;
  if (!dev) {
    CK(ensure(&ctx->d_k, &ctx->cap_k, nk));
    CK(ensure(&ctx->d_params, &ctx->cap_params, npar));
    CK(ensure(&ctx->d_energy, &ctx->cap_energy, (size_t)B));
    if (!(tiny && !want_grad && !small)) CK(cudaMemcpyAsync(ctx->d_k, k, nk * 4, cudaMemcpyHostToDevice, st));
    // the reference's asserts on k (utils.py:12-15).  Batches compiled on the host are checked now;
    // for the others the device code tolerates any key (status word -> NaN energy), so the check
    // runs on the host while the kernels do (below, before the results are copied back)
    if (B <= 16) {
      uint32_t bad = 0;
      const uint32_t cu = (uint32_t)c;
      for (size_t i = 0; i < nk; ++i) bad |= (uint32_t)((uint32_t)k[i] >= cu);
      if (bad) {
        cudaStreamSynchronize(st);
        return fail(ctx, QAS_ERR_INVALID, "structure list entry out of range: need 0 <= k < c");
      }
    }
    if (!resident) CK(cudaMemcpyAsync(ctx->d_params, params, npar * 8, cudaMemcpyHostToDevice, st));
    dk = ctx->d_k; dparams = ctx->d_params; denergy = ctx->d_energy;
    dgrad = nullptr;
    if (grad_dense) { CK(ensure(&ctx->d_dense, &ctx->cap_dense, npar)); ddense = ctx->d_dense; }
  }
  if (want_grad && (!dev || !grad_compact)) {
    CK(ensure(&ctx->d_grad, &ctx->cap_grad, ng));
    dgrad = ctx->d_grad;
  }

  ctx->small_now = small;
  EvalParams P{};
  P.k = dk; P.pool = ctx->d_pool; P.U = ctx->d_U; P.dU = ctx->d_dU; P.grad = dgrad;
  P.vtab = ctx->d_vtab; P.rb0 = ctx->rb[0]; P.rb1 = ctx->rb[1]; P.rb2 = ctx->rb[2];
  P.C_real = ctx->C_real; P.C_imag = ctx->C_imag;
  {   // class descriptors and W combinations as byte offsets, swizzled for the shared-memory path
    const bool smem_path = ctx->n <= (want_grad ? 12 : 13) && !small;
    P.cdesc = smem_path ? ctx->d_sdesc : ctx->d_cdesc_plain;
    P.cdesc_u = ctx->d_cdesc_u;
    P.hs_min = smem_path ? ctx->hs_min : 0;
    for (int j = 0; j < 8; ++j) {
      const uint32_t w = ctx->wc.w[j];
      P.wc16.w[j] = (smem_path ? qas_swz(w) : w) << 4;
    }
  }
  P.energy = denergy; P.gstate = nullptr; P.mout = nullptr; P.phase_clk = ctx->d_phase; P.work_counter = ctx->d_counter;
  P.init_basis = ctx->init_basis;
  P.B = B; P.p = p; P.c = c; P.l = l; P.n = ctx->n;
  P.ops_cap = std::max(1, p * ctx->max_expansion);
  P.init_kind = ctx->init_kind;

  const size_t rec_words = qas_rec_words(P.ops_cap, ctx->n);
  CK(ensure(&ctx->d_tape, &ctx->cap_tape, rec_words * (size_t)B));
  P.tape = ctx->d_tape;

  if (want_grad) {
    CK(ensure(&ctx->d_mout, &ctx->cap_mout, (size_t)B * P.ops_cap * 8));
    P.mout = ctx->d_mout;
  }

  auto stamp = [&](int i) {
    if (!ctx->stage_times) return;
    if (!ctx->sev[i]) cudaEventCreate(&ctx->sev[i]);
    cudaEventRecord(ctx->sev[i], st);
  };
  if (!dev && !tiny) CK(cudaEventRecord(ctx->ev0, st));
  stamp(0);
  if (want_grad) CK(cudaMemsetAsync(dgrad, 0, ng * 8, st));
  stamp(1);
  // Tiny host-pointer batches (the one-reward-at-a-time rollouts of the tree search): the tape
  // compiler and pass planner are host-callable, and one CPU core runs them in a few microseconds,
  // whereas a single GPU thread needs tens -- compile here and upload the records.
  const int swz_plan = (ctx->n <= (want_grad ? 12 : 13) && !small) ? 1 : 0;
  if (small) {
    // the thread-per-circuit kernel compiles its own tapes
  } else if (!dev && B <= 16) {
    const size_t words = rec_words * (size_t)B;
    if (ctx->cap_htape < words) {
      if (ctx->h_tape) cudaFreeHost(ctx->h_tape);
      ctx->h_tape = nullptr; ctx->cap_htape = 0;
      CK(cudaHostAlloc((void **)&ctx->h_tape, std::max<size_t>(words, 4096) * 4, cudaHostAllocDefault));
      ctx->cap_htape = std::max<size_t>(words, 4096);
    }
    uint32_t col[32], row[32];
    for (int b = 0; b < B; ++b) {
      uint32_t *rec = ctx->h_tape + (size_t)b * rec_words;
      QasTapeSink sink{rec + QAS_REC_HDR, P.ops_cap, 0, 0};
      const int stc = qas_compile_circuit(k + (size_t)b * p, p, ctx->pool.data(), c, ctx->n, col, row, sink);
      const int nops = stc ? 0 : sink.n;
      const uint32_t init_index = (ctx->init_kind == QAS_INIT_BASIS && !stc) ? qas_frame_map_basis(col, ctx->n, ctx->init_basis) : 0u;
      rec[0] = (uint32_t)nops; rec[1] = (uint32_t)stc; rec[2] = init_index;
      rec[3] = (uint32_t)qas_plan_groups(rec + QAS_REC_HDR, nops, ctx->GMAX, ctx->n, ctx->init_kind, init_index,
                                         rec + QAS_REC_HDR + (size_t)P.ops_cap * QAS_OP_WORDS, swz_plan,
                                         &ctx->hinfo, rec + qas_rec_hd_offset(P.ops_cap, ctx->n));
    }
    CK(cudaMemcpyAsync(ctx->d_tape, ctx->h_tape, words * 4, cudaMemcpyHostToDevice, st));
  } else {
    const int cthr = 64;
    // ops in shared memory only while that keeps >= 8 CTAs per SM resident (short tapes)
    size_t ops_smem_limit = 24 * 1024;
    if (const char *oe = getenv("QAS_B200_COMPILE_SMEM_KB")) ops_smem_limit = (size_t)std::max(0, atoi(oe)) * 1024;   // tuning knob
    const bool ops_smem = qas_compile_smem_words(ctx->n, P.ops_cap, p, true) * 4 * cthr <= ops_smem_limit;
    const size_t csm = qas_compile_smem_words(ctx->n, P.ops_cap, p, ops_smem) * 4 * cthr;
    if (csm > 200 * 1024) return fail(ctx, QAS_ERR_UNSUPPORTED, "structure lists too long for the tape-compile kernel");
    if (ops_smem) {
      CK(cudaFuncSetAttribute(compile_tapes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
      compile_tapes_kernel<true><<<(B + cthr - 1) / cthr, cthr, csm, st>>>(dk, ctx->d_pool, B, p, c, ctx->n, P.ops_cap, ctx->init_kind,
                                                                         ctx->init_basis, ctx->GMAX, swz_plan, ctx->hinfo, ctx->d_tape);
    } else {
      CK(cudaFuncSetAttribute(compile_tapes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
      compile_tapes_kernel<false><<<(B + cthr - 1) / cthr, cthr, csm, st>>>(dk, ctx->d_pool, B, p, c, ctx->n, P.ops_cap, ctx->init_kind,
                                                                          ctx->init_basis, ctx->GMAX, swz_plan, ctx->hinfo, ctx->d_tape);
    }
  }
  CK(cudaGetLastError());
  stamp(2);
  if (!resident) {
    build_gate_table_kernel<<<(p * c + 127) / 128, 128, 0, st>>>(ctx->d_pool, c, p, l, dparams, ctx->d_U, ctx->d_dU);
    CK(cudaGetLastError());
  }
  if (!tiny) CK(cudaEventRecord(ctx->ek0, st));
  stamp(3);
  if (want_grad) CK(launch_eval<true>(ctx, P, st)); else CK(launch_eval<false>(ctx, P, st));
  if (!tiny) { CK(cudaEventRecord(ctx->ek1, st)); ctx->ek_pending = true; }
  stamp(4);
  ctx->stats.launches += 3;
  if (want_grad && !small) {
    const size_t nthr = (size_t)B * P.ops_cap;
    grad_finalize_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(ctx->d_tape, ctx->d_mout, ctx->d_dU, B, p, l,
                                                                        P.ops_cap, ctx->n, dgrad);
    CK(cudaGetLastError());
    ctx->stats.launches += 1;
  }
  stamp(5);

  if (grad_dense) {
    int chunk = 64;            // circuits per CTA of the first stage (1024 CTAs at B = 65536)
    if (const char *ce = getenv("QAS_B200_REDUCE_CHUNK")) chunk = std::max(1, atoi(ce));   // tuning knob
    const int nchunks = (B + chunk - 1) / chunk;
    int depth_tile = p;
    while ((size_t)depth_tile * c * l * 8 > 96 * 1024 && depth_tile > 1) depth_tile = (depth_tile + 1) / 2;
    const size_t sm = (size_t)depth_tile * c * l * 8;
    CK(ensure(&ctx->d_partial, &ctx->cap_partial, (size_t)nchunks * npar));
    CK(cudaFuncSetAttribute(grad_chunk_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid(nchunks, (p + depth_tile - 1) / depth_tile);
    const int cthreads = std::min(256, std::max(64, ((std::min(depth_tile, p) * l + 31) / 32) * 32));
    grad_chunk_reduce_kernel<<<grid, cthreads, sm, st>>>(dk, dgrad, B, p, c, l, chunk, depth_tile, ctx->d_partial);
    CK(cudaGetLastError());
    stamp(6);
    const double scale = (flags & QAS_F_GRAD_SUM) ? 1.0 : 1.0 / (double)B;
    grad_final_reduce_kernel<<<(unsigned)((npar * 32 + 255) / 256), 256, 0, st>>>(ctx->d_partial, nchunks, npar, scale, ddense);
    CK(cudaGetLastError());
    ctx->stats.launches += 2;
    stamp(7);
  }
  if (ctx->stage_times) {
    cudaStreamSynchronize(st);
    const char *names[7] = {"memset", "compile_tapes", "gate_table", "eval", "finalize", "chunk_reduce", "final_reduce"};
    fprintf(stderr, "[qas stage times] B=%d:", B);
    for (int i = 0; i < (grad_dense ? 7 : 5); ++i) {
      float ms = 0.f;
      if (ctx->sev[i] && ctx->sev[i + 1] && cudaEventElapsedTime(&ms, ctx->sev[i], ctx->sev[i + 1]) == cudaSuccess)
        fprintf(stderr, " %s %.1f us", names[i], ms * 1e3);
    }
    fprintf(stderr, "\n");
  }
  ctx->stats.evals += (uint64_t)B;

  if (!dev) {
    if (!tiny) CK(cudaEventRecord(ctx->ev1, st));
    if (B > 16) {      // deferred range check of the structure lists; branch-free so that it vectorises
      uint32_t bad = 0;
      const uint32_t cu = (uint32_t)c;
      for (size_t i = 0; i < nk; ++i) bad |= (uint32_t)((uint32_t)k[i] >= cu);
      if (bad) {
        cudaStreamSynchronize(st);
        return fail(ctx, QAS_ERR_INVALID, "structure list entry out of range: need 0 <= k < c");
      }
    }
    CK(cudaMemcpyAsync(energy, denergy, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    if (grad_compact) CK(cudaMemcpyAsync(grad_compact, dgrad, ng * 8, cudaMemcpyDeviceToHost, st));
    if (grad_dense) CK(cudaMemcpyAsync(grad_dense, ddense, npar * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (!tiny && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stats.last_kernel_ms = ms;
  }
  return QAS_OK;
}

// ---------------------------------------------------------------------------------- tape export
extern "C" int qas_compile_tape_host(int n_qubits, const qas_pool_entry *pool, int c, int p,
                                     const int32_t *k, uint64_t init_basis, uint32_t *out_ops,
                                     int max_ops, int *num_ops, uint32_t *init_index) {
  if (n_qubits < 1 || n_qubits > QAS_MAX_QUBITS || !pool || c <= 0 || p <= 0 || !k || !out_ops || !num_ops)
    return QAS_ERR_INVALID;
  std::string why;
  int me = 1;
  if (validate_pool(pool, c, n_qubits, 3, why, &me)) { g_create_error = why; return QAS_ERR_INVALID; }
  uint32_t col[32], row[32];
  QasTapeSink sink{out_ops, max_ops, 0, 0};
  const int st = qas_compile_circuit(k, p, pool, c, n_qubits, col, row, sink);
  if (st == 1) { g_create_error = "structure list entry out of range"; return QAS_ERR_INVALID; }
  if (st == 2) { g_create_error = "tape buffer too small"; return QAS_ERR_NOMEM; }
  *num_ops = sink.n;
  if (init_index) *init_index = qas_frame_map_basis(col, n_qubits, init_basis);
  return QAS_OK;
}

extern "C" int qas_plan_passes_host(int n_qubits, uint32_t *ops, int num_ops, int gmax, int init_kind,
                                    uint32_t init_index, int swizzled, uint32_t *gdesc, int gdesc_words,
                                    int *num_groups) {
  if (!ops || num_ops < 0 || gmax < 1 || gmax > 3 || !num_groups || n_qubits < 1 || n_qubits > QAS_MAX_QUBITS)
    return QAS_ERR_INVALID;
  if (gdesc && gdesc_words < num_ops * QAS_GD_STRIDE(n_qubits)) return QAS_ERR_NOMEM;
  *num_groups = qas_plan_groups(ops, num_ops, gmax, n_qubits, init_kind, init_index, gdesc, swizzled);
  return QAS_OK;
}

extern "C" int qas_plan_hsupport_host(int n_qubits, const uint32_t *ops, int num_ops, int init_kind,
                                      uint32_t init_index, int w_rank, const int *w_pivots,
                                      const uint32_t *w_basis, uint32_t *hdesc, int hdesc_words) {
  if ((num_ops > 0 && !ops) || num_ops < 0 || n_qubits < 1 || n_qubits > QAS_MAX_QUBITS || w_rank < 0 || w_rank > 3 ||
      (w_rank > 0 && (!w_pivots || !w_basis)) || !hdesc)
    return QAS_ERR_INVALID;
  if (hdesc_words < QAS_HD_WORDS) return QAS_ERR_NOMEM;
  QasHInfo hi{};
  hi.RB = w_rank; hi.enable = 1;
  for (int q = 0; q < w_rank; ++q) { hi.rb[q] = w_pivots[q]; hi.wv[q] = w_basis[q]; }
  std::vector<uint32_t> tmp(ops, ops + (size_t)num_ops * QAS_OP_WORDS);
  std::vector<uint32_t> gd((size_t)std::max(1, num_ops) * QAS_GD_STRIDE(n_qubits));
  for (int w = 0; w < QAS_HD_WORDS; ++w) hdesc[w] = 0u;
  qas_plan_groups(tmp.data(), num_ops, 2, n_qubits, init_kind, init_index, gd.data(), 0, &hi, hdesc);
  return QAS_OK;
}

extern "C" int qas_compile_tape(const qas_ctx *ctx, int p, const int32_t *k, uint32_t *out_ops,
                                int max_ops, int *num_ops, uint32_t *init_index) {
  if (!ctx) return QAS_ERR_INVALID;
  return qas_compile_tape_host(ctx->n, ctx->pool.data(), ctx->c, p, k, ctx->init_basis, out_ops,
                               max_ops, num_ops, init_index);
}

// ---------------------------------------------------------------------------------- optimiser
extern "C" int qas_adam_step(void *params, const void *grad, void *m, void *v, const void *noise, uint64_t n,
                             int t, double stepsize, double beta1, double beta2, double eps,
                             double noise_factor, void *stream) {
  if (!params || !grad || !m || !v || n == 0 || t < 1) { g_create_error = "bad arguments to qas_adam_step"; return QAS_ERR_INVALID; }
  const double step = stepsize * std::sqrt(1.0 - std::pow(beta2, (double)t)) / (1.0 - std::pow(beta1, (double)t));
  adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (double *)params, (const double *)grad, (double *)m, (double *)v, (const double *)noise, (size_t)n, step, beta1,
      beta2, eps, noise_factor);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string("adam_step_kernel: ") + cudaGetErrorString(e); return QAS_ERR_CUDA; }
  return QAS_OK;
}

// ---------------------------------------------------------------------------------- FP64 peak
extern "C" int qas_fp64_peak_tflops(int device, int reps, double *tflops) {
  qas_ctx *ctx = nullptr;
  if (!tflops) return QAS_ERR_INVALID;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double *d = nullptr;
  CK(cudaMalloc(&d, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double best = 0.0;
  for (int r = 0; r < std::max(reps, 1) + 1; ++r) {
    CK(cudaEventRecord(a));
    dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
    if (r > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *tflops = best;
  return QAS_OK;
}
