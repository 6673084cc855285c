// api.cu -- host side of libqas_b200.so: context, launch logic and the extern "C" ABI declared
// in include/qas_b200.h.  No torch types, no exceptions across the boundary.
#include <chrono>
#if defined(__SSE2__) || defined(_M_X64)
#include <emmintrin.h>
#define QAS_HAVE_SSE2 1
#endif
#include "ctx.h"
#include "launch_cfg.h"

thread_local std::string g_create_error;

// The support bins of one call run on forked streams; with the driver's default of 8 hardware work queues they
// alias (with each other, and with NCCL's streams in a multi-GPU process) and serialise.  The driver reads the
// variable when the CUDA context is created: set it when the library is loaded, unless the host already did.
__attribute__((constructor)) static void qas_request_work_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

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
  if (const char *te = getenv("QAS_B200_TILED")) cx->tiled = atoi(te) != 0;
  if (const char *te = getenv("QAS_B200_TILED_FORCE_N")) cx->tiled_force_n = atoi(te);
  if (const char *te = getenv("QAS_B200_TILED_TB")) cx->tiled_tb = atoi(te) == 12 ? 12 : 11;
  if (const char *te = getenv("QAS_B200_TILED_CS")) cx->tiled_cs = atoi(te);
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
  guard(cudaStreamCreateWithFlags(&cx->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
  guard(cudaEventCreateWithFlags(&cx->copy_gate, cudaEventDisableTiming), "cudaEventCreate");
  guard(cudaEventCreateWithFlags(&cx->copy_done, cudaEventDisableTiming), "cudaEventCreate");
  guard(cudaEventCreate(&cx->ev0), "cudaEventCreate");
  guard(cudaEventCreate(&cx->ev1), "cudaEventCreate");
  guard(cudaEventCreate(&cx->ek0), "cudaEventCreate");
  guard(cudaEventCreate(&cx->ek1), "cudaEventCreate");
  guard(cudaMalloc(&cx->d_pool, sizeof(qas_pool_entry) * c), "cudaMalloc pool");
  guard(cudaMalloc(&cx->d_bins, sizeof(int) * 2 * QAS_MAX_BINS), "cudaMalloc bins");
  // support-compressed path: all-zero start states only (a basis start needs role flips the kernel does not implement)
  {
    const char *ce = getenv("QAS_B200_COMPRESS");
    // below 9 qubits the classic team kernels already run one warp per circuit and most circuits span the whole
    // register (H2O-8q: 76 %), so binning costs more than it saves; the packed tape format holds 16 qubits
    int cmin = 9;
    if (const char *me = getenv("QAS_B200_COMPRESS_MIN_N")) cmin = std::max(7, atoi(me));
    cx->compress = (!ce || atoi(ce) != 0) && init_kind == QAS_INIT_ZERO && n_qubits >= cmin && n_qubits <= 16;
    if (const char *de = getenv("QAS_B200_COMPRESS_DMIN")) cx->c_dmin = std::max(1, std::min(QAS_CT_MAX_D, atoi(de)));
    if (const char *ge = getenv("QAS_B200_GATE_IN_COMPILE")) cx->gate_in_compile = atoi(ge) != 0;
    if (const char *se = getenv("QAS_B200_SPARSE_REDUCE")) cx->sparse_reduce = atoi(se) != 0;
    cx->c_maxd = std::min(QAS_CT_MAX_D, n_qubits - 1);
    cx->c_dmin = std::min(cx->c_dmin, cx->c_maxd);
    cx->c_nbins = cx->c_maxd - cx->c_dmin + 1;
  }
  cx->d_counter = cx->d_bins ? cx->d_bins + QAS_MAX_BINS + cx->c_nbins : nullptr;
  std::vector<uint32_t> fz;
  std::vector<double> fc;
  std::vector<int> fbegin;
  if (cx->compress) {
    for (int imag = 0; imag < 2; ++imag) {
      for (auto &kv : slices) {
        if (kv.first.second != imag) continue;
        cx->sxflat.push_back((uint32_t)kv.first.first);
        fbegin.push_back((int)fz.size());
        for (auto &zc : kv.second) { fz.push_back(zc.first); fc.push_back(zc.second); }
      }
      if (!imag) cx->S_real_flat = (int)cx->sxflat.size();
    }
    fbegin.push_back((int)fz.size());
    cx->S_flat = (int)cx->sxflat.size();
    if (cx->S_flat > 65535 || (double)cx->S_flat * (double)dim * 8.0 > 4.0e9) cx->compress = false;
  }
  if (cx->compress) {
    guard(cudaMalloc(&cx->d_vflat, sizeof(double) * dim * cx->S_flat), "cudaMalloc flat sign tables");
    guard(cudaMalloc(&cx->d_sxflat, 4 * (size_t)cx->S_flat), "cudaMalloc sxflat");
    guard(cudaEventCreateWithFlags(&cx->fork_ev, cudaEventDisableTiming), "cudaEventCreate");
    for (int i = 0; i <= cx->c_nbins && i < QAS_MAX_BINS; ++i) {
      guard(cudaStreamCreateWithFlags(&cx->side[i], cudaStreamNonBlocking), "cudaStreamCreate");
      guard(cudaEventCreateWithFlags(&cx->join_ev[i], cudaEventDisableTiming), "cudaEventCreate");
    }
  }
  if (want_phase) {
    guard(cudaMalloc(&cx->d_phase, 8 * (QAS_MAX_BINS + 1) * sizeof(unsigned long long)), "cudaMalloc phase clocks");
    guard(cudaMemset(cx->d_phase, 0, 8 * (QAS_MAX_BINS + 1) * sizeof(unsigned long long)), "cudaMemset phase clocks");
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
    std::vector<uint8_t> npk((size_t)c);
    for (int i = 0; i < c; ++i) {
      const int g = pool[i].gate_id;
      npk[(size_t)i] = (uint8_t)std::min(qas_gate_nparams(g), l);
      if (g == QAS_G_CRX || g == QAS_G_CRY || g == QAS_G_CRZ || g == QAS_G_CROT || g == QAS_G_CPHASE) cx->pool_ctrl_param = true;
    }
    guard(cudaMalloc(&cx->d_npar_key, (size_t)c), "cudaMalloc npar");
    guard(cudaMemcpy(cx->d_npar_key, npk.data(), (size_t)c, cudaMemcpyHostToDevice), "copy npar");
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
    if (cx->compress) {     // flat rows [slice][physical index] for the compressed path (RB = 0 layout)
      uint32_t *d_fz = nullptr; double *d_fc = nullptr; int *d_fb = nullptr;
      guard(cudaMalloc(&d_fz, 4 * std::max<size_t>(fz.size(), 1)), "cudaMalloc fz");
      guard(cudaMalloc(&d_fc, 8 * std::max<size_t>(fc.size(), 1)), "cudaMalloc fc");
      guard(cudaMalloc(&d_fb, 4 * fbegin.size()), "cudaMalloc fbegin");
      if (rc == QAS_OK) {
        guard(cudaMemcpy(cx->d_sxflat, cx->sxflat.data(), 4 * cx->sxflat.size(), cudaMemcpyHostToDevice), "copy sxflat");
        guard(cudaMemcpy(d_fz, fz.data(), 4 * fz.size(), cudaMemcpyHostToDevice), "copy fz");
        guard(cudaMemcpy(d_fc, fc.data(), 8 * fc.size(), cudaMemcpyHostToDevice), "copy fc");
        guard(cudaMemcpy(d_fb, fbegin.data(), 4 * fbegin.size(), cudaMemcpyHostToDevice), "copy fbegin");
        const size_t ftotal = dim * cx->S_flat;
        WComb nowc{};
        build_sign_tables_kernel<<<(unsigned)((ftotal + 255) / 256), 256, 0, cx->stream>>>(
            n_qubits, cx->S_flat, 0, 0, 0, 0, nowc, cx->d_sxflat, d_fb, d_fz, d_fc, cx->d_vflat);
        guard(cudaGetLastError(), "build flat sign tables launch");
        guard(cudaStreamSynchronize(cx->stream), "build flat sign tables");
        // Slices whose Z masks span r <= 4 dimensions: V_s[i] = lut[parities of i with a basis zeta of that span].
        // z_t = xor_q a_tq zeta_q  =>  parity((i ^ x_s) & z_t) = popc(a_t & (idx(i) ^ kappa)) mod 2, with
        // idx(i) = sum_q parity(i & zeta_q) << q and kappa = idx(x_s) folded into the table.
        std::vector<uint4> sinfo((size_t)cx->S_flat);
        std::vector<double> slut((size_t)cx->S_flat * 16, 0.0);
        int lut_max = 4;
        if (const char *le = getenv("QAS_B200_C_LUT_RANK")) lut_max = std::max(0, std::min(4, atoi(le)));   // tuning knob: 0 = always the flat table
        for (int s_ = 0; s_ < cx->S_flat; ++s_) {
          // zeta: the masks chosen as basis; red[q]: zeta_q reduced against the earlier ones (echelon form, pivot =
          // its highest bit); rco[q]: coordinates of red[q] in terms of the zetas
          std::vector<uint32_t> zeta, red, rco;
          auto reduce = [&](uint32_t x, uint32_t *a) {
            for (size_t q = 0; q < red.size(); ++q)
              if ((x >> (31 - __builtin_clz(red[q]))) & 1u) { x ^= red[q]; *a ^= rco[q]; }
            return x;
          };
          bool small = lut_max > 0;
          for (int t = fbegin[s_]; t < fbegin[s_ + 1] && small; ++t) {
            uint32_t a = 0u;
            const uint32_t x = reduce(fz[t], &a);
            if (!x) continue;
            if ((int)zeta.size() >= lut_max) { small = false; break; }
            rco.push_back(a | (1u << zeta.size()));
            zeta.push_back(fz[t]);
            red.push_back(x);
          }
          if (!small) { sinfo[s_] = make_uint4(255u, 0u, 0u, 0u); continue; }
          const int r = (int)zeta.size();
          uint32_t kappa = 0u;
          for (int q = 0; q < r; ++q) kappa |= (uint32_t)(__builtin_popcount(cx->sxflat[s_] & zeta[q]) & 1) << q;
          for (uint32_t idx = 0; idx < 16u; ++idx) {
            double v = 0.0;
            for (int t = fbegin[s_]; t < fbegin[s_ + 1]; ++t) {
              uint32_t at = 0u;
              reduce(fz[t], &at);        // coordinates a_t of the term's mask (it lies in the span)
              v += (__builtin_popcount(at & (idx ^ kappa) & ((1u << r) - 1u)) & 1) ? -fc[t] : fc[t];
            }
            slut[(size_t)s_ * 16 + idx] = v;
          }
          uint32_t z4[4] = {0u, 0u, 0u, 0u};
          for (int q = 0; q < r; ++q) z4[q] = zeta[q];
          sinfo[s_] = make_uint4((uint32_t)r, z4[0] | (z4[1] << 16), z4[2] | (z4[3] << 16), 0u);
        }
        guard(cudaMalloc(&cx->d_sinfo, sizeof(uint4) * sinfo.size()), "cudaMalloc sinfo");
        guard(cudaMalloc(&cx->d_slut, sizeof(double) * slut.size()), "cudaMalloc slut");
        if (rc == QAS_OK) {
          guard(cudaMemcpy(cx->d_sinfo, sinfo.data(), sizeof(uint4) * sinfo.size(), cudaMemcpyHostToDevice), "copy sinfo");
          guard(cudaMemcpy(cx->d_slut, slut.data(), sizeof(double) * slut.size(), cudaMemcpyHostToDevice), "copy slut");
        }
      }
      cudaFree(d_fz); cudaFree(d_fc); cudaFree(d_fb);
    }
  }
  cudaFree(d_tz); cudaFree(d_tc); cudaFree(d_sbegin); cudaFree(d_sx);
  if (rc != QAS_OK) { qas_ctx_destroy(cx); return rc; }
  *out = cx;
  return QAS_OK;
}

extern "C" int qas_ctx_destroy(qas_ctx *ctx) {
  if (!ctx) return QAS_OK;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->d_bins); cudaFree(ctx->d_order); cudaFree(ctx->d_vflat); cudaFree(ctx->d_sxflat); cudaFree(ctx->d_sinfo); cudaFree(ctx->d_slut);
  if (ctx->h_bins) cudaFreeHost(ctx->h_bins);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  for (int i = 0; i < QAS_MAX_BINS; ++i) {
    if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]);
    if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
  }
  cudaFree(ctx->d_pool); cudaFree(ctx->d_npar_key); cudaFree(ctx->d_vtab); cudaFree(ctx->d_sdesc); cudaFree(ctx->d_cdesc_plain); cudaFree(ctx->d_cdesc_u);
  cudaFree(ctx->d_U); cudaFree(ctx->d_dU); cudaFree(ctx->d_k); cudaFree(ctx->d_params);
  cudaFree(ctx->d_energy); cudaFree(ctx->d_grad); cudaFree(ctx->d_partial); cudaFree(ctx->d_dense);
  if (ctx->h_k8) cudaFreeHost(ctx->h_k8);
  cudaFree(ctx->d_gstate); cudaFree(ctx->d_mout); cudaFree(ctx->d_tape); cudaFree(ctx->d_phase);
  if (ctx->h_tape) cudaFreeHost(ctx->h_tape);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ek0) cudaEventDestroy(ctx->ek0);
  if (ctx->ek1) cudaEventDestroy(ctx->ek1);
  for (int i = 0; i < 8; ++i) if (ctx->sev[i]) cudaEventDestroy(ctx->sev[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->copy_gate) cudaEventDestroy(ctx->copy_gate);
  if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
  cudaFree(ctx->d_k8);
  delete ctx;
  return QAS_OK;
}

extern "C" int qas_ctx_param_slots(const qas_ctx *ctx) { return ctx ? ctx->l : QAS_ERR_INVALID; }

extern "C" int qas_debug_phase_clocks(qas_ctx *ctx, uint64_t *out4, int reset) {
  if (!ctx || !out4) return QAS_ERR_INVALID;
  if (!ctx->d_phase) return fail(ctx, QAS_ERR_UNSUPPORTED, "create the context with QAS_B200_PHASE_CLOCKS=1");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  unsigned long long h[8 * (QAS_MAX_BINS + 1)];
  CK(cudaMemcpy(h, ctx->d_phase, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) out4[i] = h[i];
  if (reset) CK(cudaMemset(ctx->d_phase, 0, sizeof h));
  return QAS_OK;
}
// tiled kernel (kernels_t.cuh): out8 = cycles [prologue, forward, H, adjoint], forward passes, adjoint passes, circuits, 0
extern "C" int qas_debug_phase_clocks_t(qas_ctx *ctx, uint64_t *out8, int reset) {
  if (!ctx || !out8) return QAS_ERR_INVALID;
  if (!ctx->d_phase) return fail(ctx, QAS_ERR_UNSUPPORTED, "create the context with QAS_B200_PHASE_CLOCKS=1");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  unsigned long long h[8];
  CK(cudaMemcpy(h, ctx->d_phase, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  if (reset) CK(cudaMemset(ctx->d_phase, 0, sizeof h));
  return QAS_OK;
}
// per-bin clocks of the compressed kernel: out[bin][0..3] = phase sums, out[bin][4] = circuits; bins = slots c_dmin ..
extern "C" int qas_debug_phase_clocks_c(qas_ctx *ctx, uint64_t *out, int max_bins, int *dmin, int reset) {
  if (!ctx || !out || max_bins < 1) return QAS_ERR_INVALID;
  if (!ctx->d_phase) return fail(ctx, QAS_ERR_UNSUPPORTED, "create the context with QAS_B200_PHASE_CLOCKS=1");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  unsigned long long h[8 * (QAS_MAX_BINS + 1)];
  CK(cudaMemcpy(h, ctx->d_phase, sizeof h, cudaMemcpyDeviceToHost));
  for (int b_ = 0; b_ < std::min(max_bins, QAS_MAX_BINS); ++b_)
    for (int i = 0; i < 8; ++i) out[b_ * 8 + i] = h[8 * (1 + b_) + i];
  if (dmin) *dmin = ctx->c_dmin;
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

// QAS_F_K_U8: structure lists uploaded as one byte per key, widened on the device
__global__ void widen_keys_kernel(const uint8_t *src, int32_t *dst, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const uchar4 v = *reinterpret_cast<const uchar4 *>(src + i);
    *reinterpret_cast<int4 *>(dst + i) = make_int4(v.x, v.y, v.z, v.w);
  } else {
    for (size_t j = i; j < n; ++j) dst[j] = src[j];
  }
}

static QasLaunchCache g_chunk_reduce_cache;      // grad_chunk_reduce_kernel is launched from qas_eval_batch and qas_batch_mean

// The reference's asserts on k (utils.py:12-15: 0 <= k < c) for a whole batch.  Measured on the GPU boxes' hosts the
// plain loop ran at one key per 0.53 ns -- 0.69 ms for the 1.3 M keys of a LiH step, LONGER than the GPU needs for the
// step, so it, not the device, bounded the end-to-end rate (QAS_B200_HOST_TIMES=1 prints the call's host timeline).
// 16 keys per instruction: the byte maximum (pmaxub) / or-ed signed compares, then one test.
static bool keys_out_of_range_u8(const uint8_t *k, size_t nk, uint32_t c) {
  size_t i = 0;
  uint32_t mx = 0;
#if QAS_HAVE_SSE2
  __m128i m = _mm_setzero_si128();
  for (; i + 64 <= nk; i += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i *)(k + i)), b = _mm_loadu_si128((const __m128i *)(k + i + 16));
    const __m128i e = _mm_loadu_si128((const __m128i *)(k + i + 32)), f = _mm_loadu_si128((const __m128i *)(k + i + 48));
    m = _mm_max_epu8(m, _mm_max_epu8(_mm_max_epu8(a, b), _mm_max_epu8(e, f)));
  }
  uint8_t t[16];
  _mm_storeu_si128((__m128i *)t, m);
  for (int j = 0; j < 16; ++j) mx = t[j] > mx ? t[j] : mx;
#endif
  for (; i < nk; ++i) mx = k[i] > mx ? k[i] : mx;
  return mx >= c;
}
static bool keys_out_of_range_i32(const int32_t *k, size_t nk, uint32_t c) {
  size_t i = 0;
  uint32_t bad = 0;
#if QAS_HAVE_SSE2
  if (c <= 0x7fffffffu) {
    __m128i acc = _mm_setzero_si128();
    const __m128i lim = _mm_set1_epi32((int)c - 1), zero = _mm_setzero_si128();
    for (; i + 8 <= nk; i += 8) {
      const __m128i a = _mm_loadu_si128((const __m128i *)(k + i)), b = _mm_loadu_si128((const __m128i *)(k + i + 4));
      acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(_mm_cmpgt_epi32(a, lim), _mm_cmpgt_epi32(b, lim)),
                                           _mm_or_si128(_mm_cmplt_epi32(a, zero), _mm_cmplt_epi32(b, zero))));
    }
    bad = _mm_movemask_epi8(acc) ? 1u : 0u;
  }
#endif
  for (; i < nk; ++i) bad |= (uint32_t)((uint32_t)k[i] >= c);
  return bad != 0;
}

// int32 keys -> one byte per key (page-locked staging of pageable structure lists), range asserts in the same pass
static bool narrow_keys_i32_to_u8(const int32_t *k, uint8_t *h8, size_t nk, uint32_t c) {
  size_t i = 0;
  uint32_t bad = 0;
#if QAS_HAVE_SSE2
  if (c <= 256u) {
    __m128i acc = _mm_setzero_si128();
    const __m128i lim = _mm_set1_epi32((int)c - 1), zero = _mm_setzero_si128();
    for (; i + 16 <= nk; i += 16) {
      const __m128i a = _mm_loadu_si128((const __m128i *)(k + i)), b = _mm_loadu_si128((const __m128i *)(k + i + 4));
      const __m128i e = _mm_loadu_si128((const __m128i *)(k + i + 8)), f = _mm_loadu_si128((const __m128i *)(k + i + 12));
      acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(_mm_cmpgt_epi32(a, lim), _mm_cmpgt_epi32(b, lim)),
                                           _mm_or_si128(_mm_cmpgt_epi32(e, lim), _mm_cmpgt_epi32(f, lim))));
      acc = _mm_or_si128(acc, _mm_or_si128(_mm_or_si128(_mm_cmplt_epi32(a, zero), _mm_cmplt_epi32(b, zero)),
                                           _mm_or_si128(_mm_cmplt_epi32(e, zero), _mm_cmplt_epi32(f, zero))));
      _mm_storeu_si128((__m128i *)(h8 + i), _mm_packus_epi16(_mm_packs_epi32(a, b), _mm_packs_epi32(e, f)));
    }
    bad = _mm_movemask_epi8(acc) ? 1u : 0u;
  }
#endif
  for (; i < nk; ++i) {
    const uint32_t v = (uint32_t)k[i];
    bad |= (uint32_t)(v >= c);
    h8[i] = (uint8_t)v;
  }
  return bad != 0;
}

extern "C" int qas_check_keys_host(const void *k, uint64_t num_keys, int key_bytes, int c, uint8_t *narrowed, int *out_of_range) {
  if (!k || !out_of_range || c <= 0 || (key_bytes != 1 && key_bytes != 4) || (narrowed && (key_bytes != 4 || c > 256))) return QAS_ERR_INVALID;
  bool bad;
  if (key_bytes == 1) bad = keys_out_of_range_u8(static_cast<const uint8_t *>(k), (size_t)num_keys, (uint32_t)c);
  else if (narrowed) bad = narrow_keys_i32_to_u8(static_cast<const int32_t *>(k), narrowed, (size_t)num_keys, (uint32_t)c);
  else bad = keys_out_of_range_i32(static_cast<const int32_t *>(k), (size_t)num_keys, (uint32_t)c);
  *out_of_range = bad ? 1 : 0;
  return QAS_OK;
}

extern "C" int qas_eval_batch(qas_ctx *ctx, int B, int p, const int32_t *k, const double *params,
                              double *energy, double *grad_compact, double *grad_dense,
                              uint32_t flags, void *stream) {
  if (!ctx) return QAS_ERR_INVALID;
  if (B <= 0 || p <= 0 || p > 65535 || !k || (!params && !(flags & QAS_F_PARAMS_RESIDENT)) || !energy)
    return fail(ctx, QAS_ERR_INVALID, "bad arguments to qas_eval_batch");
  const bool dev = flags & QAS_F_DEVICE_PTRS;
  bool k8 = (flags & QAS_F_K_U8) != 0;
  const bool want_grad = grad_compact || grad_dense;
  const int c = ctx->c, l = ctx->l;
  if (k8 && c > 256) return fail(ctx, QAS_ERR_INVALID, "QAS_F_K_U8 needs a pool of at most 256 entries");
  static const bool host_times = []() { const char *e = getenv("QAS_B200_HOST_TIMES"); return e && atoi(e) != 0; }();
  const auto ht0 = std::chrono::steady_clock::now();
  auto hus = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::micro>(t - ht0).count(); };
  CK(cudaSetDevice(ctx->device));
  // device-pointer calls run on exactly the stream given (NULL = the CUDA default stream, which is
  // what torch.cuda.current_stream() is unless the caller changed it); host-pointer calls default to
  // the context's own stream
  cudaStream_t st = (dev || stream) ? (cudaStream_t)stream : ctx->stream;
  const size_t nk = (size_t)B * p, npar = (size_t)p * c * l, ng = (size_t)B * p * l;
  // a call enqueued on a stream that is being captured into a CUDA graph (qas_b200.mcts.driver tuning loop): no
  // timing events (they could not be queried afterwards); every buffer must have been sized by an earlier call
  bool capturing = false;
  {
    cudaStreamCaptureStatus cs_ = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs_) == cudaSuccess) capturing = cs_ != cudaStreamCaptureStatusNone;
  }

  // one-reward-at-a-time rollouts: no timing events, no structure-list upload (the tapes are compiled
  // on the host and the reward path never reads k on the device)
  const bool tiny = !dev && B <= 16;
  const bool small = use_small_kernel(ctx) && small_fits(ctx, std::max(1, p * ctx->max_expansion));   // n <= 5 kernel compiles from k itself
  const bool resident = (flags & QAS_F_PARAMS_RESIDENT) != 0;
  if (resident && (dev || ctx->resident_p != p))
    return fail(ctx, QAS_ERR_INVALID, "QAS_F_PARAMS_RESIDENT needs host pointers and a prior qas_set_params of the same depth");
  { const int rc = ensure_gate_tables(ctx, p, st); if (rc != QAS_OK) return rc; }
  if (!resident) ctx->resident_p = 0;      // the tables are about to be rebuilt from other parameters

  // QAS_F_K_U8 on a small host batch: the host-side compiler and checks below read int32 keys
  std::vector<int32_t> k_wide;
  const uint8_t *k_bytes = reinterpret_cast<const uint8_t *>(k);
  if (k8 && !dev && B <= 16) {
    k_wide.assign(k_bytes, k_bytes + nk);
    k = k_wide.data();
  }
  // Pageable int32 structure lists of a large host batch (plain NumPy arrays: what search() hands the drop-in classes):
  // the driver would stage 4 bytes per key through its own bounce buffers while the calling thread waits.  Narrow them
  // to one byte per key into a page-locked buffer instead -- the reference's range asserts (utils.py:12-15) in the same
  // pass -- and upload a quarter of the bytes asynchronously, like a QAS_F_K_U8 call.
  bool k_checked = false;
  if (!dev && !k8 && B > 16 && !capturing && c <= 256) {
    cudaPointerAttributes pa{};
    bool pageable = false;
    if (cudaPointerGetAttributes(&pa, k) == cudaSuccess) pageable = pa.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    if (pageable) {
      if (ctx->cap_hk8 < nk) {
        if (ctx->h_k8) cudaFreeHost(ctx->h_k8);
        ctx->h_k8 = nullptr; ctx->cap_hk8 = 0;
        CK(cudaHostAlloc((void **)&ctx->h_k8, std::max<size_t>(nk, 4096), cudaHostAllocDefault));
        ctx->cap_hk8 = std::max<size_t>(nk, 4096);
      }
      // the previous call's upload from this buffer has completed: every host-pointer call ends with a stream sync
      uint8_t *h8 = ctx->h_k8;
      if (narrow_keys_i32_to_u8(k, h8, nk, (uint32_t)c)) return fail(ctx, QAS_ERR_INVALID, "structure list entry out of range: need 0 <= k < c");
      k_bytes = h8;
      k8 = true;
      k_checked = true;
    }
  }
  const bool k8_dev = k8 && (dev || B > 16);      // keys reach the device as bytes and are widened there
  const int32_t *dk = k;
  const double *dparams = params;
  double *denergy = energy, *dgrad = grad_compact, *ddense = grad_dense;
  bool k_on_copy_stream = false;
  if (k8_dev) {
    CK(ensure(&ctx->d_k, &ctx->cap_k, (nk + 3) & ~(size_t)3));
    dk = ctx->d_k;
  }
  if (!dev) {
    CK(ensure(&ctx->d_k, &ctx->cap_k, (nk + 3) & ~(size_t)3));
    CK(ensure(&ctx->d_params, &ctx->cap_params, npar));
    CK(ensure(&ctx->d_energy, &ctx->cap_energy, (size_t)B));
    if (k8_dev) CK(ensure(&ctx->d_k8, &ctx->cap_k8, (nk + 3) & ~(size_t)3));
    if (B > 16 && !capturing) {
      // large batches: the upload of the structure lists runs on the copy stream, so that the memsets and the
      // gate-table kernel of this call overlap it; the tape compiler waits for it
      CK(cudaEventRecord(ctx->copy_gate, st));
      CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_gate, 0));
      if (k8_dev) CK(cudaMemcpyAsync(ctx->d_k8, k_bytes, nk, cudaMemcpyHostToDevice, ctx->copy_stream));
      else CK(cudaMemcpyAsync(ctx->d_k, k, nk * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
      CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
      k_on_copy_stream = true;
    } else if (!(tiny && !want_grad && !small)) {
      CK(cudaMemcpyAsync(ctx->d_k, k, nk * 4, cudaMemcpyHostToDevice, st));
    }
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
  ctx->small_now = small;
  const int ops_cap = std::max(1, p * ctx->max_expansion);
  // ---- which kernels evaluate this call, decided once: the tape planner, the class descriptors and the
  // launches must agree on the index layout (swizzled shared-memory states vs the plain global-memory slab)
  int cT = 0, cblk = 0;
  const int cpath = small ? 2 : qas_classic_path(ctx, want_grad, ops_cap, &cT, &cblk);
  if (cpath < 0)
    return fail(ctx, QAS_ERR_UNSUPPORTED, "tapes of this depth do not fit the shared-memory kernel of this register size and the "
                                          "global-memory kernel is not compiled for its pass / H-row configuration");
  const bool smem_path = cpath == 0;
  const int swz_plan = smem_path ? 1 : 0;
  // support-compressed path: every circuit is binned by the dimension d of its support span; d < n goes to the
  // compressed kernel of its bin, d == n stays with the classic kernel (the full bin)
  const bool use_c = ctx->compress && !small && qas_ccompile_smem_bytes(ctx, ops_cap, p) <= 96 * 1024 &&
                     (size_t)QAS_NCONST + (size_t)p * c <= 65535;
  if (want_grad && (!dev || !grad_compact)) {
    CK(ensure(&ctx->d_grad, &ctx->cap_grad, ng));
    dgrad = ctx->d_grad;
  }

  EvalParams P{};
  P.k = dk; P.pool = ctx->d_pool; P.U = ctx->d_U; P.dU = ctx->d_dU; P.grad = dgrad;
  P.vtab = ctx->d_vtab; P.rb0 = ctx->rb[0]; P.rb1 = ctx->rb[1]; P.rb2 = ctx->rb[2];
  P.C_real = ctx->C_real; P.C_imag = ctx->C_imag;
  {   // class descriptors and W combinations as byte offsets, swizzled for the shared-memory path
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
  P.ops_cap = ops_cap;
  P.init_kind = ctx->init_kind;

  const size_t rec_words = std::max(qas_rec_words(ops_cap, ctx->n),
                                    use_c ? (size_t)QAS_REC_HDR + (size_t)ops_cap * QAS_OP_WORDS + qas_cinfo_words(ctx->n, ctx->S_flat) : (size_t)0);
  CK(ensure(&ctx->d_tape, &ctx->cap_tape, rec_words * (size_t)B));
  P.tape = ctx->d_tape;
  P.rec_stride = rec_words;
  if (use_c) {
    CK(ensure(&ctx->d_order, &ctx->cap_order, (size_t)QAS_MAX_BINS * B));
    P.order = ctx->d_order + (size_t)ctx->c_nbins * B;
    P.count = ctx->d_bins + ctx->c_nbins;
  }

  if (want_grad) {
    CK(ensure(&ctx->d_mout, &ctx->cap_mout, (size_t)B * ops_cap * 8));
    P.mout = ctx->d_mout;
  }

  auto stamp = [&](int i) {
    if (!ctx->stage_times) return;
    if (!ctx->sev[i]) cudaEventCreate(&ctx->sev[i]);
    cudaEventRecord(ctx->sev[i], st);
  };
  if (!dev && !tiny && !capturing) CK(cudaEventRecord(ctx->ev0, st));
  stamp(0);
  // (Tried: zero-fill + gate table on a forked stream next to the tape compiler.  Slower -- 0.780 -> 0.798 ms per LiH
  // step: the 31 MB zero-fill then lands between the compiler and the evaluation kernels and pushes the tape
  // records out of L2.)
  // Only the dense batch mean is wanted (the compact gradients never leave the device): no zero-fill.  The reduction
  // then reads exactly the entries the kernels write -- slot j of depth i iff j < #parameters of the gate k[b][i] -- and
  // skips circuits whose energy is NaN (bad structure list / tape overflow: nothing was written; NaN parameters: the
  // entries would be zeroed by nan_to_num anyway).  Not with pools whose controlled rotations the c-tape translation
  // may drop, and not for host-compiled handfuls of circuits.
  const bool sparse_reduce = want_grad && grad_dense && !grad_compact && B > 16 && !(use_c && ctx->pool_ctrl_param) && ctx->sparse_reduce;
  if (want_grad && !sparse_reduce) { CK(cudaMemsetAsync(dgrad, 0, ng * 8, st)); ctx->stats.launches += 1; }
  if (!small) CK(cudaMemsetAsync(ctx->d_bins, 0, sizeof(int) * 2 * QAS_MAX_BINS, st));   // bin counts + work-queue heads
  stamp(1);
  // gate table first (it only needs the parameters), then the structure lists: the upload has had the memsets and
  // this kernel to hide behind
  // device-compiled compressed batches: spare CTAs of the c-tape compile kernel build the table (no launch of its own)
  const bool gate_table_in_compile = !resident && use_c && !(!small && !dev && B <= 16) && ctx->gate_in_compile;
  if (!resident && !gate_table_in_compile) {
    build_gate_table_kernel<<<(p * c + 127) / 128, 128, 0, st>>>(ctx->d_pool, c, p, l, dparams, ctx->d_U, ctx->d_dU);
    CK(cudaGetLastError());
    ctx->stats.launches += 1;
  }
  stamp(2);
  if (k_on_copy_stream) CK(cudaStreamWaitEvent(st, ctx->copy_done, 0));
  if (k8_dev) {
    const uint8_t *src8 = dev ? k_bytes : ctx->d_k8;
    widen_keys_kernel<<<(unsigned)((nk / 4 + 256) / 256), 256, 0, st>>>(src8, ctx->d_k, nk);
    CK(cudaGetLastError());
    ctx->stats.launches += 1;
  }
  // Tiny host-pointer batches (the one-reward-at-a-time rollouts of the tree search): the tape
  // compiler and pass planner are host-callable, and one CPU core runs them in a few microseconds,
  // whereas a single GPU thread needs tens -- compile here and upload the records.
  const bool host_compile = !small && !dev && B <= 16;
  int h_count[QAS_MAX_BINS] = {0};
  if (small) {
    // the thread-per-circuit kernel compiles its own tapes
  } else if (host_compile) {
    const size_t words = rec_words * (size_t)B;
    if (ctx->cap_htape < words) {
      if (ctx->h_tape) cudaFreeHost(ctx->h_tape);
      ctx->h_tape = nullptr; ctx->cap_htape = 0;
      CK(cudaHostAlloc((void **)&ctx->h_tape, std::max<size_t>(words, 4096) * 4, cudaHostAllocDefault));
      ctx->cap_htape = std::max<size_t>(words, 4096);
    }
    const size_t bwords = (size_t)QAS_MAX_BINS * (1 + (size_t)B);
    if (use_c && ctx->cap_hbins < bwords) {
      if (ctx->h_bins) cudaFreeHost(ctx->h_bins);
      ctx->h_bins = nullptr; ctx->cap_hbins = 0;
      CK(cudaHostAlloc((void **)&ctx->h_bins, std::max<size_t>(bwords, 1024) * 4, cudaHostAllocDefault));
      ctx->cap_hbins = std::max<size_t>(bwords, 1024);
    }
    uint32_t col[32], row[32], ev[32], ec[32];
    std::vector<uint32_t> mt((size_t)ops_cap);
    // a handful of circuits: ONE compressed launch sized for the largest support among them (the kernel's loops
    // follow every circuit's own d), instead of one launch per bin
    int hd_[16], dmax_ = 0;
    for (int b = 0; b < B; ++b) {
      uint32_t *rec = ctx->h_tape + (size_t)b * rec_words;
      uint32_t *rops = rec + QAS_REC_HDR, *rci = rops + (size_t)ops_cap * QAS_OP_WORDS;
      QasTapeSink sink{rops, ops_cap, 0, 0};
      const int stc = qas_compile_circuit(k + (size_t)b * p, p, ctx->pool.data(), c, ctx->n, col, row, sink);
      int nops = stc ? 0 : sink.n;
      const uint32_t init_index = (ctx->init_kind == QAS_INIT_BASIS && !stc) ? qas_frame_map_basis(col, ctx->n, ctx->init_basis) : 0u;
      int ngroups = qas_mark_groups(rops, nops, ctx->GMAX);
      int d = QAS_CT_FULL;
      if (use_c) d = qas_compress_tape(rops, &nops, &ngroups, ctx->n, ctx->sxflat.data(), ctx->S_flat, ev, ec, mt.data(), rci,
                                       rci + 1 + ctx->n, ctx->c_maxd);
      rec[0] = (uint32_t)nops; rec[1] = (uint32_t)stc; rec[3] = (uint32_t)ngroups;
      if (d >= 0) {
        rec[2] = (uint32_t)d;
      } else {
        rec[2] = init_index;
        qas_plan_marked(rops, nops, ngroups, ctx->n, ctx->init_kind, init_index, rci, swz_plan, &ctx->hinfo,
                        rec + qas_rec_hd_offset(ops_cap, ctx->n));
      }
      hd_[b] = d;
      if (d > dmax_) dmax_ = d;
    }
    if (use_c) {
      const int cbin = std::max(dmax_, ctx->c_dmin) - ctx->c_dmin;
      for (int b = 0; b < B; ++b) {
        const int bin = hd_[b] >= 0 ? cbin : ctx->c_nbins;
        ctx->h_bins[QAS_MAX_BINS + (size_t)bin * B + h_count[bin]++] = b;
      }
    }
    CK(cudaMemcpyAsync(ctx->d_tape, ctx->h_tape, words * 4, cudaMemcpyHostToDevice, st));
    if (use_c) {
      for (int i = 0; i < QAS_MAX_BINS; ++i) ctx->h_bins[i] = h_count[i];
      CK(cudaMemcpyAsync(ctx->d_bins, ctx->h_bins, sizeof(int) * QAS_MAX_BINS, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->d_order, ctx->h_bins + QAS_MAX_BINS, sizeof(int) * (size_t)(ctx->c_nbins + 1) * B, cudaMemcpyHostToDevice, st));
    }
  } else if (use_c) {
    CompileCParams C{};
    C.k = dk; C.pool = ctx->d_pool; C.sxflat = ctx->d_sxflat; C.tape = ctx->d_tape; C.bin_count = ctx->d_bins;
    C.bin_order = ctx->d_order; C.B = B; C.p = p; C.c = c; C.n = ctx->n; C.ops_cap = ops_cap; C.S = ctx->S_flat;
    C.gmax = ctx->GMAX; C.dmin = ctx->c_dmin; C.max_d = ctx->c_maxd; C.full_bin = ctx->c_nbins; C.rec_stride = rec_words;
    C.params = dparams; C.U = ctx->d_U; C.dU = ctx->d_dU; C.l = l; C.gate_blocks = gate_table_in_compile ? (p * c + 63) / 64 : 0;
    CK(qas_launch_compile_c(ctx, C, st));
    ctx->stats.launches += 1;
  } else {
    { const int rc_ = qas_launch_compile_classic(ctx, dk, B, p, ops_cap, swz_plan, rec_words, st); if (rc_ != QAS_OK) return rc_; }
    ctx->stats.launches += 1;
  }
  CK(cudaGetLastError());
  stamp(3);
  if (!tiny && !capturing) CK(cudaEventRecord(ctx->ek0, st));
  auto launch_classic = [&](cudaStream_t s_) { return want_grad ? qas_launch_eval_grad(ctx, P, s_) : qas_launch_eval_nograd(ctx, P, s_); };
  if (!use_c) {
    CK(launch_classic(st));
    ctx->stats.launches += 1;
  } else {
    // one launch per bin, largest registers first.  Device-compiled batches do not know the bin counts on the
    // host: every bin is launched (an empty one exits at once), each on its own forked stream so that the
    // persistent CTAs of a small bin fill the tail of a large one.  Host-compiled batches launch the non-empty
    // bins back to back on the call's stream.
    EvalCParams Q{};
    Q.tape = ctx->d_tape; Q.U = ctx->d_U; Q.vflat = ctx->d_vflat; Q.sinfo = ctx->d_sinfo; Q.slut = ctx->d_slut; Q.energy = denergy;
    Q.dU = ctx->d_dU; Q.grad = dgrad; Q.p = p; Q.l = l;
    Q.rec_stride = rec_words; Q.n = ctx->n; Q.ops_cap = ops_cap; Q.S = ctx->S_flat; Q.S_real = ctx->S_real_flat;
    const bool fork = !host_compile;
    if (fork) CK(cudaEventRecord(ctx->fork_ev, st));
    int nside = 0;
    bool main_used = false;
    for (int bin = ctx->c_nbins; bin >= 0; --bin) {
      if (host_compile && h_count[bin] == 0) continue;
      const int max_work = host_compile ? h_count[bin] : B;
      const bool full = bin == ctx->c_nbins;
      // the call's own stream takes the first compressed bin (the largest registers); the full bin -- classic
      // planning of its circuits, then the classic kernel -- and the smaller bins run on forked streams
      cudaStream_t s_ = st;
      const bool on_side = fork && (full || main_used);
      if (on_side) {
        s_ = ctx->side[nside];
        CK(cudaStreamWaitEvent(s_, ctx->fork_ev, 0));
      } else if (!full) {
        main_used = true;
      }
      if (full) {
        if (!host_compile) {
          CK(qas_launch_plan_full(ctx, ctx->d_tape, rec_words, P.count, P.order, B, ops_cap, swz_plan, s_));
          ctx->stats.launches += 1;
        }
        const int keepB = P.B;
        if (host_compile) P.B = max_work;
        const cudaError_t le = launch_classic(s_);
        P.B = keepB;
        CK(le);
        if (want_grad) {     // cross matrices of the full bin's circuits -> compact gradients, on the same stream
          const size_t nthr = (size_t)max_work * ops_cap;
          grad_finalize_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, s_>>>(ctx->d_tape, rec_words, ctx->d_mout, ctx->d_dU, max_work, p, l,
                                                                              ops_cap, dgrad, P.order, P.count);
          CK(cudaGetLastError());
          ctx->stats.launches += 1;
        }
      } else {
        Q.slot_bits = ctx->c_dmin + bin;
        Q.order = ctx->d_order + (size_t)bin * B;
        Q.count = ctx->d_bins + bin;
        Q.work_counter = ctx->d_bins + QAS_MAX_BINS + bin;
        Q.phase_clk = ctx->d_phase ? ctx->d_phase + 8 * (1 + bin) : nullptr;      // per-bin diagnostic clocks
        CK(qas_launch_evalc(ctx, Q, max_work, want_grad, s_));
      }
      ctx->stats.launches += 1;
      if (on_side) {
        CK(cudaEventRecord(ctx->join_ev[nside], s_));
        CK(cudaStreamWaitEvent(st, ctx->join_ev[nside], 0));
        ++nside;
      }
    }
  }
  if (!tiny && !capturing) { CK(cudaEventRecord(ctx->ek1, st)); ctx->ek_pending = true; }
  // Host-pointer calls: the energies are final once the evaluation kernels have finished -- read them back on the copy
  // stream while the finalize / batch-reduction kernels still run (LiH: 0.5 MB = 25 us of PCIe time behind 40 us of kernels)
  // Only into page-locked caller memory: a device-to-host copy into pageable pages blocks the calling thread until it
  // has run, and the reduction kernels would not even be enqueued before the evaluation kernels end (measured: 0.86 ->
  // 1.54 ms per call).
  bool energy_early = false;
  if (!dev && !tiny && !capturing && want_grad && ctx->copy_stream) {
    cudaPointerAttributes pa{};
    if (cudaPointerGetAttributes(&pa, energy) == cudaSuccess) energy_early = pa.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  if (energy_early) {
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ek1, 0));
    CK(cudaMemcpyAsync(energy, denergy, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
  }
  stamp(4);
  if (want_grad && !small && !use_c) {
    const size_t nthr = (size_t)B * ops_cap;
    grad_finalize_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, st>>>(ctx->d_tape, rec_words, ctx->d_mout, ctx->d_dU, B, p, l,
                                                                        ops_cap, dgrad, nullptr, nullptr);
    CK(cudaGetLastError());
    ctx->stats.launches += 1;
  }
  stamp(5);

  if (grad_dense) {
    int chunk = 64;            // circuits per thread group of the first stage
    if (const char *ce = getenv("QAS_B200_REDUCE_CHUNK")) chunk = std::max(1, atoi(ce));   // tuning knob
    int depth_tile = p;
    while ((size_t)depth_tile * c * l * 8 > 96 * 1024 && depth_tile > 1) depth_tile = (depth_tile + 1) / 2;
    const size_t sm1 = (size_t)depth_tile * c * l * 8;
    const int gthreads = std::min(256, std::max(64, ((std::min(depth_tile, p) * l + 31) / 32) * 32));
    // thread groups per CTA, each with its own accumulator copy over its own `chunk` circuits, summed in fixed group
    // order before the partial is written: a quarter of the partial-sum traffic of one group per CTA (LiH: 18.7 MB
    // written and re-read per 65536 circuits before)
    int ngroups_r = 4;
    if (const char *ge = getenv("QAS_B200_REDUCE_GROUPS")) ngroups_r = std::max(1, std::min(8, atoi(ge)));   // tuning knob
    while (ngroups_r > 1 && (sm1 * ngroups_r > 96 * 1024 || gthreads * ngroups_r > 512 || (B + chunk * ngroups_r - 1) / (chunk * ngroups_r) < ctx->num_sms))
      ngroups_r >>= 1;
    const int nchunks = (B + chunk * ngroups_r - 1) / (chunk * ngroups_r);
    const size_t sm = sm1 * ngroups_r + (((size_t)c + 15) & ~(size_t)15);
    CK(ensure(&ctx->d_partial, &ctx->cap_partial, (size_t)nchunks * npar));
    CK(qas_prepare_launch(g_chunk_reduce_cache, ctx->device, grad_chunk_reduce_kernel, 0, sm, nullptr));
    dim3 grid(nchunks, (p + depth_tile - 1) / depth_tile);
    grad_chunk_reduce_kernel<<<grid, gthreads * ngroups_r, sm, st>>>(dk, dgrad, B, p, c, l, chunk, depth_tile, ngroups_r, ctx->d_partial,
                                                                     sparse_reduce ? denergy : nullptr, sparse_reduce ? ctx->d_npar_key : nullptr);
    CK(cudaGetLastError());
    stamp(6);
    const double scale = (flags & QAS_F_GRAD_SUM) ? 1.0 : 1.0 / (double)B;
    grad_final_reduce_kernel<<<(unsigned)((npar * 32 + 255) / 256), 256, 0, st>>>(ctx->d_partial, nchunks, npar, npar, scale, ddense);
    CK(cudaGetLastError());
    ctx->stats.launches += 2;
    stamp(7);
  }
  if (ctx->stage_times) {
    cudaStreamSynchronize(st);
    const char *names[7] = {"memset", "gate_table", "compile_tapes", "eval", "finalize", "chunk_reduce", "final_reduce"};
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
    const auto ht1 = std::chrono::steady_clock::now();
    if (B > 16 && !k_checked) {      // deferred range check of the structure lists; branch-free so that it vectorises
      const bool bad = k8 ? keys_out_of_range_u8(k_bytes, nk, (uint32_t)c) : keys_out_of_range_i32(k, nk, (uint32_t)c);
      if (bad) {
        cudaStreamSynchronize(st);
        if (energy_early) cudaStreamSynchronize(ctx->copy_stream);      // nothing may land in caller memory after the return
        return fail(ctx, QAS_ERR_INVALID, "structure list entry out of range: need 0 <= k < c");
      }
    }
    const auto ht2 = std::chrono::steady_clock::now();
    if (energy_early) CK(cudaStreamWaitEvent(st, ctx->copy_done, 0));
    else CK(cudaMemcpyAsync(energy, denergy, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    if (grad_compact) CK(cudaMemcpyAsync(grad_compact, dgrad, ng * 8, cudaMemcpyDeviceToHost, st));
    if (grad_dense) CK(cudaMemcpyAsync(grad_dense, ddense, npar * 8, cudaMemcpyDeviceToHost, st));
    const auto ht3 = std::chrono::steady_clock::now();
    CK(cudaStreamSynchronize(st));
    if (host_times && !tiny) {      // diagnostic: where the calling thread spends a host-pointer call
      const auto ht4 = std::chrono::steady_clock::now();
      fprintf(stderr, "[qas host times] B=%d: enqueued at %.0f us, range check done %.0f, copies enqueued %.0f, synchronised %.0f\n",
              B, hus(ht1), hus(ht2), hus(ht3), hus(ht4));
    }
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

extern "C" int qas_plan_tile_pass_host(int n_qubits, const uint32_t *ops, int num_ops_total, int start, int dir, int tb,
                                       int clow, int max_ops, uint32_t *col, int *piv, int *num_ops) {
  if ((num_ops_total > 0 && !ops) || num_ops_total < 0 || n_qubits < 1 || n_qubits > QAS_TP_MAXQ || tb < 1 || tb > n_qubits ||
      clow < 0 || clow > tb || (dir != 1 && dir != -1) || max_ops < 3 || !col || !piv || !num_ops)
    return QAS_ERR_INVALID;
  *num_ops = qas_plan_tile_pass(ops, num_ops_total, start, dir, n_qubits, tb, clow, max_ops, col, piv);
  return QAS_OK;
}

extern "C" int qas_tile_group_basis_host(int g, const uint32_t *m, const uint32_t *r, int tb, uint32_t *kb, int *dim) {
  if (g < 1 || g > 2 || !m || !r || tb < g || tb > QAS_TP_MAXQ || !kb || !dim) return QAS_ERR_INVALID;
  *dim = qas_tile_group_basis(g, m, r, tb, kb);
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

extern "C" int qas_compile_ctape_host(int n_qubits, const qas_pool_entry *pool, int c, int p, const int32_t *k,
                                      int gmax, const uint32_t *xmasks, int num_xmasks, int max_d,
                                      uint32_t *out_ops, int max_ops, int *num_ops, int *num_groups, int *dim,
                                      uint32_t *basis_out, uint32_t *list_out, int *num_list, int variant) {
  if (n_qubits < 1 || n_qubits > QAS_MAX_QUBITS || !pool || c <= 0 || p <= 0 || !k || !out_ops || !num_ops || !num_groups ||
      !dim || !basis_out || gmax < 1 || gmax > 3 || num_xmasks < 0 || (num_xmasks > 0 && (!xmasks || !list_out)) || !num_list ||
      max_d < 1 || max_d > QAS_CT_MAX_D || variant < 0 || variant > 1)
    return QAS_ERR_INVALID;
  std::string why;
  int me = 1;
  if (validate_pool(pool, c, n_qubits, 3, why, &me)) { g_create_error = why; return QAS_ERR_INVALID; }
  uint32_t col[32], row[32];
  int nops = 0, ngroups = 0, d = QAS_CT_FULL;
  std::vector<uint32_t> cinfo(qas_cinfo_words(n_qubits, num_xmasks));
  if (variant == 0) {        // reference formulation (also the host path of tiny batches)
    QasTapeSink sink{out_ops, max_ops, 0, 0};
    const int st = qas_compile_circuit(k, p, pool, c, n_qubits, col, row, sink);
    if (st == 1) { g_create_error = "structure list entry out of range"; return QAS_ERR_INVALID; }
    if (st == 2) { g_create_error = "tape buffer too small"; return QAS_ERR_NOMEM; }
    nops = sink.n;
    ngroups = qas_mark_groups(out_ops, nops, gmax);
    std::vector<uint32_t> ev(n_qubits), ec(n_qubits), mt((size_t)std::max(1, nops));
    d = qas_compress_tape(out_ops, &nops, &ngroups, n_qubits, xmasks, num_xmasks, ev.data(), ec.data(), mt.data(),
                          cinfo.data(), cinfo.data() + 1 + n_qubits, max_d);
  } else {                   // device formulation: packed ops, reduced echelon basis in registers
    if (n_qubits > 16 || (size_t)QAS_NCONST + (size_t)p * c > 65535) { g_create_error = "packed c-tape compiler needs n <= 16 and p * c < 65524"; return QAS_ERR_UNSUPPORTED; }
    std::vector<uint32_t> pk((size_t)std::max(1, max_ops) * QAS_PK_WORDS);
    QasPk3Sink sink{pk.data(), max_ops, 0, 0};
    const int st = qas_compile_circuit(k, p, pool, c, n_qubits, col, row, sink);
    if (st == 1) { g_create_error = "structure list entry out of range"; return QAS_ERR_INVALID; }
    if (st == 2) { g_create_error = "tape buffer too small"; return QAS_ERR_NOMEM; }
    nops = sink.n;
    ngroups = qas_mark_groups_packed(pk.data(), nops, gmax);
    const int nops_raw = nops, ngroups_raw = ngroups;
    if (n_qubits <= 10) d = qas_compress_packed<10>(pk.data(), out_ops, &nops, &ngroups, n_qubits, xmasks, num_xmasks, cinfo.data(), cinfo.data() + 1 + n_qubits, max_d);
    else if (n_qubits <= 12) d = qas_compress_packed<12>(pk.data(), out_ops, &nops, &ngroups, n_qubits, xmasks, num_xmasks, cinfo.data(), cinfo.data() + 1 + n_qubits, max_d);
    else d = qas_compress_packed<16>(pk.data(), out_ops, &nops, &ngroups, n_qubits, xmasks, num_xmasks, cinfo.data(), cinfo.data() + 1 + n_qubits, max_d);
    if (d < 0) {
      QasPk3Sink again{pk.data(), max_ops, 0, 0};
      qas_compile_circuit(k, p, pool, c, n_qubits, col, row, again);
      qas_mark_groups_packed(pk.data(), nops_raw, gmax);
      qas_unpack_ops(pk.data(), nops_raw, out_ops);
      nops = nops_raw;
      ngroups = ngroups_raw;
    }
  }
  *num_ops = nops;
  *num_groups = ngroups;
  *dim = d;
  *num_list = 0;
  if (d >= 0) {
    for (int q = 0; q < n_qubits; ++q) basis_out[q] = cinfo[1 + q];
    *num_list = (int)cinfo[0];
    for (int s = 0; s < *num_list; ++s) list_out[s] = cinfo[1 + n_qubits + s];
  }
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

// ---------------------------------------------------------------------------------- graph-capturable tuning step
// circuitModelTuning (mcts.py:381-412) as a replayable CUDA graph: nothing in the captured work may depend on
// host state, so the step counter, the bias-corrected step size and the loss history live on the device.
//   qas_tuning_tick:  t <- t + 1 ; step <- stepsize sqrt(1 - b2^t) / (1 - b1^t) ; losses[t - 1] <- *loss
//   qas_adam_step_dev: the update of qas_adam_step with `step` read from the device
static __global__ void tuning_tick_kernel(int *t, double *step, const double *loss, double *losses, int max_losses,
                                          double stepsize, double b1, double b2) {
  const int tt = *t + 1;
  *t = tt;
  *step = stepsize * sqrt(1.0 - pow(b2, (double)tt)) / (1.0 - pow(b1, (double)tt));
  if (losses && tt - 1 < max_losses) losses[tt - 1] = *loss;
}
static __global__ void adam_step_dev_kernel(double *x, const double *grad, double *m, double *v, size_t n, const double *step,
                                            double b1, double b2, double eps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double g = nan_to_num_d(grad[i]);
  const double mi = b1 * m[i] + (1.0 - b1) * g;
  const double vi = b2 * v[i] + (1.0 - b2) * g * g;
  m[i] = mi;
  v[i] = vi;
  x[i] = x[i] - *step * mi / (sqrt(vi) + eps);
}
extern "C" int qas_tuning_tick(void *t_dev, void *step_dev, const void *loss_dev, void *losses_dev, int max_losses,
                               double stepsize, double beta1, double beta2, void *stream) {
  if (!t_dev || !step_dev || !loss_dev) { g_create_error = "bad arguments to qas_tuning_tick"; return QAS_ERR_INVALID; }
  tuning_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((int *)t_dev, (double *)step_dev, (const double *)loss_dev,
                                                       (double *)losses_dev, max_losses, stepsize, beta1, beta2);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string("tuning_tick_kernel: ") + cudaGetErrorString(e); return QAS_ERR_CUDA; }
  return QAS_OK;
}
extern "C" int qas_adam_step_dev(void *params, const void *grad, void *m, void *v, uint64_t n, const void *step_dev,
                                 double beta1, double beta2, double eps, void *stream) {
  if (!params || !grad || !m || !v || n == 0 || !step_dev) { g_create_error = "bad arguments to qas_adam_step_dev"; return QAS_ERR_INVALID; }
  adam_step_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (double *)params, (const double *)grad, (double *)m, (double *)v, (size_t)n, (const double *)step_dev, beta1, beta2, eps);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string("adam_step_dev_kernel: ") + cudaGetErrorString(e); return QAS_ERR_CUDA; }
  return QAS_OK;
}

// ---------------------------------------------------------------------------------- batch mean of given gradients
// stack -> nan_to_num -> mean | sum (mcts.py:345-347) of per-circuit compact gradients that are already on the
// device, with the deterministic two-stage reduction of qas_eval_batch.  Used where the per-circuit gradient is
// not a plain adjoint gradient (VQLS: quotient rule of two Pauli sums).
extern "C" int qas_batch_mean(qas_ctx *ctx, int B, int p, const int32_t *k, const double *grad_compact,
                              double *grad_dense, uint32_t flags, void *stream) {
  if (!ctx || B <= 0 || p <= 0 || !k || !grad_compact || !grad_dense) return fail(ctx, QAS_ERR_INVALID, "bad arguments to qas_batch_mean");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int c = ctx->c, l = ctx->l;
  const size_t npar = (size_t)p * c * l;
  const int chunk = 64, nchunks = (B + chunk - 1) / chunk;
  int depth_tile = p;
  while ((size_t)depth_tile * c * l * 8 > 96 * 1024 && depth_tile > 1) depth_tile = (depth_tile + 1) / 2;
  const size_t sm = (size_t)depth_tile * c * l * 8 + (((size_t)c + 15) & ~(size_t)15);
  CK(ensure(&ctx->d_partial, &ctx->cap_partial, (size_t)nchunks * npar));
  CK(qas_prepare_launch(g_chunk_reduce_cache, ctx->device, grad_chunk_reduce_kernel, 0, sm, nullptr));
  dim3 grid(nchunks, (p + depth_tile - 1) / depth_tile);
  const int cthreads = std::min(256, std::max(64, ((std::min(depth_tile, p) * l + 31) / 32) * 32));
  grad_chunk_reduce_kernel<<<grid, cthreads, sm, st>>>(k, grad_compact, B, p, c, l, chunk, depth_tile, 1, ctx->d_partial, nullptr, nullptr);
  CK(cudaGetLastError());
  const double scale = (flags & QAS_F_GRAD_SUM) ? 1.0 : 1.0 / (double)B;
  grad_final_reduce_kernel<<<(unsigned)((npar * 32 + 255) / 256), 256, 0, st>>>(ctx->d_partial, nchunks, npar, npar, scale, grad_dense);
  CK(cudaGetLastError());
  ctx->stats.launches += 2;
  return QAS_OK;
}

// VQLS cost (qml_models_legacy.py:1880-1951 restated as two Pauli-sum expectations N, D on n qubits):
//   loss = 1/2 - |N| / (2 n |D|),   dloss = -sign(N) / (2 n |D|) dN + |N| sign(D) / (2 n D^2) dD
// elementwise over the batch; gN is overwritten with the loss gradient (device pointers, asynchronous).
static __global__ void vqls_combine_kernel(const double *N, const double *D, double *gN, const double *gD, int B, int pl,
                                           double nq, double *loss) {
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= (size_t)B * (size_t)(pl > 0 ? pl : 1)) return;
  const int b = (int)(pl > 0 ? x / pl : x);
  const double nv = N[b], dv = D[b], an = fabs(nv), ad = fabs(dv);
  if (pl == 0 || x % pl == 0) loss[b] = 0.5 - 0.5 * an / (nq * ad);
  if (pl > 0) {
    const double sn = nv > 0.0 ? 1.0 : (nv < 0.0 ? -1.0 : 0.0), sd = dv > 0.0 ? 1.0 : (dv < 0.0 ? -1.0 : 0.0);
    const double wN = (-0.5 / nq) * sn / ad, wD = (0.5 / nq) * an * sd / (ad * ad);
    gN[x] = wN * gN[x] + wD * gD[x];
  }
}
extern "C" int qas_vqls_combine(const void *N, const void *D, void *gN, const void *gD, int B, int pl, int n_qubits,
                                void *loss, void *stream) {
  if (!N || !D || !loss || B <= 0 || pl < 0 || n_qubits < 1 || (pl > 0 && (!gN || !gD))) { g_create_error = "bad arguments to qas_vqls_combine"; return QAS_ERR_INVALID; }
  const size_t total = (size_t)B * (size_t)(pl > 0 ? pl : 1);
  vqls_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const double *)N, (const double *)D, (double *)gN, (const double *)gD, B, pl, (double)n_qubits, (double *)loss);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string("vqls_combine_kernel: ") + cudaGetErrorString(e); return QAS_ERR_CUDA; }
  return QAS_OK;
}

// ---------------------------------------------------------------------------------- shard reduction
extern "C" int qas_reduce_rows(const void *rows, int nrows, uint64_t row_stride, uint64_t total, double scale, void *out,
                               void *stream) {
  if (!rows || !out || nrows <= 0 || total == 0 || row_stride < total) { g_create_error = "bad arguments to qas_reduce_rows"; return QAS_ERR_INVALID; }
  grad_final_reduce_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const double *)rows, nrows, (size_t)row_stride, (size_t)total, scale, (double *)out);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_create_error = std::string("grad_final_reduce_kernel: ") + cudaGetErrorString(e); return QAS_ERR_CUDA; }
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
