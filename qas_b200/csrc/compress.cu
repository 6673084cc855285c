// compress.cu -- support-compressed register: translation of a compiled tape into coordinates of
// its support span (host-only translation unit).
//
// From a basis state |b0> a circuit only populates b0 ^ D, D = span of the xor masks of its 2x2 ops.
// With a basis B_0..B_{d-1} of D and i(t) = b0 ^ xor_k t_k B_k the evaluation becomes that of a
// d-qubit register started in |0..0>:
//   m  -> m'  = coordinates of m                     (pair mask)
//   r  -> r'  , r'_k = parity(B_k & r)               (logical-bit parity mask), f  = parity(b0 & r)
//   rc -> rc' , likewise                             (control parity mask),     fc = parity(b0 & rc)
// f / fc are constants that flip the logical bit (f = 1 swaps the roles of the pair's amplitudes);
// they are stored in meta bits 28 / 29 and are zero for the all-zero start state of the chemistry
// searches.  An X mask x of the Hamiltonian lies inside D or drops out of the energy and of every
// derivative; inside, its coordinates x' say which compressed amplitude it pairs with.
// No reference counterpart.  The evaluation kernels do not use this yet: tests/test_host.py runs the
// translated tape through the NumPy interpreter (tests/compressed_register.py) against the oracle.
#include <cstdint>
#include <vector>

#include "../../include/qas_b200.h"
#include "tape.h"

#define QAS_META_FLIP (1u << 28)
#define QAS_META_CFLIP (1u << 29)

namespace {
struct SpanBasis {
  // reduced echelon basis, highest-bit pivots; vec[k] sorted by ascending pivot
  std::vector<uint32_t> vec;
  std::vector<int> piv;
  bool add(uint32_t x) {
    for (size_t k = 0; k < vec.size(); ++k) if ((x >> piv[k]) & 1u) x ^= vec[k];
    if (!x) return false;
    const int p = 31 - __builtin_clz(x);
    for (size_t k = 0; k < vec.size(); ++k) if ((vec[k] >> p) & 1u) vec[k] ^= x;
    size_t at = 0;
    while (at < vec.size() && piv[at] < p) ++at;
    vec.insert(vec.begin() + at, x);
    piv.insert(piv.begin() + at, p);
    return true;
  }
  // coordinates of x (bit k <-> vec[k]); false if x is outside the span
  bool coords(uint32_t x, uint32_t *t) const {
    uint32_t c = 0;
    for (size_t k = 0; k < vec.size(); ++k) if ((x >> piv[k]) & 1u) { x ^= vec[k]; c |= 1u << k; }
    *t = c;
    return x == 0;
  }
  uint32_t dual(uint32_t r) const {      // r'_k = parity(B_k & r)
    uint32_t c = 0;
    for (size_t k = 0; k < vec.size(); ++k) c |= (uint32_t)(__builtin_popcount(vec[k] & r) & 1) << k;
    return c;
  }
};
}  // namespace

extern "C" int qas_compress_tape_host(int n_qubits, uint32_t *ops, int num_ops, uint32_t init_index,
                                      const uint32_t *xmasks, int num_xmasks, int *dim_out,
                                      uint32_t *basis_out, uint32_t *xcoords_out) {
  if ((num_ops > 0 && !ops) || num_ops < 0 || n_qubits < 1 || n_qubits > QAS_MAX_QUBITS || !dim_out ||
      !basis_out || (num_xmasks > 0 && (!xmasks || !xcoords_out)) || num_xmasks < 0)
    return QAS_ERR_INVALID;
  SpanBasis B;
  for (int j = 0; j < num_ops; ++j) {
    const uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
    if (qas_meta_kind(o[4]) == QAS_K_U1) B.add(o[0]);
  }
  const int d = (int)B.vec.size();
  *dim_out = d;
  for (int k = 0; k < d; ++k) basis_out[k] = B.vec[k];
  for (int j = 0; j < num_ops; ++j) {
    uint32_t *o = ops + (size_t)j * QAS_OP_WORDS;
    uint32_t mp = 0;
    if (qas_meta_kind(o[4]) == QAS_K_U1 && !B.coords(o[0], &mp)) return QAS_ERR_INVALID;   // cannot happen
    const uint32_t r = o[1], rc = o[2];
    uint32_t meta = o[4] & ~(QAS_META_FLIP | QAS_META_CFLIP);
    if (__builtin_popcount(init_index & r) & 1) meta |= QAS_META_FLIP;
    if (rc && (__builtin_popcount(init_index & rc) & 1)) meta |= QAS_META_CFLIP;
    o[0] = mp;
    o[1] = B.dual(r);
    o[2] = B.dual(rc);
    o[4] = meta;
    // a control whose parity mask vanishes on D is constant on the support: encode it so that the
    // interpreter can tell "no control" (rc == 0 upstream) from "control parity = fc everywhere"
    if (rc && o[2] == 0u) o[2] = 0x80000000u;
  }
  for (int s = 0; s < num_xmasks; ++s) {
    uint32_t t = 0;
    xcoords_out[s] = B.coords(xmasks[s], &t) ? t : 0xffffffffu;
  }
  return QAS_OK;
}
