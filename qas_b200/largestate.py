"""Large-state mode: one statevector of 21-34 qubits, resident in HBM, optionally sharded over
2/4/8 GPUs (BASELINE.json config 5: "synthetic 26-30-qubit hardware-efficient ansatz from the
qas gate set, one statevector sharded across 2/4/8 B200 with NCCL qubit swaps").

The reference has no counterpart (largest model: 16 qubits on one CPU,
qas/qml_models/kagome_heisenberg_model.py:62); the gate vocabulary and conventions are the
reference's (qas/qml_models/qml_gate_ops.py:66-105, wire 0 = most significant bit).

Layout: with W = 2^g ranks, the top g bit POSITIONS of the n-bit amplitude index are the rank;
each rank owns a slab of 2^(n-g) complex128 amplitudes.  A map logical wire -> bit position is
kept on the host.  Gates whose target sits on a local position run in shared-memory-tiled HBM
passes (libqas_b200: qas_sv_apply); controls and diagonal phases on rank bits need no data
movement (they are resolved per rank).  When the next gates need a target that is currently a
rank bit, ONE all-to-all (torch.distributed.all_to_all_single over NCCL/NVLink) swaps the g rank
positions with the top g local positions -- the slab is already contiguous in those bits, so no
pack/unpack pass is needed -- and the wire map is updated.  A dependency-aware scheduler runs
every gate that is executable before paying for a swap.
"""
import ctypes
import math

import numpy as np

from qas_b200 import _lib
from qas_b200.hamiltonians import word_to_masks

_R = 1.0 / math.sqrt(2.0)


# ----------------------------------------------------------------------------- gate lowering
def _rz(t):
    return np.array([[np.exp(-0.5j * t), 0], [0, np.exp(0.5j * t)]])


def _ry(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def _rx(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]])


def _ps(t):
    return np.array([[1, 0], [0, np.exp(1j * t)]])


_H = np.array([[_R, _R], [_R, -_R]], dtype=np.complex128)
_S = np.array([[1, 0], [0, 1j]])
_CONST = {
    "Hadamard": _H, "PauliX": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "PauliY": np.array([[0, -1j], [1j, 0]]), "PauliZ": np.array([[1, 0], [0, -1]], dtype=np.complex128),
    "S": _S, "Sdg": _S.conj(), "T": np.array([[1, 0], [0, np.exp(0.25j * math.pi)]]),
    "Tdg": np.array([[1, 0], [0, np.exp(-0.25j * math.pi)]]),
    "SX": 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]),
}
_X = _CONST["PauliX"]


def _matrix_1q(name, th):
    if name in _CONST:
        return _CONST[name]
    if name == "Rot":
        return _rz(th[2]) @ _ry(th[1]) @ _rz(th[0])
    if name == "RX":
        return _rx(th[0])
    if name == "RY":
        return _ry(th[0])
    if name == "RZ":
        return _rz(th[0])
    if name in ("PhaseShift", "U1"):
        return _ps(th[0])
    if name == "U2":
        return _R * np.array([[1, -np.exp(1j * th[1])], [np.exp(1j * th[0]), np.exp(1j * (th[0] + th[1]))]])
    if name == "U3":
        c, s = math.cos(th[0] / 2), math.sin(th[0] / 2)
        return np.array([[c, -np.exp(1j * th[2]) * s], [np.exp(1j * th[1]) * s, np.exp(1j * (th[1] + th[2])) * c]])
    raise ValueError("unsupported gate %s" % name)


def lower_gate(name, wires, th):
    """One vocabulary gate -> primitive ops on LOGICAL wires:
    ("u1", target, control|None, 2x2)  |  ("pdiag", [wires], d0, d1)."""
    th = list(th) if th is not None else []
    w = list(wires)
    if name == "PlaceHolder":
        return []
    if len(w) == 1:
        return [("u1", w[0], None, _matrix_1q(name, th))]
    a, b = w
    if name == "CNOT":
        return [("u1", b, a, _X)]
    if name in ("CZ", "CY"):
        return [("u1", b, a, _CONST["Pauli" + name[1]])]
    if name in ("CRX", "CRY", "CRZ", "CRot", "CPhase", "ControlledPhaseShift"):
        inner = {"CRX": "RX", "CRY": "RY", "CRZ": "RZ", "CRot": "Rot"}.get(name, "PhaseShift")
        return [("u1", b, a, _matrix_1q(inner, th))]
    if name == "SWAP":
        return [("u1", b, a, _X), ("u1", a, b, _X), ("u1", b, a, _X)]
    zz = lambda t: ("pdiag", [a, b], np.exp(-0.5j * t), np.exp(0.5j * t))  # noqa: E731
    if name == "IsingZZ":
        return [zz(th[0])]
    had = [("u1", a, None, _H), ("u1", b, None, _H)]
    sh, hsd = _S @ _H, _H @ _S.conj()
    pre = [("u1", a, None, hsd), ("u1", b, None, hsd)]
    post = [("u1", a, None, sh), ("u1", b, None, sh)]
    if name == "IsingXX":
        return had + [zz(th[0])] + had
    if name == "IsingYY":
        return pre + [zz(th[0])] + post
    if name == "ISWAP":
        return lower_gate("SWAP", w, None) + [("u1", b, a, _CONST["PauliZ"]), ("u1", a, None, _S), ("u1", b, None, _S)]
    if name in ("SISWAP", "SQISW"):
        t = -math.pi / 4
        return pre + [zz(t)] + post + had + [zz(t)] + had
    raise ValueError("unsupported gate %s" % name)


def hea_gate_list(n, layers, seed=0):
    """Hardware-efficient ansatz of SURVEY.md 8(d): per layer Rot on every wire, then the
    nearest-neighbour CNOT ladder (0,1),(1,2),...; angles from default_rng(seed)."""
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(layers):
        for w in range(n):
            gates.append(("Rot", [w], list(rng.standard_normal(3))))
        for w in range(n - 1):
            gates.append(("CNOT", [w, w + 1], None))
    return gates


def zz_chain_terms(n):
    """sum_i Z_i Z_{i+1} as [(coeff, word)]."""
    out = []
    for i in range(n - 1):
        w = ["I"] * n
        w[i] = w[i + 1] = "Z"
        out.append((1.0, "".join(w)))
    return out


# ----------------------------------------------------------------------------- CUDA backend
class CudaSvBackend:
    """Slab kernels of libqas_b200 on torch CUDA tensors (torch = device memory + NCCL only)."""

    def __init__(self, device=0, tile_bits=0):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.tile_bits = tile_bits
        self.passes = 0
        self.tma_passes = 0
        self.kernel_ms = 0.0
        self.bytes = 0.0
        self._swap_events = []

    def alloc(self, n_local):
        return self.torch.empty((1 << n_local, 2), dtype=self.torch.float64, device=self.device)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        if rc != _lib.QAS_OK:
            raise RuntimeError("libqas_b200 (sv) error %d: %s" % (rc, self.lib.qas_sv_last_error().decode()))

    def init(self, state, n_local, index, amp):
        self._check(self.lib.qas_sv_init(state.data_ptr(), n_local, int(index), float(amp.real), float(amp.imag),
                                         self._stream()))

    def apply(self, state, n_local, gates):
        """gates: [(kind, target, control, mask, 2x2 complex)] on local bit positions."""
        if not gates:
            return
        arr = (_lib.SvGate * len(gates))()
        for i, (kind, target, control, mask, m) in enumerate(gates):
            arr[i].kind, arr[i].target, arr[i].control, arr[i].mask = kind, target, control, mask
            flat = np.asarray(m, dtype=np.complex128).reshape(4)
            for q in range(4):
                arr[i].m[2 * q], arr[i].m[2 * q + 1] = flat[q].real, flat[q].imag
        info = _lib.SvInfo()
        self._check(self.lib.qas_sv_apply(state.data_ptr(), n_local, arr, len(gates), self.tile_bits,
                                          self._stream(), 1, ctypes.byref(info)))
        self.passes += info.passes
        self.tma_passes += info.tma_passes
        self.kernel_ms += info.last_ms
        self.bytes += info.passes * info.bytes_per_pass

    def expval(self, state, n_local, n_total, rank_bits, masks):
        arr = (_lib.PauliTerm * len(masks))(*[_lib.PauliTerm(x, z, c) for x, z, c in masks])
        out = ctypes.c_double(0.0)
        self._check(self.lib.qas_sv_expval(state.data_ptr(), n_local, n_total, int(rank_bits), arr, len(masks),
                                           ctypes.byref(out), self._stream()))
        return out.value

    def all_to_all(self, out, inp, group):
        import torch.distributed as dist
        e0 = self.torch.cuda.Event(enable_timing=True)
        e1 = self.torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_to_all_single(out.view(-1), inp.view(-1), group=group)
        e1.record()
        self._swap_events.append((e0, e1))

    def swap_ms(self, reset=True):
        """device time of the rank-bit swaps since the last call (synchronises)"""
        self.torch.cuda.synchronize(self.device)
        ms = sum(a.elapsed_time(b) for a, b in self._swap_events)
        if reset:
            self._swap_events = []
        return ms

    def all_reduce_sum(self, value, group):
        import torch.distributed as dist
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t, group=group)
        return float(t.item())


# ----------------------------------------------------------------------------- orchestration
class LargeStateSimulator:
    def __init__(self, n_qubits, backend=None, group=None, device=0, tile_bits=0):
        import torch.distributed as dist
        self.n = int(n_qubits)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.g = int(round(math.log2(self.world)))
        assert 1 << self.g == self.world, "world size must be a power of two"
        self.n_local = self.n - self.g
        assert self.n_local >= 2 * self.g, "slab too small for the rank-bit swap"
        self.backend = backend if backend is not None else CudaSvBackend(device, tile_bits)
        self.state = self.backend.alloc(self.n_local)
        self.scratch = self.backend.alloc(self.n_local) if self.world > 1 else None
        self.swaps = 0
        self.reset()

    # wire w -> bit position; initially wire 0 is the most significant bit (reference convention)
    def reset(self, init=("zero", 0)):
        self.pos = [self.n - 1 - w for w in range(self.n)]
        kind, arg = init
        if kind == "plus":
            self.backend.init(self.state, self.n_local, -1, complex(2.0 ** (-0.5 * self.n)))
        else:
            index = int(arg) if kind == "basis" else 0
            hi, lo = index >> self.n_local, index & ((1 << self.n_local) - 1)
            self.backend.init(self.state, self.n_local, lo if hi == self.rank else (1 << self.n_local), 1.0 + 0j)
            # an index beyond the slab sets nothing: ranks that do not own the basis state stay zero
        return self

    def _swap_rank_bits(self):
        """all-to-all: rank positions [n_local, n) <-> top local positions [n_local-g, n_local)."""
        self.backend.all_to_all(self.scratch, self.state, self.group)
        self.state, self.scratch = self.scratch, self.state
        lo, g = self.n_local - self.g, self.g
        self.pos = [p + g if lo <= p < self.n_local else (p - g if p >= self.n_local else p) for p in self.pos]
        self.swaps += 1

    def _resolve(self, op):
        """primitive op on logical wires -> gate on local bit positions for THIS rank (or None)."""
        nl = self.n_local
        if op[0] == "u1":
            _, t, ctl, m = op
            pt = self.pos[t]
            control = -1
            if ctl is not None:
                pc = self.pos[ctl]
                if pc >= nl:
                    if not (self.rank >> (pc - nl)) & 1:
                        return None
                else:
                    control = pc
            return (_lib.QAS_SV_U1, pt, control, 0, m)
        _, wires, d0, d1 = op
        mask, flip = 0, 0
        for w in wires:
            p = self.pos[w]
            if p >= nl:
                flip ^= (self.rank >> (p - nl)) & 1
            else:
                mask |= 1 << p
        if flip:
            d0, d1 = d1, d0
        return (_lib.QAS_SV_PDIAG, 0, -1, mask, np.array([[d0, 0], [0, d1]]))

    def run(self, gate_list):
        """gate_list: [(name, wires, params or None)] -- the reference's toList format."""
        pending = []
        for name, wires, th in gate_list:
            pending.extend(lower_gate(name, wires, th))
        while pending:
            batch, rest, blocked = [], [], set()
            for op in pending:
                qs = ([op[1]] + ([op[2]] if op[2] is not None else [])) if op[0] == "u1" else list(op[1])
                if blocked.intersection(qs) or (op[0] == "u1" and self.pos[op[1]] >= self.n_local):
                    rest.append(op)
                    blocked.update(qs)
                else:
                    batch.append(op)
            local = [g for g in (self._resolve(op) for op in batch) if g is not None]
            self.backend.apply(self.state, self.n_local, local)
            pending = rest
            if pending:
                assert self.world > 1
                self._swap_rank_bits()
        return self

    def expval(self, terms):
        """sum_t c_t <psi|P_t|psi> over the whole (sharded) register; identical on every rank."""
        def masks():
            out = []
            for c, word in terms:
                x = z = 0
                for w, ch in enumerate(word):
                    bit = 1 << self.pos[w]
                    if ch in "XY":
                        x |= bit
                    if ch in "ZY":
                        z |= bit
                out.append((x, z, float(c)))
            return out
        mk = masks()
        local_mask = (1 << self.n_local) - 1
        if any(x & ~local_mask for x, _, _ in mk):
            self._swap_rank_bits()
            mk = masks()
            if any(x & ~local_mask for x, _, _ in mk):
                raise NotImplementedError("observable has X/Y on both the rank wires and the top local wires")
        e = self.backend.expval(self.state, self.n_local, self.n, self.rank, mk)
        return self.backend.all_reduce_sum(e, self.group) if self.world > 1 else e
