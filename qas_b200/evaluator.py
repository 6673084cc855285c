"""BatchEvaluator -- Python face of the C ABI (include/qas_b200.h).

One evaluator = one (n qubits, initial state, Hamiltonian, operator pool) family, i.e. what a
reference model class fixes at class level.  ``evaluate`` takes a batch of structure lists
that share one parameter super-set and returns energies plus gradients, replacing the
serial per-circuit loops of the reference search driver (qas/mcts/mcts.py:298-305, :340-347).

Host arrays go through ``qas_eval_batch`` with host pointers (the library stages them and
returns synchronously); ``evaluate_device`` passes CUDA tensors' data pointers and is
asynchronous on the given stream.  PyTorch is used only as a container for device memory.
"""
import ctypes

import numpy as np

from qas_b200 import _lib
from qas_b200.hamiltonians import pauli_terms
from qas_b200.qml_models.qml_gate_ops import pool_table

_INIT = {"zero": _lib.QAS_INIT_ZERO, "basis": _lib.QAS_INIT_BASIS, "plus": _lib.QAS_INIT_PLUS}


class QasError(RuntimeError):
    pass


class BatchEvaluator:
    def __init__(self, n_qubits, terms, op_pool, init=("zero", 0), l=3, device=0):
        lib = _lib.load()
        self._lib = lib
        self.n_qubits = int(n_qubits)
        self.l = int(l)
        self.device = int(device)
        table = pool_table(op_pool)
        self.c = len(table)
        pool_arr = (_lib.PoolEntry * self.c)(*[_lib.PoolEntry(*e) for e in table])
        pt = pauli_terms(terms)
        term_arr = (_lib.PauliTerm * len(pt))(*[_lib.PauliTerm(x, z, c) for x, z, c in pt])
        kind, arg = init
        h = ctypes.c_void_p()
        rc = lib.qas_ctx_create(ctypes.byref(h), self.device, self.n_qubits, _INIT[kind], int(arg),
                                term_arr, len(pt), pool_arr, self.c, self.l)
        if rc != _lib.QAS_OK:
            raise QasError("qas_ctx_create failed (%d): %s" % (rc, lib.qas_last_error(None).decode()))
        self._h = h

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.qas_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == _lib.QAS_OK:
            return
        msg = self._lib.qas_last_error(self._h).decode()
        if rc == _lib.QAS_ERR_INVALID:
            # the reference raises AssertionError on the same conditions (utils.py:12-15)
            raise AssertionError(msg)
        raise QasError("libqas_b200 error %d: %s" % (rc, msg))

    # ------------------------------------------------------------------ host-array entry
    def evaluate(self, ks, params, want_grad=True, dense=True, compact=False, avg=True, out=None):
        """ks: int array [B, p]; params: float64 [p, c, l].  A uint8 array (pools of <= 256 entries) is uploaded
        as it is -- a quarter of the bytes -- and widened on the device (QAS_F_K_U8).
        out: optional dict of preallocated C-contiguous float64 result arrays ("energy" [B], "grad_dense" [p,c,l],
        "grad_compact" [B,p,l]) -- e.g. page-locked memory, so that the read-back is a true asynchronous DMA instead
        of the driver's staged copy into fresh pageable pages.
        Returns dict(energy[B], grad_dense[p,c,l] | None, grad_compact[B,p,l] | None)."""
        ks = np.asarray(ks)
        k8 = ks.dtype == np.uint8 and self.c <= 256
        ks = np.ascontiguousarray(ks if k8 else ks.astype(np.int32, copy=False))
        if ks.ndim == 1:
            ks = ks[None, :]
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        B, p = ks.shape
        assert params.shape == (p, self.c, self.l), (params.shape, (p, self.c, self.l))
        def _buf(name, shape):
            a = out.get(name) if out else None
            if a is None:
                return np.empty(shape, dtype=np.float64)
            assert a.dtype == np.float64 and a.shape == tuple(shape) and a.flags["C_CONTIGUOUS"], name
            return a
        energy = _buf("energy", (B,))
        gd = _buf("grad_dense", (p, self.c, self.l)) if (want_grad and dense) else None
        gc = _buf("grad_compact", (B, p, self.l)) if (want_grad and compact) else None
        flags = (0 if avg else _lib.QAS_F_GRAD_SUM) | (_lib.QAS_F_K_U8 if k8 else 0)
        rc = self._lib.qas_eval_batch(
            self._h, B, p, ks.ctypes.data, params.ctypes.data, energy.ctypes.data,
            gc.ctypes.data if gc is not None else None,
            gd.ctypes.data if gd is not None else None, flags, None)
        self._check(rc)
        return {"energy": energy, "grad_dense": gd, "grad_compact": gc}

    # ------------------------------------------------------------------ device-tensor entry
    def evaluate_device(self, ks, params, energy, grad_compact=None, grad_dense=None, avg=True,
                        stream=None):
        """All arguments are CUDA tensors on this evaluator's device (int32 / float64,
        contiguous); outputs are written in place, asynchronously on `stream` (a
        torch.cuda.Stream or None for the context's own stream)."""
        B, p = ks.shape
        assert ks.is_cuda and params.is_cuda and energy.is_cuda
        assert ks.is_contiguous() and params.is_contiguous() and energy.is_contiguous()
        assert tuple(params.shape) == (p, self.c, self.l)
        assert energy.numel() == B
        import torch
        flags = _lib.QAS_F_DEVICE_PTRS | (0 if avg else _lib.QAS_F_GRAD_SUM) | (_lib.QAS_F_K_U8 if ks.dtype == torch.uint8 else 0)
        sp = ctypes.c_void_p(stream.cuda_stream) if stream is not None else None
        rc = self._lib.qas_eval_batch(
            self._h, B, p, ks.data_ptr(), params.data_ptr(), energy.data_ptr(),
            grad_compact.data_ptr() if grad_compact is not None else None,
            grad_dense.data_ptr() if grad_dense is not None else None, flags, sp)
        self._check(rc)

    # ------------------------------------------------------------------ introspection
    def compile_tape(self, k):
        """Host-side run of the tape compiler on one structure list ->
        (ops uint32[num_ops, 5] in last-gate-first order, physical index of the initial basis state)."""
        k = np.ascontiguousarray(np.asarray(k, dtype=np.int32))
        cap = max(1, 10 * len(k))
        ops = np.zeros((cap, _lib.QAS_OP_WORDS), dtype=np.uint32)
        n = ctypes.c_int(0)
        idx = ctypes.c_uint32(0)
        self._check(self._lib.qas_compile_tape(self._h, len(k), k.ctypes.data, ops.ctypes.data, cap,
                                               ctypes.byref(n), ctypes.byref(idx)))
        return ops[:n.value].copy(), idx.value

    def stats(self):
        s = _lib.Stats()
        self._check(self._lib.qas_get_stats(self._h, ctypes.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}


def compile_tape_host(n_qubits, op_pool, k, init_basis=0):
    """Tape compiler without a GPU/context (used by CPU tests and tooling)."""
    lib = _lib.load()
    table = pool_table(op_pool)
    pool_arr = (_lib.PoolEntry * len(table))(*[_lib.PoolEntry(*e) for e in table])
    k = np.ascontiguousarray(np.asarray(k, dtype=np.int32))
    cap = max(1, 10 * len(k))
    ops = np.zeros((cap, _lib.QAS_OP_WORDS), dtype=np.uint32)
    n = ctypes.c_int(0)
    idx = ctypes.c_uint32(0)
    rc = lib.qas_compile_tape_host(n_qubits, pool_arr, len(table), len(k), k.ctypes.data,
                                   int(init_basis), ops.ctypes.data, cap, ctypes.byref(n),
                                   ctypes.byref(idx))
    if rc == _lib.QAS_ERR_INVALID:
        raise AssertionError(lib.qas_last_error(None).decode())
    if rc != _lib.QAS_OK:
        raise QasError("qas_compile_tape_host failed (%d)" % rc)
    return ops[:n.value].copy(), idx.value


def fp64_peak_tflops(device=0, reps=5):
    lib = _lib.load()
    out = ctypes.c_double(0.0)
    rc = lib.qas_fp64_peak_tflops(device, reps, ctypes.byref(out))
    if rc != _lib.QAS_OK:
        raise QasError("qas_fp64_peak_tflops failed (%d): %s" % (rc, lib.qas_last_error(None).decode()))
    return out.value
