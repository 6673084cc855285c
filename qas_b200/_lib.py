"""ctypes binding of libqas_b200.so (the C ABI declared in include/qas_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
CPU fallback: if the shared object is missing this module raises at first use.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QAS_B200_LIB", os.path.join(_HERE, "lib", "libqas_b200.so"))

QAS_OK = 0
QAS_ERR_INVALID = 1
QAS_ERR_CUDA = 2
QAS_ERR_UNSUPPORTED = 3
QAS_ERR_NOMEM = 4

QAS_INIT_ZERO = 0
QAS_INIT_BASIS = 1
QAS_INIT_PLUS = 2

QAS_F_DEVICE_PTRS = 0x1
QAS_F_GRAD_SUM = 0x2
QAS_F_PARAMS_RESIDENT = 0x4
QAS_F_K_U8 = 0x8

QAS_OP_WORDS = 5

# order must match enum qas_gate_id in include/qas_b200.h
GATE_IDS = {name: i for i, name in enumerate([
    "PlaceHolder", "Hadamard", "PauliX", "PauliY", "PauliZ", "S", "Sdg", "T", "Tdg", "SX",
    "Rot", "RX", "RY", "RZ", "PhaseShift", "U1", "U2", "U3",
    "CNOT", "CZ", "CY", "SWAP", "ISWAP", "SISWAP",
    "CRX", "CRY", "CRZ", "CRot", "CPhase",
    "IsingXX", "IsingYY", "IsingZZ",
])}
GATE_IDS["SQISW"] = GATE_IDS["SISWAP"]
GATE_IDS["ControlledPhaseShift"] = GATE_IDS["CPhase"]


class PoolEntry(ctypes.Structure):
    _fields_ = [("gate_id", ctypes.c_int32), ("wire0", ctypes.c_int32), ("wire1", ctypes.c_int32)]


class PauliTerm(ctypes.Structure):
    _fields_ = [("xmask", ctypes.c_uint64), ("zmask", ctypes.c_uint64), ("coeff", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [("evals", ctypes.c_uint64), ("launches", ctypes.c_uint64),
                ("last_kernel_ms", ctypes.c_double), ("last_eval_kernel_ms", ctypes.c_double),
                ("teams_per_cta", ctypes.c_int),
                ("threads_per_team", ctypes.c_int), ("ctas_per_sm", ctypes.c_int),
                ("smem_bytes_per_cta", ctypes.c_int), ("path", ctypes.c_int)]


class SvGate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("target", ctypes.c_int32), ("control", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("mask", ctypes.c_uint64), ("m", ctypes.c_double * 8)]


class SvInfo(ctypes.Structure):
    _fields_ = [("passes", ctypes.c_int), ("tile_bits", ctypes.c_int), ("bytes_per_pass", ctypes.c_double),
                ("last_ms", ctypes.c_double), ("tma_passes", ctypes.c_int)]


QAS_SV_U1 = 0
QAS_SV_PDIAG = 1

_lib = None

# every symbol include/qas_b200.h declares (tests check the export list against this)
EXPORTS = [
    "qas_abi_version", "qas_build_info", "qas_ctx_create", "qas_ctx_destroy", "qas_last_error",
    "qas_eval_batch", "qas_compile_tape", "qas_compile_tape_host", "qas_plan_passes_host",
    "qas_plan_hsupport_host", "qas_plan_tile_pass_host", "qas_tile_group_basis_host", "qas_compile_ctape_host", "qas_check_keys_host",
    "qas_get_stats", "qas_debug_phase_clocks", "qas_debug_phase_clocks_c", "qas_debug_phase_clocks_t", "qas_ctx_param_slots", "qas_adam_step", "qas_set_params",
    "qas_fp64_peak_tflops", "qas_reduce_rows", "qas_peer_push_row", "qas_peer_wait_reduce", "qas_batch_mean", "qas_vqls_combine", "qas_tuning_tick", "qas_adam_step_dev",
]
# include/qas_b200_sv.h
SV_EXPORTS = ["qas_sv_init", "qas_sv_apply", "qas_sv_plan", "qas_sv_expval", "qas_sv_norm2", "qas_sv_last_error"]


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libqas_b200.so not found at %s -- build it with `python -c \"import __graft_entry__ as g; "
            "g.build()\"`; there is no CPU fallback for the evaluator" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    lib.qas_abi_version.restype = ctypes.c_int
    lib.qas_build_info.restype = ctypes.c_char_p
    lib.qas_last_error.restype = ctypes.c_char_p
    lib.qas_last_error.argtypes = [vp]
    lib.qas_ctx_create.restype = ctypes.c_int
    lib.qas_ctx_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_uint64, ctypes.POINTER(PauliTerm), ctypes.c_int,
                                   ctypes.POINTER(PoolEntry), ctypes.c_int, ctypes.c_int]
    lib.qas_ctx_destroy.restype = ctypes.c_int
    lib.qas_ctx_destroy.argtypes = [vp]
    lib.qas_eval_batch.restype = ctypes.c_int
    lib.qas_eval_batch.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp,
                                   ctypes.c_uint32, vp]
    lib.qas_compile_tape.restype = ctypes.c_int
    lib.qas_compile_tape.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint32)]
    lib.qas_compile_tape_host.restype = ctypes.c_int
    lib.qas_compile_tape_host.argtypes = [ctypes.c_int, ctypes.POINTER(PoolEntry), ctypes.c_int,
                                          ctypes.c_int, vp, ctypes.c_uint64, vp, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_uint32)]
    lib.qas_plan_passes_host.restype = ctypes.c_int
    lib.qas_plan_passes_host.argtypes = [ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                         ctypes.c_int, vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.qas_plan_tile_pass_host.restype = ctypes.c_int
    lib.qas_plan_tile_pass_host.argtypes = [ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, vp, vp, ctypes.POINTER(ctypes.c_int)]
    lib.qas_check_keys_host.restype = ctypes.c_int
    lib.qas_check_keys_host.argtypes = [vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_int)]
    lib.qas_tile_group_basis_host.restype = ctypes.c_int
    lib.qas_tile_group_basis_host.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_int)]
    lib.qas_plan_hsupport_host.restype = ctypes.c_int
    lib.qas_plan_hsupport_host.argtypes = [ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int,
                                           vp, vp, vp, ctypes.c_int]
    pi = ctypes.POINTER(ctypes.c_int)
    lib.qas_compile_ctape_host.restype = ctypes.c_int
    lib.qas_compile_ctape_host.argtypes = [ctypes.c_int, ctypes.POINTER(PoolEntry), ctypes.c_int, ctypes.c_int, vp,
                                           ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int,
                                           pi, pi, pi, vp, vp, pi, ctypes.c_int]
    lib.qas_set_params.restype = ctypes.c_int
    lib.qas_set_params.argtypes = [vp, ctypes.c_int, vp]
    lib.qas_adam_step.restype = ctypes.c_int
    lib.qas_adam_step.argtypes = [vp, vp, vp, vp, vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
    lib.qas_reduce_rows.restype = ctypes.c_int
    lib.qas_reduce_rows.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double, vp, vp]
    lib.qas_peer_push_row.restype = ctypes.c_int
    lib.qas_peer_push_row.argtypes = [vp, ctypes.c_uint64, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.c_int, ctypes.c_uint64, vp]
    lib.qas_peer_wait_reduce.restype = ctypes.c_int
    lib.qas_peer_wait_reduce.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double, vp, vp,
                                         ctypes.c_uint64, vp, vp]
    lib.qas_tuning_tick.restype = ctypes.c_int
    lib.qas_tuning_tick.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
    lib.qas_adam_step_dev.restype = ctypes.c_int
    lib.qas_adam_step_dev.argtypes = [vp, vp, vp, vp, ctypes.c_uint64, vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, vp]
    lib.qas_batch_mean.restype = ctypes.c_int
    lib.qas_batch_mean.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_uint32, vp]
    lib.qas_vqls_combine.restype = ctypes.c_int
    lib.qas_vqls_combine.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.qas_debug_phase_clocks.restype = ctypes.c_int
    lib.qas_debug_phase_clocks.argtypes = [vp, vp, ctypes.c_int]
    lib.qas_debug_phase_clocks_c.restype = ctypes.c_int
    lib.qas_debug_phase_clocks_c.argtypes = [vp, vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.qas_get_stats.restype = ctypes.c_int
    lib.qas_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    lib.qas_fp64_peak_tflops.restype = ctypes.c_int
    lib.qas_fp64_peak_tflops.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.qas_sv_last_error.restype = ctypes.c_char_p
    lib.qas_sv_init.restype = ctypes.c_int
    lib.qas_sv_init.argtypes = [vp, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_double, vp]
    lib.qas_sv_apply.restype = ctypes.c_int
    lib.qas_sv_apply.argtypes = [vp, ctypes.c_int, ctypes.POINTER(SvGate), ctypes.c_int, ctypes.c_int, vp,
                                 ctypes.c_int, ctypes.POINTER(SvInfo)]
    lib.qas_sv_plan.restype = ctypes.c_int
    lib.qas_sv_plan.argtypes = [ctypes.c_int, ctypes.POINTER(SvGate), ctypes.c_int, ctypes.c_int,
                                ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
    lib.qas_sv_expval.restype = ctypes.c_int
    lib.qas_sv_expval.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(PauliTerm),
                                  ctypes.c_int, ctypes.POINTER(ctypes.c_double), vp]
    lib.qas_sv_norm2.restype = ctypes.c_int
    lib.qas_sv_norm2.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double), vp]
    _lib = lib
    return lib
