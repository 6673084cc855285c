"""qas_b200 -- B200-native batched evaluator behind the QAS model/pool plugin surface.

Only the hot path of peiyong-addwater/QAS is here: rewards and
adjoint gradients of batches of candidate circuits, computed by hand-written sm_100a CUDA
kernels through the C ABI in include/qas_b200.h.  There is no CPU fallback.
"""
import os as _os

# The evaluator launches one kernel per support bin on forked streams (csrc/api.cu); with the driver's default of 8
# hardware work queues those streams alias with each other and with NCCL's as soon as a process group exists, and the
# bins serialise (measured: 0.83 -> 0.90 ms per LiH step under torchrun).  The driver reads this variable when the
# CUDA context is created, so it has to be in the environment before the first CUDA call of the process.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__version__ = "0.1.0"
