"""qas_b200 -- B200-native batched evaluator behind the QAS model/pool plugin surface.

Only the hot path of peiyong-addwater/QAS is here: rewards and
adjoint gradients of batches of candidate circuits, computed by hand-written sm_100a CUDA
kernels through the C ABI in include/qas_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
