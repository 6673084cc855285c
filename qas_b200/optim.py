"""Adam with the update rule of ``qml.AdamOptimizer`` -- the optimiser every reference script passes
to ``search`` / ``circuitModelTuning`` (qas/mcts/mcts.py:226,351,387,407).  PennyLane is not a
dependency here; the rule (PennyLane 0.29 documentation) is

    m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g*g ;  t <- t+1
    x <- x - stepsize * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)

with b1 = 0.9, b2 = 0.99, eps = 1e-8.  ``apply_grad(grad, args)`` mirrors the call shape used
upstream (``optimizer.apply_grad(grad=..., args=params)``)."""
import numpy as np


class AdamOptimizer:
    def __init__(self, stepsize=0.01, beta1=0.9, beta2=0.99, eps=1e-8):
        self.stepsize, self.beta1, self.beta2, self.eps = stepsize, beta1, beta2, eps
        self.fm = None
        self.sm = None
        self.t = 0

    def apply_grad(self, grad, args):
        x = np.asarray(args, dtype=np.float64)
        g = np.asarray(grad, dtype=np.float64)
        if self.fm is None:
            self.fm = np.zeros_like(x)
            self.sm = np.zeros_like(x)
        self.t += 1
        self.fm = self.beta1 * self.fm + (1.0 - self.beta1) * g
        self.sm = self.beta2 * self.sm + (1.0 - self.beta2) * g * g
        step = self.stepsize * np.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)
        return x - step * self.fm / (np.sqrt(self.sm) + self.eps)

    def reset(self):
        self.fm = self.sm = None
        self.t = 0


class DeviceAdam:
    """The same rule on the GPU (libqas_b200: qas_adam_step): parameters, moments and the gradient
    stay in HBM, so an optimisation loop over a fixed circuit never moves them over PCIe.
    ``params`` is a CUDA float64 tensor updated in place by ``step``."""

    def __init__(self, params, stepsize=0.01, beta1=0.9, beta2=0.99, eps=1e-8):
        import torch
        from qas_b200 import _lib
        assert params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
        self._lib = _lib.load()
        self.params = params
        self.stepsize, self.beta1, self.beta2, self.eps = stepsize, beta1, beta2, eps
        self.fm = torch.zeros_like(params)
        self.sm = torch.zeros_like(params)
        self.t = 0

    def step(self, grad, noise=None, noise_factor=0.0, stream=None):
        """grad (and noise, if given): CUDA float64 tensors of params' shape."""
        import ctypes
        assert grad.is_cuda and grad.is_contiguous() and grad.numel() == self.params.numel()
        self.t += 1
        sp = ctypes.c_void_p(stream.cuda_stream) if stream is not None else None
        rc = self._lib.qas_adam_step(self.params.data_ptr(), grad.data_ptr(), self.fm.data_ptr(), self.sm.data_ptr(),
                                     noise.data_ptr() if noise is not None else None, self.params.numel(), self.t,
                                     self.stepsize, self.beta1, self.beta2, self.eps, float(noise_factor), sp)
        if rc != 0:
            raise RuntimeError("qas_adam_step failed (%d): %s" % (rc, self._lib.qas_last_error(None).decode()))
        return self.params
