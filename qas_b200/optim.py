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
