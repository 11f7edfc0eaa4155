"""Flux.Adam / update! mirror ([EXT Flux 0.14.22]; LV_driver_KANODE.jl:219,287; PDE scripts `ADAM(1e-2)`).

The reference keeps `p` on the host and updates 240 numbers per iteration; `Adam.update` does the same on a numpy
vector.  For device-resident training (large PDE models, data-parallel) `Adam.update_dev` runs the fused kernel
`kanode_adam_step_dev` on the all-reduced gradient sum.
"""
from __future__ import annotations

import numpy as np


class Adam:
    def __init__(self, eta: float = 1e-3, beta=(0.9, 0.999), eps: float = 1e-8):
        self.eta, self.beta, self.eps = float(eta), (float(beta[0]), float(beta[1])), float(eps)
        self.t = 0
        self.m = None
        self.v = None

    def update(self, p: np.ndarray, grad: np.ndarray) -> np.ndarray:
        """update!(opt, p, grad): in place on the host vector p (any float dtype)."""
        if self.m is None:
            self.m = np.zeros_like(p, dtype=np.float64); self.v = np.zeros_like(p, dtype=np.float64)
        self.t += 1
        b1, b2 = self.beta
        g = np.asarray(grad, dtype=np.float64)
        self.m = b1 * self.m + (1 - b1) * g
        self.v = b2 * self.v + (1 - b2) * g * g
        p -= (self.eta * (self.m / (1 - b1**self.t)) / (np.sqrt(self.v / (1 - b2**self.t)) + self.eps)).astype(p.dtype)
        return p

    def update_dev(self, ode, d_p, d_grad, d_m, d_v, grad_scale: float = 1.0) -> None:
        """Device pointers (ints) of np float32 each; see include/kanode.h:kanode_adam_step_dev."""
        from . import abi
        self.t += 1
        rc = ode.lib.kanode_adam_step_dev(ode.h, d_p, d_grad, d_m, d_v, self.t, self.eta, self.beta[0], self.beta[1],
                                          self.eps, grad_scale)
        abi.check(ode.lib, ode.h, rc, "kanode_adam_step_dev")
