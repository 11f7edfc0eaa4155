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


class DeviceTrainer:
    """The reference's training iteration (LV_driver_KANODE.jl:280-291) resident on the device: `grad = Zgrad(loss, p)[1]`,
    `update!(opt, p, grad)`, `loss_train(p)`, `loss_test(p)` per step() without a host round trip (kanode_train_step_dev).
    Data, parameters, Adam moments and the loss history live in HBM; `losses()` / `params()` read them back."""

    def __init__(self, ode, u0, target, tspan, saveat, *, u0_test=None, target_test=None, tspan_test=None, saveat_test=None,
                 eta: float = 5e-4, beta=(0.9, 0.999), eps: float = 1e-8, abstol: float = 1e-6, reltol: float = 1e-3,
                 max_iters: int = 100000, device: int = 0):
        import ctypes as C

        import torch
        from . import abi
        if ode.dtype != np.float32:
            raise ValueError("device-resident training runs the fp32 kernels")
        self.ode, self.C, self.abi = ode, C, abi
        dev = torch.device("cuda", device)
        f32 = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
        self.u0 = f32(np.asarray(u0).reshape(-1, ode.n)); self.B = self.u0.shape[0]
        self.sa = np.ascontiguousarray(saveat, dtype=np.float64)
        self.tg = f32(np.asarray(target).reshape(self.B, self.sa.size, ode.n))
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.test = target_test is not None
        if self.test:
            self.u0t = f32(np.asarray(u0 if u0_test is None else u0_test).reshape(-1, ode.n)); self.Bt = self.u0t.shape[0]
            self.sat = np.ascontiguousarray(saveat_test, dtype=np.float64)
            self.tgt = f32(np.asarray(target_test).reshape(self.Bt, self.sat.size, ode.n))
            self.t1t = float(tspan_test[1])
        self.hist = torch.zeros((max_iters, 3), dtype=torch.float64, device=dev)
        self.it = 0
        self.abstol, self.reltol = abstol, reltol
        f = ode.lib.kanode_train_begin
        f.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        abi.check(ode.lib, ode.h, f(ode.h, eta, beta[0], beta[1], eps), "kanode_train_begin")
        self._step = ode.lib.kanode_train_step_dev
        self._step.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                               C.c_float, C.c_float, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]

    def step(self) -> None:
        """One iteration, enqueued on the handle's stream (returns before it has run)."""
        t = self.test
        rc = self._step(self.ode.h, self.u0.data_ptr(), self.B, self.tspan[0], self.tspan[1], self.sa.ctypes.data, self.sa.size,
                        self.tg.data_ptr(), self.abstol, self.reltol, self.u0t.data_ptr() if t else None, self.Bt if t else 0,
                        self.t1t if t else 0.0, self.sat.ctypes.data if t else None, self.sat.size if t else 0,
                        self.tgt.data_ptr() if t else None, self.hist[self.it].data_ptr())
        self.abi.check(self.ode.lib, self.ode.h, rc, "kanode_train_step_dev")
        self.it += 1

    def losses(self) -> np.ndarray:
        """[iterations, 3]: loss(p_k) (with the regulariser when set), loss_train(p_{k+1}), loss_test(p_{k+1})."""
        self.ode.lib.kanode_sync(self.ode.h)
        return self.hist[:self.it].cpu().numpy()

    def params(self) -> np.ndarray:
        p = np.empty(self.ode.np_, np.float32)
        rc = self.ode.lib.kanode_train_params(self.ode.h, p.ctypes.data_as(self.C.c_void_p))
        self.abi.check(self.ode.lib, self.ode.h, rc, "kanode_train_params")
        return p
