"""Host-side mirror of the reference's ODE / gradient surface over the C ABI (include/kanode.h).

Mirrors the three call shapes the reference drivers use:
  NeuralODE(chain, tspan, Tsit5(); saveat)(u0, p, st) -> (sol, st)        LV/LV_driver_KANODE.jl:180-184
  ODEProblem(rc_kanode, u0, tspan, p; saveat) + solve(prob, Tsit5())      PDE/Allen-Cahn_Source.jl:96-99
  Zygote.gradient(loss, p)[1],  loss(p) = mean(abs2, X .- predict(p))     LV/LV_driver_KANODE.jl:197-203,284

All compute happens in libkanode_b200.so (CUDA, sm_100a).  There is no CPU fallback: if the library is missing or no
B200-class device is usable the constructors raise KanodeError.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import abi
from .layers import Chain


class Tsit5:
    """Marker for the integrator argument of NeuralODE / solve (the only one on the hot path)."""


@dataclass
class Stats:
    naccept: np.ndarray
    nreject: np.ndarray
    nf: np.ndarray
    retcode: np.ndarray

    @classmethod
    def from_raw(cls, raw) -> "Stats":
        a = np.frombuffer(raw, dtype=np.int32).reshape(-1, 4)
        return cls(a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy(), a[:, 3].copy())


@dataclass
class ODESolution:
    """What `node(u0, p, st)[1]` gives in the reference: `.t`, `.u`, `Array(sol)`, `.stats`, `.retcode`."""
    t: np.ndarray              # save times
    array: np.ndarray          # [batch, nsave, n]
    stats: Stats

    @property
    def u(self):
        return [self.array[:, i, :] for i in range(self.array.shape[1])]

    @property
    def retcode(self):
        return [abi.RETCODE_NAMES[int(r)] for r in self.stats.retcode]

    def __array__(self, dtype=None, copy=None):
        """Array(sol): [n, nsave] for a single trajectory (Julia layout), else [batch, nsave, n]."""
        a = self.array[0].T if self.array.shape[0] == 1 else self.array
        return a.astype(dtype) if dtype is not None else a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class KanOde:
    """Owns one kanode_handle: a KDense chain as an ODE right-hand side on one GPU."""

    def __init__(self, chain: Chain, rhs_kind: int = abi.RHS_CHAIN, n_state: int | None = None,
                 lap_coef: float = 0.0, dx: float = 1.0, device: int = 0, stream: int | None = None,
                 dtype=np.float32, devices: Sequence[int] | None = None):
        self.lib = abi.load_library()
        self.chain = chain
        self.desc = chain.desc(rhs_kind, n_state, lap_coef, dx)
        self.n = int(self.desc.n_state)
        self.n_out = chain.layers[-1].out_dims if rhs_kind == abi.RHS_MAP else self.n
        self.np_ = int(self.lib.kanode_param_count(C.byref(self.desc)))
        if self.np_ == 0:
            raise abi.KanodeError("invalid model descriptor")
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("dtype must be float32 or float64")
        self._suf = "" if self.dtype == np.float32 else "_f64"
        self._real = C.c_float if self.dtype == np.float32 else C.c_double
        h = C.c_void_p()
        if devices is not None and len(devices) > 1:
            # one handle, several GPUs of the box: every batch is sharded over them inside the library (kanode_create_multi)
            devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            rc = self.lib.kanode_create_multi(C.byref(self.desc), devs, len(devices), C.byref(h))
        else:
            rc = self.lib.kanode_create(C.byref(self.desc), int(device if not devices else devices[0]), C.c_void_p(stream), C.byref(h))
        abi.check(self.lib, None, rc, "kanode_create")
        self.h = h
        self.device = int(device if not devices else devices[0])

    def sync(self):
        """Block until the handle's stream(s) are idle (kanode_sync)."""
        abi.check(self.lib, self.h, self.lib.kanode_sync(self.h), "kanode_sync")

    def close(self):
        if getattr(self, "h", None):
            self.lib.kanode_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _fn(self, name):
        f = getattr(self.lib, name + self._suf)
        if self._suf:   # the _f64 variants take double tolerances; declare lazily
            f.restype = C.c_int
        return f

    def _arr(self, x, shape=None):
        a = np.ascontiguousarray(x, dtype=self.dtype)
        return a if shape is None else a.reshape(shape)

    # ---- parameters ----------------------------------------------------------------------------------
    def set_params(self, p) -> None:
        p = self._arr(p).reshape(-1)
        rc = self._fn("kanode_set_params")(self.h, _ptr(p), C.c_size_t(p.size))
        abi.check(self.lib, self.h, rc, "kanode_set_params")

    # ---- RHS / VJP -----------------------------------------------------------------------------------
    def rhs(self, u):
        u = self._arr(u).reshape(-1, self.n)
        du = np.empty((u.shape[0], self.n_out), self.dtype)            # n_out != n only for a map handle (abi.RHS_MAP)
        rc = self._fn("kanode_rhs")(self.h, _ptr(u), _ptr(du), C.c_int64(u.shape[0]))
        abi.check(self.lib, self.h, rc, "kanode_rhs")
        return du

    def vjp(self, u, lam):
        u = self._arr(u).reshape(-1, self.n)
        lam = self._arr(lam).reshape(-1, self.n_out)
        ubar = np.empty_like(u); pbar = np.empty(self.np_, self.dtype)
        rc = self._fn("kanode_vjp")(self.h, _ptr(u), _ptr(lam), _ptr(ubar), _ptr(pbar), C.c_int64(u.shape[0]))
        abi.check(self.lib, self.h, rc, "kanode_vjp")
        return ubar, pbar

    # ---- solve / loss+gradient -------------------------------------------------------------------------
    def solve(self, u0, tspan, saveat, abstol=1e-6, reltol=1e-3) -> ODESolution:
        u0 = self._arr(u0).reshape(-1, self.n)
        B = u0.shape[0]
        sa = np.ascontiguousarray(saveat, dtype=np.float64).reshape(-1)
        out = np.empty((B, sa.size, self.n), self.dtype)
        stats = (abi.Stats * max(B, 1))()
        f = self._fn("kanode_solve")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32,
                      self._real, self._real, C.c_void_p, C.c_void_p]
        rc = f(self.h, _ptr(u0), B, float(tspan[0]), float(tspan[1]), _ptr(sa), sa.size, abstol, reltol,
               _ptr(out), stats)
        abi.check(self.lib, self.h, rc, "kanode_solve")
        return ODESolution(sa, out, Stats.from_raw(stats))

    def loss_grad(self, u0, tspan, saveat, target, abstol=1e-6, reltol=1e-3, want_du0=True, want_stats=True,
                  allow_failed=False):
        """loss = mean(abs2, target - predict) and d loss / d p (what `Zygote.gradient(loss, p)[1]` returns in the reference,
        LV_driver_KANODE.jl:197-203,284).  `want_du0` / `want_stats` add d loss / d u0 and the per-trajectory solver
        statistics to the result (two more device-to-host copies of batch-sized arrays).  A trajectory whose forward or
        adjoint solve does not return Success raises KanodeError (the reference's `loss` throws on a short solution);
        `allow_failed=True` returns the result anyway (the failed trajectories are left out of loss / grad)."""
        u0 = self._arr(u0).reshape(-1, self.n)
        B = u0.shape[0]
        sa = np.ascontiguousarray(saveat, dtype=np.float64).reshape(-1)
        target = self._arr(target).reshape(B, sa.size, self.n)
        loss = self._real(0)
        grad = np.empty(self.np_, self.dtype)
        du0 = np.empty_like(u0) if want_du0 else None
        fst = (abi.Stats * B)() if want_stats else None
        bst = (abi.Stats * B)() if want_stats else None
        f = self._fn("kanode_loss_grad")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                      self._real, self._real, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = f(self.h, _ptr(u0), B, float(tspan[0]), float(tspan[1]), _ptr(sa), sa.size, _ptr(target), abstol, reltol,
               C.byref(loss), _ptr(grad), _ptr(du0), fst, bst)
        if not (rc == abi.ERR_SOLVER and allow_failed):
            abi.check(self.lib, self.h, rc, "kanode_loss_grad")
        out = dict(loss=float(loss.value), grad=grad, solver_failed=(rc == abi.ERR_SOLVER))
        if want_du0:
            out["du0"] = du0
        if want_stats:
            out["fwd_stats"] = Stats.from_raw(fst); out["bwd_stats"] = Stats.from_raw(bst)
        return out

    def loss_grad_replay(self, u0, tspan, saveat, target, fwd_t, bwd_t, abstol=1e-6, reltol=1e-3):
        """dt-replay: loss / gradient / predictions with the accepted-step END times of another run
        (`fwd_t`, `bwd_t`: [batch, max_steps], NaN padded — e.g. the fp64 oracle's) instead of the step-size controller."""
        u0 = self._arr(u0).reshape(-1, self.n)
        B = u0.shape[0]
        sa = np.ascontiguousarray(saveat, dtype=np.float64).reshape(-1)
        target = self._arr(target).reshape(B, sa.size, self.n)
        fwd_t = np.ascontiguousarray(fwd_t, dtype=np.float64).reshape(B, -1)
        bwd_t = np.ascontiguousarray(bwd_t, dtype=np.float64).reshape(B, -1)
        if fwd_t.shape != bwd_t.shape:
            raise ValueError("fwd_t and bwd_t must have the same [batch, max_steps] shape")
        loss = self._real(0)
        grad = np.empty(self.np_, self.dtype); du0 = np.empty_like(u0); out = np.empty_like(target)
        fst = (abi.Stats * B)(); bst = (abi.Stats * B)()
        f = self._fn("kanode_loss_grad_replay")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                      self._real, self._real, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.c_void_p, C.c_void_p, C.c_void_p]
        rc = f(self.h, _ptr(u0), B, float(tspan[0]), float(tspan[1]), _ptr(sa), sa.size, _ptr(target), abstol, reltol,
               _ptr(fwd_t), _ptr(bwd_t), fwd_t.shape[1], C.byref(loss), _ptr(grad), _ptr(du0), _ptr(out), fst, bst)
        abi.check(self.lib, self.h, rc, "kanode_loss_grad_replay")
        return dict(loss=float(loss.value), grad=grad, du0=du0, out=out, fwd_stats=Stats.from_raw(fst),
                    bwd_stats=Stats.from_raw(bst))

    def solve_adjoint(self, u0, tspan, saveat, dL_dout, abstol=1e-6, reltol=1e-3, allow_failed=False):
        """Pullback of the solve for an arbitrary loss: given dL/dpred [batch, nsave, n] returns the predictions of the dense
        forward solve, grad = sum_b (d pred_b/d p)^T dL_dout[b] and du0 (kanode_solve_adjoint; what Zygote.gradient(loss, p)
        runs for the reference's loss, LV_driver_KANODE.jl:197-203,284 / Burgers_Surrogate.jl:105-107,191)."""
        u0 = self._arr(u0).reshape(-1, self.n)
        B = u0.shape[0]
        sa = np.ascontiguousarray(saveat, dtype=np.float64).reshape(-1)
        cot = self._arr(dL_dout).reshape(B, sa.size, self.n)
        out = np.empty_like(cot); grad = np.empty(self.np_, self.dtype); du0 = np.empty_like(u0)
        fst = (abi.Stats * B)(); bst = (abi.Stats * B)()
        f = self._fn("kanode_solve_adjoint")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, self._real, self._real,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = f(self.h, _ptr(u0), B, float(tspan[0]), float(tspan[1]), _ptr(sa), sa.size, abstol, reltol, _ptr(cot), _ptr(out),
               _ptr(grad), _ptr(du0), fst, bst)
        if not (rc == abi.ERR_SOLVER and allow_failed):
            abi.check(self.lib, self.h, rc, "kanode_solve_adjoint")
        return dict(out=out, grad=grad, du0=du0, fwd_stats=Stats.from_raw(fst), bwd_stats=Stats.from_raw(bst))

    def edge_activations(self, layer: int, x):
        """act[k, i, o] = C[o,(i,g)]·basis_g(norm(x[k,i])) + W[o,i]·swish(x[k,i]) of layer `layer` (0-based) at its inputs
        x [K, I] (LV/Activation_getter.jl); act.sum(axis=1) is the layer output."""
        L = self.chain.layers[layer]
        x = self._arr(x).reshape(-1, L.in_dims)
        act = np.empty((x.shape[0], L.in_dims, L.out_dims), self.dtype)
        rc = self._fn("kanode_edge_activations")(self.h, C.c_int32(layer), _ptr(x), _ptr(act), C.c_int64(x.shape[0]))
        abi.check(self.lib, self.h, rc, "kanode_edge_activations")
        return act

    def set_regularizer(self, act_reg: float = 0.0, entropy_reg: float = 0.0) -> None:
        """reg_loss(p, act_reg, entropy_reg) is added to every loss_grad from now on (LV_driver_KANODE.jl:187-201:
        `sparse_on == 1` uses (5e-4, 0)); (0, 0) switches it off."""
        self.lib.kanode_set_regularizer.argtypes = [C.c_void_p, C.c_double, C.c_double]
        rc = self.lib.kanode_set_regularizer(self.h, float(act_reg), float(entropy_reg))
        abi.check(self.lib, self.h, rc, "kanode_set_regularizer")

    def reg_loss(self, act_reg: float = 1.0, entropy_reg: float = 1.0):
        """reg_loss of the current parameters and its gradient (float32)."""
        loss = C.c_double(0); grad = np.empty(self.np_, np.float32)
        self.lib.kanode_reg_loss.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        rc = self.lib.kanode_reg_loss(self.h, float(act_reg), float(entropy_reg), C.byref(loss), _ptr(grad))
        abi.check(self.lib, self.h, rc, "kanode_reg_loss")
        return float(loss.value), grad

    def launch_count(self) -> int:
        return int(self.lib.kanode_launch_count(self.h))


class NeuralODE:
    """NeuralODE(model, tspan, Tsit5(); saveat, abstol, reltol) — [EXT DiffEqFlux 4.0.0] call surface.

    `node(u0, p, st)` returns `(sol, st)` like the reference (LV_driver_KANODE.jl:183); `node.loss_and_grad(u0, p, X)`
    is `loss(p)` and `Zygote.gradient(loss, p)[1]` of LV_driver_KANODE.jl:197-203,284 in one call.
    """

    def __init__(self, model: Chain, tspan: Sequence[float], alg: Tsit5 | None = None, *, saveat=(),
                 abstol: float = 1e-6, reltol: float = 1e-3, device: int = 0, dtype=np.float32):
        if alg is not None and not isinstance(alg, Tsit5):
            raise NotImplementedError("only Tsit5() is on the hot path")
        self.model, self.tspan = model, (float(tspan[0]), float(tspan[1]))
        self.saveat = np.asarray(saveat, dtype=np.float64)
        self.abstol, self.reltol = abstol, reltol
        self.ode = KanOde(model, device=device, dtype=dtype)

    def __call__(self, u0, p, st=None):
        self.ode.set_params(p)
        return self.ode.solve(u0, self.tspan, self.saveat, self.abstol, self.reltol), st

    def loss_and_grad(self, u0, p, target):
        self.ode.set_params(p)
        r = self.ode.loss_grad(u0, self.tspan, self.saveat, target, self.abstol, self.reltol)
        return r["loss"], r["grad"], r

    def pullback(self, u0, p):
        """`pred, back = Zygote.pullback(p -> Array(node(u0, p, st)[1]), p)`: back(dL_dpred) -> (dL/dp, dL/du0)."""
        self.ode.set_params(p)
        sol = self.ode.solve(u0, self.tspan, self.saveat, self.abstol, self.reltol)

        def back(dL_dpred):
            self.ode.set_params(p)
            r = self.ode.solve_adjoint(u0, self.tspan, self.saveat, dL_dpred, self.abstol, self.reltol)
            return r["grad"], r["du0"]
        return sol.array, back

    def gradient(self, loss_and_cotangent, u0, p):
        """`Zygote.gradient(loss, p)[1]` for the reference's unchanged `loss(p)`: `loss_and_cotangent(pred)` returns
        (loss value, dloss/dpred) for pred [batch, nsave, n] — any loss of the predictions (mean(abs2, X - pred), a transposed
        target, extra terms), differentiated by the caller; the solve's pullback comes from kanode_solve_adjoint."""
        pred, back = self.pullback(u0, p)
        value, cot = loss_and_cotangent(pred)
        return value, back(cot)[0]


class SourceODE(NeuralODE):
    """`rc_kanode(u,p,t) = s*D*lap*u + kan1_.(u)` + `solve(prob, Tsit5())`  (PDE/Allen-Cahn_Source.jl:90-99,
    PDE/Fisher-KPP_Source.jl:95-104): periodic 3-point Laplacian plus a pointwise 1->1 KAN."""

    def __init__(self, model: Chain, n_state: int, lap_coef: float, dx: float, tspan, alg=None, *, saveat=(),
                 abstol: float = 1e-6, reltol: float = 1e-3, device: int = 0, dtype=np.float32):
        if alg is not None and not isinstance(alg, Tsit5):
            raise NotImplementedError("only Tsit5() is on the hot path")
        self.model, self.tspan = model, (float(tspan[0]), float(tspan[1]))
        self.saveat = np.asarray(saveat, dtype=np.float64)
        self.abstol, self.reltol = abstol, reltol
        self.ode = KanOde(model, abi.RHS_SOURCE_LAPLACIAN, n_state, lap_coef, dx, device=device, dtype=dtype)
