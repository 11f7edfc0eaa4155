"""ctypes wrapper of oracle/libkanode_oracle.so (TEST INFRASTRUCTURE ONLY).

Follows oracle/kanode_oracle.cpp, which restates LV/src/kdense.jl:109-130, LV/src/utils.jl:8-21 and the
un-vendored SciML solver/adjoint algorithms.  The descriptor structs are the public ones of include/kanode.h
(kan_odes_b200.abi) — types only; nothing of the product's compute path is used here.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from kan_odes_b200 import abi

ORACLE_DIR = Path(__file__).resolve().parent
ORACLE_LIB = ORACLE_DIR / "libkanode_oracle.so"


def build_oracle(force: bool = False) -> Path:
    src = ORACLE_DIR / "kanode_oracle.cpp"
    if force or not ORACLE_LIB.exists() or ORACLE_LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B"], check=True, capture_output=True)
    return ORACLE_LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """dtype=np.float64 restates what the reference drivers run; np.float32 mirrors device arithmetic."""

    def __init__(self, desc: abi.Desc, dtype=np.float64):
        build_oracle()
        self.lib = C.CDLL(str(ORACLE_LIB))
        self.desc = desc
        self.dtype = np.dtype(dtype)
        self.suf = "f64" if self.dtype == np.float64 else "f32"
        self.lib.kanode_oracle_param_count.restype = C.c_size_t
        self.np_ = int(self.lib.kanode_oracle_param_count(C.byref(desc)))
        if self.np_ == 0:
            raise ValueError("invalid descriptor")
        self.n = int(desc.n_state)

    def set_threads(self, n: int) -> int:
        """OpenMP threads of the batched entry points; returns the count in effect (n <= 0: query only)."""
        self.lib.kanode_oracle_set_threads.restype = C.c_int
        return int(self.lib.kanode_oracle_set_threads(C.c_int(n)))

    def _fn(self, name):
        f = getattr(self.lib, f"kanode_oracle_{name}_{self.suf}")
        f.restype = C.c_int
        return f

    def _a(self, x, shape=None):
        a = np.ascontiguousarray(x, dtype=self.dtype)
        return a if shape is None else a.reshape(shape)

    def rhs(self, p, u):
        u = self._a(u).reshape(-1, self.n); p = self._a(p)
        du = np.empty_like(u)
        rc = self._fn("rhs")(C.byref(self.desc), _ptr(p), _ptr(u), _ptr(du), C.c_int64(u.shape[0]))
        assert rc == 0, rc
        return du

    def vjp(self, p, u, lam):
        u = self._a(u).reshape(-1, self.n); lam = self._a(lam).reshape(-1, self.n); p = self._a(p)
        ubar = np.empty_like(u); pbar = np.empty(self.np_, self.dtype)
        rc = self._fn("vjp")(C.byref(self.desc), _ptr(p), _ptr(u), _ptr(lam), _ptr(ubar), _ptr(pbar),
                             C.c_int64(u.shape[0]))
        assert rc == 0, rc
        return ubar, pbar

    def solve(self, p, u0, tspan, saveat, abstol=1e-6, reltol=1e-3, step_cap=0):
        u0 = self._a(u0).reshape(-1, self.n); p = self._a(p)
        B = u0.shape[0]
        sa = np.ascontiguousarray(saveat, dtype=np.float64)
        out = np.empty((B, sa.size, self.n), self.dtype)
        stats = (abi.Stats * B)()
        step_t = np.full((B, step_cap), np.nan) if step_cap else None
        rc = self._fn("solve")(C.byref(self.desc), _ptr(p), _ptr(u0), C.c_int64(B), C.c_double(tspan[0]),
                               C.c_double(tspan[1]), _ptr(sa), C.c_int(sa.size), C.c_double(abstol),
                               C.c_double(reltol), _ptr(out), stats, _ptr(step_t), C.c_int(step_cap))
        assert rc == 0, rc
        st = np.frombuffer(stats, dtype=np.int32).reshape(B, 4).copy()
        return (out, st, step_t) if step_cap else (out, st)

    def loss_grad(self, p, u0, tspan, saveat, target, abstol=1e-6, reltol=1e-3, want_out=False, step_cap=0):
        """step_cap > 0 additionally returns `fwd_t` / `bwd_t` [B, step_cap]: END times of the accepted forward / adjoint
        steps (NaN padded) — the sequences kanode_loss_grad_replay takes."""
        u0 = self._a(u0).reshape(-1, self.n); p = self._a(p)
        B = u0.shape[0]
        sa = np.ascontiguousarray(saveat, dtype=np.float64)
        target = self._a(target).reshape(B, sa.size, self.n)
        loss = C.c_double(0.0)
        grad = np.empty(self.np_, self.dtype); du0 = np.empty_like(u0)
        fst = (abi.Stats * B)(); bst = (abi.Stats * B)()
        out = np.empty((B, sa.size, self.n), self.dtype) if want_out else None
        args = [C.byref(self.desc), _ptr(p), _ptr(u0), C.c_int64(B), C.c_double(tspan[0]), C.c_double(tspan[1]), _ptr(sa),
                C.c_int(sa.size), _ptr(target), C.c_double(abstol), C.c_double(reltol), C.byref(loss), _ptr(grad),
                _ptr(du0), fst, bst, _ptr(out)]
        fwd_t = bwd_t = None
        if step_cap:
            fwd_t = np.full((B, step_cap), np.nan); bwd_t = np.full((B, step_cap), np.nan)
            rc = self._fn("loss_grad_steps")(*args, _ptr(fwd_t), _ptr(bwd_t), C.c_int(step_cap))
        else:
            rc = self._fn("loss_grad")(*args)
        assert rc == 0, rc
        f = np.frombuffer(fst, dtype=np.int32).reshape(B, 4).copy()
        b = np.frombuffer(bst, dtype=np.int32).reshape(B, 4).copy()
        res = dict(loss=loss.value, grad=grad, du0=du0, fwd_stats=f, bwd_stats=b)
        if want_out:
            res["out"] = out
        if step_cap:
            res["fwd_t"] = fwd_t; res["bwd_t"] = bwd_t
        return res

    def adjoint(self, p, u0, tspan, saveat, cot, abstol=1e-6, reltol=1e-3):
        """Pullback of the solve with caller-supplied cotangents dL/dpred [B, nsave, n] (kanode_solve_adjoint's counterpart):
        grad = sum_b (d pred_b/d p)^T cot_b (no 1/B), du0, predictions, statistics."""
        u0 = self._a(u0).reshape(-1, self.n); p = self._a(p)
        B = u0.shape[0]
        sa = np.ascontiguousarray(saveat, dtype=np.float64)
        cot = self._a(cot).reshape(B, sa.size, self.n)
        grad = np.empty(self.np_, self.dtype); du0 = np.empty_like(u0); out = np.empty_like(cot)
        fst = (abi.Stats * B)(); bst = (abi.Stats * B)()
        rc = self._fn("adjoint")(C.byref(self.desc), _ptr(p), _ptr(u0), C.c_int64(B), C.c_double(tspan[0]), C.c_double(tspan[1]),
                                 _ptr(sa), C.c_int(sa.size), _ptr(cot), C.c_double(abstol), C.c_double(reltol), _ptr(grad),
                                 _ptr(du0), fst, bst, _ptr(out))
        assert rc == 0, rc
        return dict(grad=grad, du0=du0, out=out, fwd_stats=np.frombuffer(fst, dtype=np.int32).reshape(B, 4).copy(),
                    bwd_stats=np.frombuffer(bst, dtype=np.int32).reshape(B, 4).copy())

    def map(self, p, x):
        """The chain as a map x [K, I_first] -> y [K, O_last] (direct layer call, kdense.jl:109-130); desc.rhs_kind = RHS_MAP."""
        I = int(self.desc.layers[0].in_dims); O = int(self.desc.layers[self.desc.n_layers - 1].out_dims)
        x = self._a(x).reshape(-1, I); p = self._a(p)
        y = np.empty((x.shape[0], O), self.dtype)
        rc = self._fn("map")(C.byref(self.desc), _ptr(p), _ptr(x), _ptr(y), C.c_int64(x.shape[0]))
        assert rc == 0, rc
        return y

    def edge_activations(self, p, layer, x):
        """act[k, i, o] of layer `layer` at its inputs x [K, I_l] (LV/Activation_getter.jl)."""
        L = self.desc.layers[layer]
        x = self._a(x).reshape(-1, int(L.in_dims)); p = self._a(p)
        act = np.empty((x.shape[0], int(L.in_dims), int(L.out_dims)), self.dtype)
        rc = self._fn("edge_activations")(C.byref(self.desc), C.c_int(layer), _ptr(p), _ptr(x), _ptr(act), C.c_int64(x.shape[0]))
        assert rc == 0, rc
        return act

    def reg_loss(self, p, act_reg=1.0, entropy_reg=1.0):
        """reg_loss (LV_driver_KANODE.jl:187-194) and its gradient."""
        p = self._a(p); loss = C.c_double(0); grad = np.empty_like(p)
        rc = self._fn("reg_loss")(_ptr(p), C.c_size_t(p.size), C.c_double(act_reg), C.c_double(entropy_reg), C.byref(loss), _ptr(grad))
        assert rc == 0, rc
        return float(loss.value), grad

    # small helpers for the unit tests
    def fastpower(self, x, y):
        self.lib.kanode_oracle_fastpower.restype = C.c_float
        return float(self.lib.kanode_oracle_fastpower(C.c_double(x), C.c_double(y)))

    def interp_weights(self, th):
        b = np.empty(7)
        self.lib.kanode_oracle_interp_weights(C.c_double(th), _ptr(b))
        return b

    def tableau(self):
        c = np.empty(6); a = np.empty((7, 7)); bt = np.empty(7)
        self.lib.kanode_oracle_tableau(_ptr(c), _ptr(a), _ptr(bt))
        return c, a, bt
