"""Oracle integrator + adjoint against independent references (scipy DOP853, finite differences)."""
import numpy as np
import pytest
import torch
from scipy.integrate import solve_ivp

from conftest import glorot_params, lv_chain, lv_targets, source_chain, surrogate_chain
from kan_odes_b200 import abi
from kdense_ref import rhs_torch
from oracle import Oracle


def _scipy_solve(chain, desc, p, u0, tspan, saveat):
    pt = torch.tensor(p)

    def f(t, u):
        return rhs_torch(chain, desc, pt, torch.tensor(u)[None, :])[0].numpy()
    s = solve_ivp(f, tspan, u0, method="DOP853", t_eval=saveat, rtol=1e-12, atol=1e-13)
    return s.y.T


def test_tight_tolerance_solution_matches_scipy(lv_saveat):
    chain = lv_chain(); desc = chain.desc()
    p = glorot_params(chain, seed=0).astype(np.float64)
    orc = Oracle(desc)
    out, st = orc.solve(p, [[1.0, 1.0]], (0.0, 3.5), lv_saveat, abstol=1e-11, reltol=1e-11)
    ref = _scipy_solve(chain, desc, p, [1.0, 1.0], (0.0, 3.5), lv_saveat)
    assert st[0, 3] == abi.RET_SUCCESS
    assert np.max(np.abs(out[0] - ref)) < 5e-9 * max(1.0, np.abs(ref).max())


def test_default_tolerance_solution_error_is_of_order_reltol(lv_saveat):
    chain = lv_chain(); desc = chain.desc()
    p = glorot_params(chain, seed=0).astype(np.float64)
    orc = Oracle(desc)
    out, st = orc.solve(p, [[1.0, 1.0]], (0.0, 3.5), lv_saveat)          # abstol=1e-6, reltol=1e-3 (defaults)
    ref = _scipy_solve(chain, desc, p, [1.0, 1.0], (0.0, 3.5), lv_saveat)
    err = np.max(np.abs(out[0] - ref)) / (1e-3 * np.abs(ref).max())      # global error in units of reltol*|u|
    assert 1e-3 < err < 5.0
    na, nr, nf, rc = st[0]
    assert rc == 0 and nf == 3 + 6 * (na + nr)                           # 2 initdt + 1 FSAL + 6/attempt


def test_saveat_values_are_interpolated_not_stepped_to(lv_saveat):
    """saveat must not change the step sequence: stats identical for any saveat."""
    chain = lv_chain(); desc = chain.desc()
    p = glorot_params(chain, seed=0).astype(np.float64)
    orc = Oracle(desc)
    _, st1 = orc.solve(p, [[1.0, 1.0]], (0.0, 3.5), lv_saveat)
    _, st2 = orc.solve(p, [[1.0, 1.0]], (0.0, 3.5), [3.5])
    assert (st1 == st2).all()


def test_source_rhs_solve_matches_scipy():
    chain = source_chain(10)
    n = 41
    desc = chain.desc(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=-1e-4, dx=0.05)
    p = glorot_params(chain, seed=3).astype(np.float64)
    x = np.linspace(-1, 1, n)
    u0 = x**2 * np.cos(np.pi * x)                                        # AC_Source.jl:44
    sa = np.linspace(0, 1, 11)
    orc = Oracle(desc)
    out, st = orc.solve(p, [u0], (0.0, 1.0), sa, abstol=1e-11, reltol=1e-11)
    ref = _scipy_solve(chain, desc, p, u0, (0.0, 1.0), sa)
    assert st[0, 3] == 0 and np.max(np.abs(out[0] - ref)) < 1e-8


def _torch_rk4_loss_grad(chain, desc, p, u0, saveat, tg, t_end, nsub=40):
    """Independent gradient: discretise-then-optimise through a fine fixed-step RK4 with torch autograd (fp64)."""
    pt = torch.tensor(p, requires_grad=True)
    u = torch.tensor(np.asarray(u0, dtype=np.float64))
    f = lambda v: rhs_torch(chain, desc, pt, v)
    times = list(saveat) + ([t_end] if t_end > saveat[-1] else [])
    outs, t = [], 0.0
    for ts in times:
        if ts > t:
            h = (ts - t) / nsub
            for _ in range(nsub):
                k1 = f(u); k2 = f(u + h / 2 * k1); k3 = f(u + h / 2 * k2); k4 = f(u + h * k3)
                u = u + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
            t = ts
        outs.append(u)
    out = torch.stack(outs[:len(saveat)], 1)
    loss = ((out - torch.tensor(tg))**2).mean()
    loss.backward()
    return loss.item(), pt.grad.numpy()


def _fd_grad(orc, p, u0, tspan, sa, tg, idx, h=1e-4, tol=1e-13):
    g = np.zeros(len(idx))
    for j, i in enumerate(idx):
        pp = p.copy(); pp[i] += h
        pm = p.copy(); pm[i] -= h
        lp = np.mean((orc.solve(pp, u0, tspan, sa, abstol=tol, reltol=tol)[0] - tg)**2)
        lm = np.mean((orc.solve(pm, u0, tspan, sa, abstol=tol, reltol=tol)[0] - tg)**2)
        g[j] = (lp - lm) / (2 * h)
    return g


def test_adjoint_gradient_lv_vs_autograd_and_finite_differences(lv_saveat):
    chain = lv_chain(); desc = chain.desc()
    p = glorot_params(chain, seed=0).astype(np.float64)
    orc = Oracle(desc)
    u0 = np.array([[1.0, 1.0], [0.7, 1.6]])
    tg = lv_targets(u0, lv_saveat)
    r = orc.loss_grad(p, u0, (0.0, 3.5), lv_saveat, tg, abstol=1e-12, reltol=1e-12, want_out=True)
    assert abs(r["loss"] - np.mean((r["out"] - tg)**2)) < 1e-14
    loss_t, grad_t = _torch_rk4_loss_grad(chain, desc, p, u0, lv_saveat, tg, 3.5, nsub=24)
    gmax = np.abs(grad_t).max()
    assert abs(loss_t - r["loss"]) < 1e-9 * max(1.0, loss_t)
    assert np.max(np.abs(r["grad"] - grad_t)) < 2e-8 * gmax
    idx = np.random.default_rng(0).choice(240, 12, replace=False)
    fd = _fd_grad(orc, p, u0, (0.0, 3.5), lv_saveat, tg, idx)
    assert np.max(np.abs(r["grad"][idx] - fd)) < 1e-6 * gmax
    # default tolerances (abstol=1e-6, reltol=1e-3): same gradient to solver accuracy
    r2 = orc.loss_grad(p, u0, (0.0, 3.5), lv_saveat, tg)
    assert np.linalg.norm(r2["grad"] - r["grad"]) < 2e-2 * np.linalg.norm(r["grad"])


def test_adjoint_gradient_surrogate_and_source_vs_autograd():
    # Burgers-like surrogate at n=12 (Burgers_Surrogate.jl:82-107; save times :68, none at T)
    chain = surrogate_chain(12, 4, 5); desc = chain.desc()
    p = glorot_params(chain, seed=1).astype(np.float64)
    x = np.linspace(-1, 1, 12)
    u0 = -np.sin(np.pi * x)[None, :]
    sa = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9])
    tg = u0[:, None, :] * np.exp(-sa)[None, :, None]
    orc = Oracle(desc)
    r = orc.loss_grad(p, u0, (0.0, 1.0), sa, tg, abstol=1e-12, reltol=1e-12)
    loss_t, grad_t = _torch_rk4_loss_grad(chain, desc, p, u0, sa, tg, 1.0, nsub=100)
    assert abs(loss_t - r["loss"]) < 1e-10
    assert np.max(np.abs(r["grad"] - grad_t)) < 2e-8 * np.abs(grad_t).max()
    # hidden-source model (Allen-Cahn_Source.jl:90-104; saveat=dt includes both ends)
    chain = source_chain(6)
    desc = chain.desc(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=9, lap_coef=-1e-4, dx=0.25)
    p = glorot_params(chain, seed=2).astype(np.float64)
    x = np.linspace(-1, 1, 9)
    u0 = (x**2 * np.cos(np.pi * x))[None, :]
    sa = np.linspace(0, 1, 6)
    tg = u0[:, None, :] * (1 + sa)[None, :, None]
    orc = Oracle(desc)
    r = orc.loss_grad(p, u0, (0.0, 1.0), sa, tg, abstol=1e-12, reltol=1e-12)
    loss_t, grad_t = _torch_rk4_loss_grad(chain, desc, p, u0, sa, tg, 1.0, nsub=100)
    assert abs(loss_t - r["loss"]) < 1e-10
    assert np.max(np.abs(r["grad"] - grad_t)) < 2e-8 * np.abs(grad_t).max()


def test_backward_solve_structure(lv_saveat):
    """tstops at every save time force >= nsave backward steps (a14, a16); z(T)=0 gives the 1e-6 start-up ramp."""
    chain = lv_chain(); desc = chain.desc()
    p = glorot_params(chain, seed=0).astype(np.float64)
    orc = Oracle(desc)
    tg = lv_targets([[1.0, 1.0]], lv_saveat)
    r = orc.loss_grad(p, [[1.0, 1.0]], (0.0, 3.5), lv_saveat, tg)
    na, nr, nf, rc = r["bwd_stats"][0]
    assert rc == 0 and na >= 34 + 5
    # nf = 3 (init) + 6 per attempt + one FSAL re-evaluation per save-time jump (35 here; t=0 is a save time)
    assert nf == 3 + 6 * (na + nr) + 35 - 1 or nf == 3 + 6 * (na + nr) + 35


def test_float32_oracle_vs_float64(lv_saveat):
    """fp32 state arithmetic cannot reproduce fp64 step sequences when the field is non-trivial: the embedded error
    estimate of a conservative step (EEst ~ 1e-5) is below fp32 round-off of the stage values, so the next dt is
    noise-dominated (DESIGN.md 'precision').  What must hold: fp32 results stay within solver tolerance of fp64, and
    for the reference's actual training start (p = glorot/1e5, LV_driver_KANODE.jl:175) they agree tightly."""
    chain = lv_chain(); desc = chain.desc()
    p = glorot_params(chain, seed=0)
    u0 = np.random.default_rng(1234).uniform(0.5, 2.0, (64, 2))
    tg = lv_targets(u0, lv_saveat)
    r64 = Oracle(desc, np.float64).loss_grad(p, u0, (0.0, 3.5), lv_saveat, tg, want_out=True)
    r32 = Oracle(desc, np.float32).loss_grad(p, u0, (0.0, 3.5), lv_saveat, tg, want_out=True)
    assert np.abs(r32["out"] - r64["out"]).max() < 5e-3 * np.abs(r64["out"]).max()
    assert np.abs(r32["grad"] - r64["grad"]).max() < 5e-3 * np.abs(r64["grad"]).max()
    p0 = p * np.float32(1e-5)
    r64 = Oracle(desc, np.float64).loss_grad(p0, u0, (0.0, 3.5), lv_saveat, tg, want_out=True)
    r32 = Oracle(desc, np.float32).loss_grad(p0, u0, (0.0, 3.5), lv_saveat, tg, want_out=True)
    assert (r64["fwd_stats"] == r32["fwd_stats"]).all() and (r64["bwd_stats"] == r32["bwd_stats"]).all()
    assert np.abs(r32["out"] - r64["out"]).max() < 1e-5 * np.abs(r64["out"]).max()
    assert np.abs(r32["grad"] - r64["grad"]).max() < 1e-5 * np.abs(r64["grad"]).max()
