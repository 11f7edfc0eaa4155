"""Parity of the generic (block-per-trajectory) CUDA path against the CPU oracle at the REFERENCE's own sizes:
  Burgers surrogate      [41,10,41]  G=5  softsign   PDE examples/Burgers_Surrogate.jl:82-107
  Allen-Cahn surrogate   [41,10,41]  G=10 softsign   PDE examples/Allen-Cahn_Surrogate.jl:80-107 (n=42 with the BC node)
  Schrodinger surrogate  [402,10,402] G=10           PDE examples/Schrodinger_Surrogate.jl:68,89-114
  Allen-Cahn / Fisher-KPP hidden source  [1,1] G=10 + periodic Laplacian  Allen-Cahn_Source.jl:50-54,76-104, Fisher-KPP_Source.jl
plus chains the small-model registry does not cover.  fp64 instantiation: identical step counts, 1e-8 relative.
fp32: solver accuracy (see tests/test_gpu_lv.py for why step sequences decorrelate in fp32)."""
import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import glorot_params, lv_chain, source_chain, surrogate_chain
from kan_odes_b200 import abi
from oracle import Oracle

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-300)


def _check_full(chain, kw, p, u0, tspan, saveat, tg, batch_note=""):
    desc = chain.desc(**kw)
    orc = Oracle(desc, np.float64)
    lam = np.random.default_rng(7).normal(size=u0.shape)
    ode = K.KanOde(chain, kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0),
                   dtype=np.float64)
    ode.set_params(p)
    assert _relmax(ode.rhs(u0), orc.rhs(p, u0)) < 1e-12
    ub, pb = ode.vjp(u0, lam); ub_r, pb_r = orc.vjp(p, u0, lam)
    assert _relmax(ub, ub_r) < 1e-11 and _relmax(pb, pb_r) < 1e-11
    sol = ode.solve(u0, tspan, saveat)
    out_r, st_r = orc.solve(p, u0, tspan, saveat)
    assert (sol.stats.naccept == st_r[:, 0]).all() and (sol.stats.nreject == st_r[:, 1]).all()
    assert (sol.stats.nf == st_r[:, 2]).all() and (sol.stats.retcode == 0).all()
    assert _relmax(sol.array, out_r) < 1e-8
    r = ode.loss_grad(u0, tspan, saveat, tg)
    ref = orc.loss_grad(p, u0, tspan, saveat, tg)
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all()
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all(), (r["bwd_stats"].naccept, ref["bwd_stats"][:, 0])
    assert (r["bwd_stats"].nreject == ref["bwd_stats"][:, 1]).all()
    assert (r["bwd_stats"].nf == ref["bwd_stats"][:, 2]).all()
    assert abs(r["loss"] - ref["loss"]) < 1e-9 * abs(ref["loss"])
    assert _relmax(r["grad"], ref["grad"]) < 1e-7       # wide block reductions sum in a different order
    assert _relmax(r["du0"], ref["du0"]) < 1e-7
    ode.close()
    # fp32 instantiation: RHS/VJP arithmetic at 1e-5, solve + gradient at solver accuracy
    ode = K.KanOde(chain, kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0),
                   dtype=np.float32)
    ode.set_params(p)
    assert _relmax(ode.rhs(u0), orc.rhs(p, u0)) < 2e-5
    ub, pb = ode.vjp(u0, lam)
    assert _relmax(ub, ub_r) < 5e-5 and _relmax(pb, pb_r) < 5e-5
    r32 = ode.loss_grad(u0, tspan, saveat, tg)
    assert (r32["fwd_stats"].retcode == 0).all() and (r32["bwd_stats"].retcode == 0).all()
    assert abs(r32["loss"] - ref["loss"]) < 5e-3 * abs(ref["loss"])
    assert _relmax(r32["grad"], ref["grad"]) < 2e-2
    ode.close()


def test_burgers_surrogate_reference_size():
    n = 41
    chain = surrogate_chain(n, 10, 5)
    p = glorot_params(chain, seed=0)
    x = np.linspace(-1, 1, n)
    amp = np.random.default_rng(3).uniform(0.5, 1.5, 3)
    u0 = -amp[:, None] * np.sin(np.pi * x)[None, :]                 # u(0,x) = -sin(pi x), Burgers_Surrogate.jl:48
    saveat = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9])               # dt_train :68
    tg = u0[:, None, :] * np.exp(-saveat)[None, :, None]
    _check_full(chain, {}, p, u0, (0.0, 1.0), saveat, tg)


def test_allen_cahn_surrogate_reference_size():
    n = 42                                                          # u0 = [-1.0; prob.u0] AC_Surrogate.jl:93
    chain = surrogate_chain(n, 10, 10)
    p = glorot_params(chain, seed=1)
    x = np.linspace(-1, 1, n)
    u0 = (x**2 * np.cos(np.pi * x))[None, :] * np.array([[1.0], [0.7]])
    saveat = np.array([0.1, 0.3, 0.5, 0.7, 0.9])
    tg = u0[:, None, :] * (1 - 0.5 * saveat)[None, :, None]
    _check_full(chain, {}, p, u0, (0.0, 1.0), saveat, tg)


def test_schrodinger_surrogate_reference_size():
    n = 402                                                         # [Re; Im] on 201 nodes, Schrodinger_Surrogate.jl:68
    chain = surrogate_chain(n, 10, 10)
    p = glorot_params(chain, seed=2)
    x = np.linspace(-5, 5, 201)
    u0 = np.concatenate([2 / np.cosh(x), np.zeros(201)])[None, :] * np.array([[1.0], [0.9]])
    saveat = np.array([0.1, 0.3, 0.5, 0.7, 0.9, 1.1, 1.3, 1.5])     # dt_train :73
    tg = u0[:, None, :] * np.cos(saveat)[None, :, None]
    _check_full(chain, {}, p, u0, (0.0, np.pi / 2), saveat, tg)


def test_allen_cahn_hidden_source_reference_size():
    n = 41
    chain = source_chain(10)
    kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=-1e-4, dx=0.05)   # AC_Source.jl:36,58,92
    p = glorot_params(chain, seed=3)
    x = np.linspace(-1, 1, n)
    u0 = (x**2 * np.cos(np.pi * x))[None, :] * np.array([[1.0], [1.3]])                # :44
    saveat = np.linspace(0, 1, 101)                                                    # saveat = dt = 0.01 :37,97
    tg = u0[:, None, :] * np.exp(0.5 * saveat)[None, :, None]
    _check_full(chain, kw, p, u0, (0.0, 1.0), saveat, tg)


def test_fisher_kpp_hidden_source_reference_size():
    n = 26                                                                              # x = 0:0.04:1 Fisher-KPP_Source.jl:41-44
    chain = source_chain(10)
    kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=0.01, dx=0.04)    # +D*lap :36,64,97
    p = glorot_params(chain, seed=4)
    x = np.arange(n) * 0.04
    rho0 = (np.tanh((x - 0.4) / 0.02) - np.tanh((x - 0.6) / 0.02)) / 2                 # :48-50
    saveat = np.linspace(0, 5, 11)                                                     # dt = T/10
    tg = rho0[None, None, :] * (1 + 0.1 * saveat)[None, :, None]
    _check_full(chain, kw, p, rho0[None, :], (0.0, 5.0), saveat, tg)


def test_chains_outside_the_small_registry():
    # an LV-shaped model of another width/grid (LV/trend_plotter.py sweep) goes through the generic path
    chain = lv_chain(5, 3)
    p = glorot_params(chain, seed=5)
    u0 = np.random.default_rng(1).uniform(0.5, 2.0, (5, 2))
    saveat = np.arange(35) * 0.1
    tg = np.random.default_rng(2).uniform(0, 3, (5, 35, 2))
    _check_full(chain, {}, p, u0, (0.0, 3.5), saveat, tg)
    # three layers, rswaf basis, sigmoid normalizer, one layer without the base branch (kdense.jl:122-127, utils.jl:27-43)
    chain = K.Chain(K.KDense(6, 4, 6, basis_func=K.rswaf, normalizer=K.sigmoid, use_base_act=False),
                    K.KDense(4, 7, 4, basis_func=K.rswaf, normalizer=K.softsign),
                    K.KDense(7, 6, 5, basis_func=K.iqf, normalizer=K.tanh_fast))
    p = glorot_params(chain, seed=6)
    u0 = np.random.default_rng(3).uniform(-1, 1, (3, 6))
    saveat = np.array([0.0, 0.25, 0.5, 1.0])
    tg = np.zeros((3, 4, 6))
    _check_full(chain, {}, p, u0, (0.0, 1.0), saveat, tg)


def test_burgers_1024_batch_runs_and_matches_oracle_fp64():
    """BASELINE configs[2] shape (n=1024) at a small batch: fp64 parity with the oracle."""
    n = 1024
    chain = surrogate_chain(n, 10, 5)
    p = glorot_params(chain, seed=0)
    x = np.linspace(-1, 1, n)
    amp = np.random.default_rng(3).uniform(0.5, 1.5, 2)
    u0 = -amp[:, None] * np.sin(np.pi * x)[None, :]
    saveat = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9])
    tg = u0[:, None, :] * np.exp(-saveat)[None, :, None]
    ode = K.KanOde(chain, dtype=np.float64); ode.set_params(p)
    r = ode.loss_grad(u0, (0.0, 1.0), saveat, tg)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, (0.0, 1.0), saveat, tg)
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert _relmax(r["grad"], ref["grad"]) < 1e-8


def test_allen_cahn_surrogate_4096_wide_layer_fp64():
    """BASELINE configs[3] shape: [4096,10,4096] G=10 (the 'wide layer'), one IC, fp64 parity with the oracle."""
    n = 4096
    chain = surrogate_chain(n, 10, 10)
    p = glorot_params(chain, seed=1)
    x = np.linspace(-1, 1, n)
    u0 = (x**2 * np.cos(np.pi * x))[None, :]
    saveat = np.array([0.1, 0.3, 0.5, 0.7, 0.9])
    tg = u0[:, None, :] * (1 - 0.5 * saveat)[None, :, None]
    ode = K.KanOde(chain, dtype=np.float64); ode.set_params(p)
    r = ode.loss_grad(u0, (0.0, 1.0), saveat, tg)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, (0.0, 1.0), saveat, tg)
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all()
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all() and (r["bwd_stats"].nf == ref["bwd_stats"][:, 2]).all()
    assert abs(r["loss"] - ref["loss"]) < 1e-9 * abs(ref["loss"])
    assert _relmax(r["grad"], ref["grad"]) < 1e-7


_SCHROD = {}


def _schrodinger_case():
    """BASELINE configs[4] shape: [32768,10,32768] G=10 (Re/Im split on a 16,384-point grid), one IC; the fp64 oracle result is
    computed once (about a minute of CPU) and shared by the fp64 and the fp32 / tcgen05 tests."""
    if not _SCHROD:
        n = 32768
        chain = surrogate_chain(n, 10, 10)
        p = glorot_params(chain, seed=2)
        x = np.linspace(-5, 5, 16384)
        u0 = np.concatenate([2 / np.cosh(x), np.zeros_like(x)])[None, :]
        saveat = np.array([0.1, 0.3, 0.5, 0.7, 0.9, 1.1, 1.3, 1.5])
        tg = u0[:, None, :] * np.cos(saveat)[None, :, None]
        orc = Oracle(chain.desc(), np.float64)
        _SCHROD.update(chain=chain, p=p, u0=u0, saveat=saveat, tg=tg, orc=orc,
                       ref=orc.loss_grad(p, u0, (0.0, np.pi / 2), saveat, tg))
    return _SCHROD


def test_schrodinger_16384_fp64():
    """BASELINE configs[4] shape, one IC, fp64 parity."""
    c = _schrodinger_case(); ref = c["ref"]
    ode = K.KanOde(c["chain"], dtype=np.float64); ode.set_params(c["p"])
    r = ode.loss_grad(c["u0"], (0.0, np.pi / 2), c["saveat"], c["tg"])
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all()
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert abs(r["loss"] - ref["loss"]) < 1e-9 * abs(ref["loss"])
    assert _relmax(r["grad"], ref["grad"]) < 1e-7


def _tc_vs_oracle(chain, p, u0, tspan, saveat, tg, orc, ref, monkeypatch, label):
    """fp32 with the tcgen05 contraction kernels (wide_l2_fwd_tc, wide_l2_vjp_tc: n % 128 == 0) against the fp64 oracle, and
    against the CUDA-core kernels of the same engine (KANODE_WIDE_TC=0) to show that the tensor-core kernels are the ones that
    ran.  RHS arithmetic within 2e-5; loss / gradient of the adaptive fp32 solve within solver accuracy (fp32 cannot follow the
    fp64 step sequence)."""
    res = {}
    for tc in (1, 0):
        monkeypatch.setenv("KANODE_WIDE", "1"); monkeypatch.setenv("KANODE_WIDE_TC", str(tc))
        ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
        res[tc] = (ode.rhs(u0), ode.loss_grad(u0, tspan, saveat, tg))
        ode.close()
    ref_rhs = orc.rhs(p, u0)
    e_rhs = _relmax(res[1][0], ref_rhs)
    assert e_rhs < 2e-5 and _relmax(res[0][0], ref_rhs) < 2e-5
    assert not np.array_equal(res[1][0], res[0][0])                       # 3xTF32 on TMEM vs fp32 FFMA: close, not bit-identical
    r = res[1][1]
    assert (r["fwd_stats"].retcode == 0).all() and (r["bwd_stats"].retcode == 0).all()
    e_g, e_l = _relmax(r["grad"], ref["grad"]), abs(r["loss"] - ref["loss"]) / abs(ref["loss"])
    same = ((r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]) & (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0])).mean()
    print(f"{label}: fp32/tcgen05 vs fp64 oracle: rhs {e_rhs:.2e}, loss {e_l:.2e}, gradient {e_g:.2e}, identical step counts {100 * same:.0f}% of ICs; "
          f"tcgen05 vs CUDA cores: gradient {_relmax(r['grad'], res[0][1]['grad'].astype(np.float64)):.2e}")
    assert e_g < 5e-5 and e_l < 5e-6 and e_rhs < 5e-6                    # measured on B200: rhs 1-2e-6, loss 4e-8..3e-7, gradient 4e-6..1.1e-5


@pytest.mark.parametrize("name,n,G,B", [("burgers1024", 1024, 5, 64), ("ac4096", 4096, 10, 32)])
def test_tcgen05_kernels_vs_oracle_at_config_sizes_batched_fp32(name, n, G, B, monkeypatch):
    """BASELINE configs[2] / configs[3] at their model size and a real batch (the TC kernels had only been compared
    with the oracle at n = 256 / 384)."""
    import bench
    chain, kw, p, u0, ts, sa, tg = bench.pde_make(name, B, np.random.default_rng(3))
    orc = Oracle(chain.desc(), np.float64); orc.set_threads(__import__("os").cpu_count() or 1)
    ref = orc.loss_grad(p, u0, ts, sa, tg)
    _tc_vs_oracle(chain, p, u0, ts, sa, tg, orc, ref, monkeypatch, f"{name} x {B}")


def test_tcgen05_kernels_vs_oracle_schrodinger_16384_fp32(monkeypatch):
    c = _schrodinger_case()
    _tc_vs_oracle(c["chain"], c["p"], c["u0"], (0.0, np.pi / 2), c["saveat"], c["tg"], c["orc"], c["ref"], monkeypatch, "schrodinger16384 x 1")


def test_hidden_source_4096_fp64():
    """BASELINE configs[3] hidden-source shape at N=4096 (periodic Laplacian + 1->1 KAN).  The Allen-Cahn script's sign
    (-1e-4*lap, Allen-Cahn_Source.jl:92) is anti-diffusive and blows up at dx=2/4095 (growth rate 4*419/s), so the
    stable Fisher-KPP sign (+D*lap, Fisher-KPP_Source.jl:97) is used with D/dx^2 = 419."""
    n = 4096
    chain = source_chain(10)
    kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=1e-4, dx=2.0 / (n - 1))
    p = glorot_params(chain, seed=3)
    x = np.linspace(-1, 1, n)
    u0 = (x**2 * np.cos(np.pi * x))[None, :]
    saveat = np.linspace(0, 0.2, 21)
    tg = u0[:, None, :] * np.exp(0.5 * saveat)[None, :, None]
    ode = K.KanOde(chain, kw["rhs_kind"], n, kw["lap_coef"], kw["dx"], dtype=np.float64); ode.set_params(p)
    r = ode.loss_grad(u0, (0.0, 0.2), saveat, tg)
    ref = Oracle(chain.desc(**kw), np.float64).loss_grad(p, u0, (0.0, 0.2), saveat, tg)
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all() and (r["fwd_stats"].retcode == 0).all()
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert _relmax(r["grad"], ref["grad"]) < 1e-7
    # the reference's anti-diffusive sign at this resolution: both sides report the blow-up instead of a number
    bad = dict(kw, lap_coef=-1e-4)
    ode2 = K.KanOde(chain, bad["rhs_kind"], n, bad["lap_coef"], bad["dx"], dtype=np.float64); ode2.set_params(p)
    sol = ode2.solve(u0, (0.0, 1.0), [1.0])
    _, st = Oracle(chain.desc(**bad), np.float64).solve(p, u0, (0.0, 1.0), [1.0])
    assert sol.stats.retcode[0] != 0 and st[0, 3] != 0
