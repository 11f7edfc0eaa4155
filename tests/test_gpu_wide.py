"""The batched lockstep ("wide") engine for [n,10,n] surrogates (kan_odes_b200/csrc/kanode_wide.cuh) against
  (a) the block-per-trajectory kernels of the same library (KANODE_WIDE=0) — two independent CUDA formulations of the
      reference's per-IC semantics: identical step counts, states/gradients to reduction-order round-off (fp64);
  (b) the CPU oracle (fp64 identical step counts; fp32 at solver accuracy),
on ragged shapes: n not a multiple of the 128-unit tiles, batches that do not fill the IC tiles, ICs whose step counts
differ (finished ICs are masked while the others keep stepping), save times equal to the end time (lambda jump at
t1), a dense record that overflows and is regrown, and the empty time span.
Reference semantics: PDE examples/Burgers_Surrogate.jl:82-107, Allen-Cahn_Surrogate.jl:80-107, Schrodinger_Surrogate.jl:89-114."""
import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import glorot_params, surrogate_chain
from oracle import Oracle

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-300)


def _problem(n, G, B, seed, t_end=1.0, save_end=False):
    chain = surrogate_chain(n, 10, G)
    p = glorot_params(chain, seed=seed)
    x = np.linspace(-1, 1, n)
    rng = np.random.default_rng(seed + 100)
    amp = rng.uniform(0.2, 2.5, (B, 1))                      # spread amplitudes -> different step counts per IC
    u0 = -amp * np.sin(np.pi * x)[None, :] + 0.1 * rng.normal(size=(B, n))
    saveat = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9]) * t_end
    if save_end:
        saveat = np.append(saveat, t_end)
    tg = u0[:, None, :] * np.exp(-saveat)[None, :, None]
    return chain, p, u0, (0.0, t_end), saveat, tg


def _run(chain, p, u0, tspan, saveat, tg, dtype, wide, monkeypatch, cap=None):
    monkeypatch.setenv("KANODE_WIDE", "1" if wide else "0")
    ode = K.KanOde(chain, dtype=dtype)
    ode.set_params(p)
    if cap:
        ode.lib.kanode_set_record_capacity(ode.h, cap)
    launches0 = ode.launch_count()
    sol = ode.solve(u0, tspan, saveat)
    r = ode.loss_grad(u0, tspan, saveat, tg)
    r["launches"] = ode.launch_count() - launches0
    ode.close()
    return sol, r


@pytest.mark.parametrize("n,G,B,save_end", [(130, 5, 11, False), (257, 10, 9, True), (41, 5, 3, False)])
def test_wide_matches_block_per_trajectory_fp64(n, G, B, save_end, monkeypatch):
    chain, p, u0, tspan, saveat, tg = _problem(n, G, B, seed=n, save_end=save_end)
    sw, rw = _run(chain, p, u0, tspan, saveat, tg, np.float64, True, monkeypatch)
    sg, rg = _run(chain, p, u0, tspan, saveat, tg, np.float64, False, monkeypatch)
    assert rw["launches"] > 10 * rg["launches"]                     # the lockstep engine really ran (many small launches)
    for a, b in ((sw.stats, sg.stats), (rw["fwd_stats"], rg["fwd_stats"]), (rw["bwd_stats"], rg["bwd_stats"])):
        assert (a.naccept == b.naccept).all() and (a.nreject == b.nreject).all()
        assert (a.nf == b.nf).all() and (a.retcode == 0).all() and (b.retcode == 0).all()
    assert len(set(rw["bwd_stats"].naccept.tolist())) > 1 or B < 4      # ICs with different step counts were masked
    assert _relmax(sw.array, sg.array) < 1e-12
    assert abs(rw["loss"] - rg["loss"]) < 1e-12 * abs(rg["loss"])
    assert _relmax(rw["grad"], rg["grad"]) < 1e-10
    assert _relmax(rw["du0"], rg["du0"]) < 1e-10


def test_wide_matches_oracle_ragged_batch(monkeypatch):
    chain, p, u0, tspan, saveat, tg = _problem(300, 10, 21, seed=5)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, tspan, saveat, tg)
    _, r = _run(chain, p, u0, tspan, saveat, tg, np.float64, True, monkeypatch)
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all() and (r["fwd_stats"].nf == ref["fwd_stats"][:, 2]).all()
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all() and (r["bwd_stats"].nreject == ref["bwd_stats"][:, 1]).all()
    assert (r["bwd_stats"].nf == ref["bwd_stats"][:, 2]).all()
    assert abs(r["loss"] - ref["loss"]) < 1e-10 * abs(ref["loss"])
    assert _relmax(r["grad"], ref["grad"]) < 1e-7 and _relmax(r["du0"], ref["du0"]) < 1e-7   # reduction order differs from the oracle
    # fp32 instantiation: solver accuracy (step sequences decorrelate in fp32, see tests/test_gpu_lv.py)
    _, r32 = _run(chain, p, u0, tspan, saveat, tg, np.float32, True, monkeypatch)
    assert (r32["fwd_stats"].retcode == 0).all() and (r32["bwd_stats"].retcode == 0).all()
    assert abs(r32["loss"] - ref["loss"]) < 5e-3 * abs(ref["loss"])
    assert _relmax(r32["grad"], ref["grad"]) < 2e-2


def test_wide_dense_record_overflow_is_regrown(monkeypatch):
    chain, p, u0, tspan, saveat, tg = _problem(130, 5, 6, seed=9)
    _, full = _run(chain, p, u0, tspan, saveat, tg, np.float64, True, monkeypatch)
    _, small = _run(chain, p, u0, tspan, saveat, tg, np.float64, True, monkeypatch, cap=2)   # 2 steps cannot hold the solve
    assert (small["fwd_stats"].retcode == 0).all() and (small["fwd_stats"].naccept == full["fwd_stats"].naccept).all()
    assert _relmax(small["grad"], full["grad"]) < 1e-13


def test_wide_repeated_calls_are_deterministic_and_track_new_parameters(monkeypatch):
    """Deterministic two-level reductions: the same call gives bit-identical results; a parameter update is picked up
    (the transposed layer-1 weights are refreshed) and the attempt-count prediction of the host loop stays valid."""
    chain, p, u0, tspan, saveat, tg = _problem(200, 5, 8, seed=3)
    monkeypatch.setenv("KANODE_WIDE", "1")
    ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
    a = ode.loss_grad(u0, tspan, saveat, tg); b = ode.loss_grad(u0, tspan, saveat, tg)
    assert np.array_equal(a["grad"], b["grad"]) and a["loss"] == b["loss"]
    p2 = (p * 0.5).astype(np.float32)
    ode.set_params(p2); c = ode.loss_grad(u0, tspan, saveat, tg)
    ode.set_params(p); d = ode.loss_grad(u0, tspan, saveat, tg)
    ode.close()
    ref2 = Oracle(chain.desc(), np.float64).loss_grad(p2, u0, tspan, saveat, tg)
    assert _relmax(c["grad"], ref2["grad"]) < 2e-2 and abs(c["loss"] - ref2["loss"]) < 5e-3 * abs(ref2["loss"])
    assert np.array_equal(a["grad"], d["grad"])


def test_wide_empty_time_span(monkeypatch):
    chain, p, u0, _, _, _ = _problem(64, 5, 2, seed=1)
    monkeypatch.setenv("KANODE_WIDE", "1")
    ode = K.KanOde(chain, dtype=np.float64); ode.set_params(p)
    sol = ode.solve(u0, (0.5, 0.5), [0.5])
    ode.close()
    assert (sol.stats.retcode == 0).all() and (sol.stats.naccept == 0).all()
    assert np.array_equal(sol.array[:, 0, :], u0)


@pytest.mark.parametrize("n,G,B", [(256, 10, 5), (384, 5, 40)])
def test_wide_tensor_core_kernels_match_cuda_core_fp32(n, G, B, monkeypatch):
    """fp32: the tcgen05 contraction kernels (3xTF32, TMEM accumulator; layer-2 forward for any n, layer-2 reverse for
    n % 128 == 0) against the CUDA-core kernels of the same engine (KANODE_WIDE_TC=0), and both against the fp64 oracle."""
    chain, p, u0, tspan, saveat, tg = _problem(n, G, B, seed=n + 1)
    orc = Oracle(chain.desc(), np.float64)
    res = {}
    for tc in (2, 1, 0):                                                   # 2: also the fused basis-expansion + contraction kernel
        monkeypatch.setenv("KANODE_WIDE", "1"); monkeypatch.setenv("KANODE_WIDE_TC", str(tc))
        ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
        rhs = ode.rhs(u0)
        r = ode.loss_grad(u0, tspan, saveat, tg)
        ode.close()
        res[tc] = (rhs, r)
    ref_rhs = orc.rhs(p, u0)
    for tc in (2, 1, 0):
        assert _relmax(res[tc][0], ref_rhs) < 2e-5, tc                      # RHS arithmetic within the fp32 parity budget
    assert _relmax(res[1][0], res[0][0]) < 5e-6
    for a, b in ((res[1][1], res[0][1]), (res[2][1], res[0][1])):
        _compare_tc(a, b)
    ref = orc.loss_grad(p, u0, tspan, saveat, tg)
    a = res[2][1]
    assert _relmax(a["grad"], ref["grad"]) < 2e-2 and abs(a["loss"] - ref["loss"]) < 5e-3 * abs(ref["loss"])


def _compare_tc(a, b):
    assert (a["fwd_stats"].retcode == 0).all() and (a["bwd_stats"].retcode == 0).all()
    same = (a["fwd_stats"].naccept == b["fwd_stats"].naccept).all() and (a["bwd_stats"].naccept == b["bwd_stats"].naccept).all() \
        and (a["bwd_stats"].nreject == b["bwd_stats"].nreject).all()
    tol = 2e-4 if same else 2e-2                                            # different fp32 step sequences: solver accuracy only
    assert abs(a["loss"] - b["loss"]) < tol * abs(b["loss"])
    assert _relmax(a["grad"], b["grad"]) < tol, (same, _relmax(a["grad"], b["grad"]))


@pytest.mark.parametrize("n,G,B", [(300, 5, 7), (41, 10, 3)])
def test_hidden_source_lockstep_matches_block_per_trajectory(n, G, B, monkeypatch):
    """Hidden-source model (periodic Laplacian + pointwise KDense(1,1,G); Allen-Cahn_Source.jl:50-54,90-99) through the
    lockstep engine of csrc/kanode_wsrc.cuh against the block-per-trajectory kernels and the oracle: ragged n (block-edge
    neighbours, periodic wrap inside a partial block), ICs with different step counts, save time at the end."""
    from conftest import source_chain
    from kan_odes_b200 import abi
    chain = source_chain(G)
    p = glorot_params(chain, seed=G)
    kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=2e-4, dx=2.0 / (n - 1))
    x = np.linspace(-1, 1, n)
    amp = np.random.default_rng(n).uniform(0.3, 2.0, (B, 1))
    u0 = amp * (x**2 * np.cos(np.pi * x))[None, :]
    tspan, saveat = (0.0, 0.5), np.linspace(0, 0.5, 6)
    tg = u0[:, None, :] * np.exp(0.5 * saveat)[None, :, None]
    res = {}
    for wide in (1, 0):
        monkeypatch.setenv("KANODE_WIDE", str(wide))
        ode = K.KanOde(chain, kw["rhs_kind"], n, kw["lap_coef"], kw["dx"], dtype=np.float64); ode.set_params(p)
        l0 = ode.launch_count()
        sol = ode.solve(u0, tspan, saveat); r = ode.loss_grad(u0, tspan, saveat, tg)
        res[wide] = (sol, r, ode.launch_count() - l0)
        ode.close()
    (sw, rw, lw), (sg, rg, lg) = res[1], res[0]
    assert lw > 10 * lg                                                 # the lockstep engine really ran
    for a, b in ((sw.stats, sg.stats), (rw["fwd_stats"], rg["fwd_stats"]), (rw["bwd_stats"], rg["bwd_stats"])):
        assert (a.naccept == b.naccept).all() and (a.nreject == b.nreject).all() and (a.nf == b.nf).all() and (a.retcode == 0).all()
    assert _relmax(sw.array, sg.array) < 1e-12 and _relmax(rw["grad"], rg["grad"]) < 1e-9 and _relmax(rw["du0"], rg["du0"]) < 1e-9
    ref = Oracle(chain.desc(**kw), np.float64).loss_grad(p, u0, tspan, saveat, tg)
    assert (rw["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all() and _relmax(rw["grad"], ref["grad"]) < 1e-7
    monkeypatch.setenv("KANODE_WIDE", "1")
    ode = K.KanOde(chain, kw["rhs_kind"], n, kw["lap_coef"], kw["dx"], dtype=np.float32); ode.set_params(p)
    r32 = ode.loss_grad(u0, tspan, saveat, tg); ode.close()
    assert (r32["bwd_stats"].retcode == 0).all() and _relmax(r32["grad"], ref["grad"]) < 2e-2


def test_gpass_timing_entry_point(monkeypatch):
    """kanode_last_gpass_timing: available after a wide-engine step launched directly (KANODE_WIDE_GRAPH=0), an error after an
    LV (thread-per-trajectory) step; kanode_last_timing works for both."""
    import ctypes as C
    from conftest import lv_chain
    monkeypatch.setenv("KANODE_WIDE", "1"); monkeypatch.setenv("KANODE_WIDE_GRAPH", "0")
    chain, p, u0, tspan, saveat, tg = _problem(256, 5, 4, seed=2)
    ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
    r = ode.loss_grad(u0, tspan, saveat, tg)
    ms, passes, m3 = C.c_float(), C.c_int32(), (C.c_float * 3)()
    assert ode.lib.kanode_last_gpass_timing(ode.h, C.byref(ms), C.byref(passes)) == 0
    assert passes.value >= int(r["bwd_stats"].naccept.max()) and ms.value > 0
    assert ode.lib.kanode_last_timing(ode.h, m3) == 0 and m3[1] > ms.value
    ode.close()
    lv = lv_chain()
    ode = K.KanOde(lv, dtype=np.float32); ode.set_params(glorot_params(lv, seed=0))
    ode.loss_grad(np.ones((4, 2)), (0.0, 1.0), [0.5, 1.0], np.ones((4, 2, 2)))
    assert ode.lib.kanode_last_gpass_timing(ode.h, C.byref(ms), C.byref(passes)) != 0
    ode.close()
