"""Parity of the CUDA path (through the C ABI) against the CPU oracle — Lotka-Volterra KAN-ODE [2,10,2] G=5
(BASELINE.json configs[0] and a slice of configs[1]).

Tolerances (north_star: 1e-5 relative, same accepted-step count):
  * fp64 instantiation vs fp64 oracle: identical step counts, 1e-8 relative — same algorithm, same precision.
  * fp32 instantiation vs fp64 oracle at the reference's training start p = glorot/1e5
    (LV_driver_KANODE.jl:175): identical step counts, 1e-5 relative.
  * fp32 on a non-trivial field: the embedded error estimate of conservative steps is below fp32 round-off of the
    stage values, so step sequences decorrelate (shown for the fp32 ORACLE too in tests/test_oracle_solve.py); the
    bound is solver accuracy (5e-3), and the RHS/VJP arithmetic itself is checked at 1e-5.
"""
import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import glorot_params, lv_chain, lv_targets
from oracle import Oracle

pytestmark = pytest.mark.gpu

TSPAN = (0.0, 3.5)


@pytest.fixture(scope="module")
def setup_lv(lv_saveat):
    chain = lv_chain()
    p = glorot_params(chain, seed=0)
    u0 = np.random.default_rng(1234).uniform(0.5, 2.0, (300, 2))   # not a multiple of the block size
    u0[0] = (1.0, 1.0)                                              # the reference's u0 (LV_driver:117)
    tg = lv_targets(u0, lv_saveat)
    return chain, p, u0, tg


def _relmax(a, b):
    return np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max()


def test_rhs_and_vjp(setup_lv):
    chain, p, u0, _ = setup_lv
    orc = Oracle(chain.desc(), np.float64)
    lam = np.random.default_rng(3).normal(size=u0.shape)
    du_ref = orc.rhs(p, u0); ub_ref, pb_ref = orc.vjp(p, u0, lam)
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        ode = K.KanOde(chain, dtype=dtype); ode.set_params(p)
        assert _relmax(ode.rhs(u0), du_ref) < tol
        ub, pb = ode.vjp(u0, lam)
        assert _relmax(ub, ub_ref) < tol and _relmax(pb, pb_ref) < tol
        ode.close()


def test_solve_fp64_matches_oracle_stepwise(setup_lv, lv_saveat):
    chain, p, u0, _ = setup_lv
    out_ref, st_ref = Oracle(chain.desc(), np.float64).solve(p, u0, TSPAN, lv_saveat)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat, dtype=np.float64)
    sol, _ = node(u0, p, None)
    assert (sol.stats.naccept == st_ref[:, 0]).all() and (sol.stats.nreject == st_ref[:, 1]).all()
    assert (sol.stats.nf == st_ref[:, 2]).all() and (sol.stats.retcode == 0).all()
    assert _relmax(sol.array, out_ref) < 1e-8
    assert np.asarray(K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat, dtype=np.float64)(u0[:1], p)[0]).shape == (2, 35)


def test_loss_grad_fp64_matches_oracle(setup_lv, lv_saveat):
    chain, p, u0, tg = setup_lv
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat, dtype=np.float64)
    loss, grad, info = node.loss_and_grad(u0, p, tg)
    assert (info["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all()
    assert (info["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert (info["bwd_stats"].nreject == ref["bwd_stats"][:, 1]).all()
    assert (info["bwd_stats"].nf == ref["bwd_stats"][:, 2]).all()
    assert abs(loss - ref["loss"]) < 1e-10 * ref["loss"]
    assert _relmax(grad, ref["grad"]) < 1e-8
    assert _relmax(info["du0"], ref["du0"]) < 1e-8


def test_fp32_at_reference_training_start(setup_lv, lv_saveat):
    """p = glorot/1e5: the state the reference driver differentiates at iteration 1 (LV_driver_KANODE.jl:175,284)."""
    chain, p, u0, tg = setup_lv
    p0 = (p * np.float32(1e-5)).astype(np.float32)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p0, u0, TSPAN, lv_saveat, tg, want_out=True)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat, dtype=np.float32)
    sol, _ = node(u0, p0)
    assert _relmax(sol.array, ref["out"]) < 1e-5
    loss, grad, info = node.loss_and_grad(u0, p0, tg)
    assert (info["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all()
    assert (info["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert abs(loss - ref["loss"]) < 1e-5 * ref["loss"]
    assert _relmax(grad, ref["grad"]) < 1e-5


def test_fp32_nontrivial_field_within_solver_accuracy(setup_lv, lv_saveat):
    chain, p, u0, tg = setup_lv
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg, want_out=True)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat, dtype=np.float32)
    sol, _ = node(u0, p)
    assert (sol.stats.retcode == 0).all()
    assert _relmax(sol.array, ref["out"]) < 5e-3
    loss, grad, info = node.loss_and_grad(u0, p, tg)
    assert abs(loss - ref["loss"]) < 5e-3 * ref["loss"]
    assert _relmax(grad, ref["grad"]) < 5e-3
    assert np.abs(info["fwd_stats"].naccept - ref["fwd_stats"][:, 0]).max() <= 2


def test_single_trajectory_reference_config(lv_saveat):
    """BASELINE configs[0]: u0=(1,1), tspan (0,3.5), saveat 0:0.1:3.4, true-LV targets (LV_driver:111-127)."""
    chain = lv_chain()
    p = glorot_params(chain, seed=0)
    u0 = np.array([[1.0, 1.0]])
    tg = lv_targets(u0, lv_saveat)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat, dtype=np.float64)
    loss, grad, info = node.loss_and_grad(u0, p, tg)
    assert int(info["fwd_stats"].naccept[0]) == int(ref["fwd_stats"][0, 0])
    assert int(info["bwd_stats"].naccept[0]) == int(ref["bwd_stats"][0, 0])
    assert _relmax(grad, ref["grad"]) < 1e-8


def test_save_time_at_both_ends_of_tspan(setup_lv):
    """saveat containing t0 and T: the adjoint jump at T fires at initialisation (PresetTimeCallback), the one at
    t0 after the last step (Allen-Cahn_Source.jl:97 uses saveat=dt, which includes both ends)."""
    chain, p, u0, _ = setup_lv
    sa = np.linspace(0.0, 3.5, 8)
    tg = np.random.default_rng(5).uniform(0, 3, (6, 8, 2))
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0[:6], TSPAN, sa, tg)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=sa, dtype=np.float64)
    loss, grad, info = node.loss_and_grad(u0[:6], p, tg)
    assert (info["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert (info["bwd_stats"].nf == ref["bwd_stats"][:, 2]).all()
    assert _relmax(grad, ref["grad"]) < 1e-8 and _relmax(info["du0"], ref["du0"]) < 1e-8


def test_edge_cases(setup_lv, lv_saveat):
    chain, p, u0, tg = setup_lv
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=[3.5], dtype=np.float64)
    sol, _ = node(u0[:3], p)                                       # only the end point saved
    ref, _ = Oracle(chain.desc()).solve(p, u0[:3], TSPAN, [3.5])
    assert _relmax(sol.array, ref) < 1e-9
    empty = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat)(np.zeros((0, 2)), p)[0]
    assert empty.array.shape == (0, 35, 2)
    with pytest.raises(K.KanodeError):                             # save time outside tspan
        K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=[4.0])(u0[:1], p)
    with pytest.raises(K.KanodeError):                             # wrong parameter count
        node.ode.set_params(p[:-1])
    # tiny record capacity: the host API grows it and still returns the same answer
    ode = K.KanOde(chain, dtype=np.float64); ode.set_params(p)
    ode.lib.kanode_set_record_capacity(ode.h, 2)
    r = ode.loss_grad(u0[:5], TSPAN, lv_saveat, tg[:5])
    ref = Oracle(chain.desc()).loss_grad(p, u0[:5], TSPAN, lv_saveat, tg[:5])
    assert _relmax(r["grad"], ref["grad"]) < 1e-8


def test_full_size_ensemble_properties(lv_saveat):
    """BASELINE configs[1] at full size (65,536 trajectories, fp32): size-independent properties.
    (a) trajectories are independent: a permuted batch gives the permuted outputs bit for bit;
    (b) sharding: per-shard unnormalised sums add up to the full-batch result (what the multi-GPU path relies on);
    (c) a random sample of trajectories agrees with the fp64 oracle within solver accuracy;
    (d) duplicated initial conditions give identical trajectories and step counts."""
    chain = lv_chain()
    p = glorot_params(chain, seed=0)
    B = 65536
    rng = np.random.default_rng(1234)
    u0 = rng.uniform(0.5, 2.0, (B, 2)).astype(np.float32)
    u0[B // 2:B // 2 + 64] = u0[:64]                                   # (d) duplicates
    tg = rng.uniform(0.0, 3.0, (B, 35, 2)).astype(np.float32)
    ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
    sol = ode.solve(u0, TSPAN, lv_saveat)
    assert (sol.stats.retcode == 0).all() and np.isfinite(sol.array).all()
    assert np.array_equal(sol.array[:64], sol.array[B // 2:B // 2 + 64])
    assert np.array_equal(sol.stats.naccept[:64], sol.stats.naccept[B // 2:B // 2 + 64])
    perm = rng.permutation(B)
    sol_p = ode.solve(u0[perm], TSPAN, lv_saveat)
    assert np.array_equal(sol_p.array, sol.array[perm])                 # (a)
    full = ode.loss_grad(u0, TSPAN, lv_saveat, tg)
    halves = [ode.loss_grad(u0[s], TSPAN, lv_saveat, tg[s]) for s in (slice(0, 40000), slice(40000, B))]
    g = (halves[0]["grad"].astype(np.float64) * 40000 + halves[1]["grad"].astype(np.float64) * (B - 40000)) / B
    l = (halves[0]["loss"] * 40000 + halves[1]["loss"] * (B - 40000)) / B
    assert _relmax(g, full["grad"].astype(np.float64)) < 2e-6 and abs(l - full["loss"]) < 2e-6 * full["loss"]   # (b)
    idx = rng.choice(B, 256, replace=False)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0[idx], TSPAN, lv_saveat, tg[idx], want_out=True)
    assert _relmax(sol.array[idx], ref["out"]) < 5e-3                    # (c)
    sub = ode.loss_grad(u0[idx], TSPAN, lv_saveat, tg[idx])
    assert _relmax(sub["grad"], ref["grad"]) < 5e-3 and _relmax(sub["du0"], ref["du0"]) < 2e-2


def test_training_loop_reduces_loss_and_adam_kernel_matches_host():
    """The reference's training iteration (LV_driver_KANODE.jl:279-305) on the CUDA path: the loss goes down from the
    p/1e5 start; the fused device Adam kernel reproduces the host update."""
    import ctypes as C
    import torch
    sys_path_root = str(__import__("pathlib").Path(__file__).resolve().parent.parent)
    import sys
    sys.path.insert(0, sys_path_root)
    from examples.train_lv import lv_data
    t, X = lv_data()
    chain = lv_chain()
    p = (glorot_params(chain, 0).astype(np.float64) / 1e5)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=t[:35], dtype=np.float64)
    Xtr = X[:, :35].T[None]
    opt = K.Adam(5e-3)
    losses = []
    for _ in range(60):
        loss, grad, _ = node.loss_and_grad(np.array([[1.0, 1.0]]), p, Xtr)
        opt.update(p, grad); losses.append(loss)
    assert losses[-1] < 0.7 * losses[0] and np.isfinite(losses).all()
    # device Adam vs host Adam
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=240).astype(np.float32); g = rng.normal(size=240).astype(np.float32)
    host = K.Adam(1e-2); ph = p0.astype(np.float64).copy(); host.update(ph, 0.5 * g); host.update(ph, 0.5 * g)
    d_p, d_g = torch.tensor(p0, device="cuda"), torch.tensor(g, device="cuda")
    d_m, d_v = torch.zeros(240, device="cuda"), torch.zeros(240, device="cuda")
    dev = K.Adam(1e-2)
    for _ in range(2):
        dev.update_dev(node.ode, d_p.data_ptr(), d_g.data_ptr(), d_m.data_ptr(), d_v.data_ptr(), grad_scale=0.5)
    node.ode.lib.kanode_sync(node.ode.h)
    assert np.allclose(d_p.cpu().numpy(), ph, rtol=2e-5, atol=1e-6)


def test_scheduled_backward_with_long_trajectory_warps_matches_plain(lv_saveat):
    """Second call with the same batch: the step margins recorded by the first call decide the launch order of the adjoint
    solves (trajectories sorted by margin: risky ones first, kanode_lg.cu).  The per-trajectory arithmetic does not depend on
    the order and the per-warp gradient partials are summed in a fixed order: identical step counts, gradients equal to
    summation round-off, and the long solves still match the oracle."""
    chain = lv_chain()
    p = glorot_params(chain, seed=0)
    B = 8192
    rng = np.random.default_rng(77)
    u0 = rng.uniform(0.5, 2.0, (B, 2))
    tg = rng.uniform(0.0, 3.0, (B, 35, 2))
    ode = K.KanOde(chain, dtype=np.float64); ode.set_params(p)
    r1 = ode.loss_grad(u0, TSPAN, lv_saveat, tg)          # no history: plain launch
    r2 = ode.loss_grad(u0, TSPAN, lv_saveat, tg)          # ordered launch
    r3 = ode.loss_grad(u0, TSPAN, lv_saveat, tg)
    att = r1["bwd_stats"].naccept + r1["bwd_stats"].nreject
    assert (att > att.mean() + 4).sum() > 0, "workload has no long trajectories; the test would not exercise the warp kernel"
    for r in (r2, r3):
        assert np.array_equal(r["bwd_stats"].naccept, r1["bwd_stats"].naccept)
        assert np.array_equal(r["bwd_stats"].nreject, r1["bwd_stats"].nreject)
        assert np.array_equal(r["bwd_stats"].nf, r1["bwd_stats"].nf)
        assert _relmax(r["grad"], r1["grad"]) < 1e-11 and _relmax(r["du0"], r1["du0"]) < 1e-11
        assert abs(r["loss"] - r1["loss"]) < 1e-13 * r1["loss"]
    # against the oracle on the whole batch: two fp64 implementations (different FMA contraction / summation order) keep
    # the same step sequence except on a few ill-conditioned long solves where round-off is amplified past an
    # accept/reject or tstop-clipping decision
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg)
    same = (r2["bwd_stats"].naccept == ref["bwd_stats"][:, 0]) & (r2["fwd_stats"].naccept == ref["fwd_stats"][:, 0])
    print(f"fp64 CUDA vs oracle: identical fwd+bwd accepted-step counts on {same.mean() * 100:.3f}% of {B} trajectories")
    assert same.mean() > 0.995
    assert _relmax(r2["du0"][same], ref["du0"][same]) < 1e-7
    long_idx = np.argsort(-att)[:64]
    assert same[long_idx].mean() > 0.8
    assert _relmax(r2["grad"], ref["grad"]) < 1e-4          # the few diverged solves differ at solver accuracy
    # fp32: the scheduled path gives the same gradient to rounding
    ode32 = K.KanOde(chain, dtype=np.float32); ode32.set_params(p)
    a = ode32.loss_grad(u0, TSPAN, lv_saveat, tg); b = ode32.loss_grad(u0, TSPAN, lv_saveat, tg)
    assert _relmax(b["grad"], a["grad"].astype(np.float64)) < 1e-4


def test_target_copy_overlapping_the_forward_solve_gives_the_same_step(lv_saveat, monkeypatch):
    """Host entry point: the target travels on a second stream while the forward solve runs on u0; dL/du(t_s) and the loss
    are then formed from the stored predictions (loss_dg_kernel) instead of inside the forward kernel.  Same arithmetic on the
    same values: identical gradient and statistics, loss equal up to the order of the atomic partial sums."""
    chain = lv_chain()
    p = glorot_params(chain, seed=0)
    rng = np.random.default_rng(41)
    u0 = rng.uniform(0.5, 2.0, (1500, 2)); tg = rng.uniform(0.0, 3.0, (1500, 35, 2))
    res = {}
    for mode in ("0", "2"):                                   # 0: every copy on the main stream, 2: overlap at any size
        monkeypatch.setenv("KANODE_OVERLAP_H2D", mode)
        for dt in (np.float32, np.float64):
            ode = K.KanOde(chain, dtype=dt); ode.set_params(p)
            res[mode, dt] = ode.loss_grad(u0, TSPAN, lv_saveat, tg)
            ode.close()
    for dt in (np.float32, np.float64):
        a, b = res["0", dt], res["2", dt]
        assert np.array_equal(a["grad"], b["grad"]) and np.array_equal(a["du0"], b["du0"])
        assert np.array_equal(a["bwd_stats"].nf, b["bwd_stats"].nf) and np.array_equal(a["fwd_stats"].nf, b["fwd_stats"].nf)
        assert abs(a["loss"] - b["loss"]) <= 1e-6 * abs(a["loss"])


def test_lean_loss_grad_call_matches_full(lv_saveat):
    """`want_du0=False, want_stats=False` (the Zygote.gradient(loss, p) shape of the call) returns the same loss / gradient."""
    chain = lv_chain()
    p = glorot_params(chain, seed=0)
    u0 = np.random.default_rng(5).uniform(0.5, 2.0, (300, 2))
    tg = lv_targets(u0[:1], lv_saveat).repeat(300, axis=0)
    ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
    full = ode.loss_grad(u0, TSPAN, lv_saveat, tg)
    lean = ode.loss_grad(u0, TSPAN, lv_saveat, tg, want_du0=False, want_stats=False)
    ode.close()
    assert set(lean) == {"loss", "grad", "solver_failed"} and lean["loss"] == full["loss"] and np.array_equal(lean["grad"], full["grad"])


def test_device_resident_training_matches_the_host_loop():
    """kanode_train_step_dev (grad -> Adam -> loss_train -> loss_test on the device, LV_driver_KANODE.jl:280-291) against the same
    iteration driven from the host with a round trip per call (host-pointer loss_grad, the Adam kernel on the copied gradient,
    parameters read back and set again, two host-pointer forward solves): 100-iteration loss curves equal to 1e-6."""
    import torch
    from examples.train_lv import lv_data
    t, X = lv_data()
    chain = lv_chain()
    p0 = (glorot_params(chain, 0).astype(np.float64) / 1e5).astype(np.float32)
    u0 = np.array([[1.0, 1.0]], np.float32)
    Xtr = X[:, :35].T[None].astype(np.float32); Xte = X.T[None].astype(np.float32)
    eta, iters = 5e-4, 100                                       # Flux.Adam(5e-4), LV_driver_KANODE.jl:219
    # host-driven loop
    ode = K.KanOde(chain, dtype=np.float32)
    opt = K.Adam(eta)
    d_p = torch.tensor(p0, device="cuda"); d_m = torch.zeros_like(d_p); d_v = torch.zeros_like(d_p)
    p = p0.copy()
    host = np.zeros((iters, 3))
    for k in range(iters):
        ode.set_params(p)
        r = ode.loss_grad(u0, TSPAN, t[:35], Xtr, want_du0=False, want_stats=False)
        d_g = torch.tensor(r["grad"], device="cuda")
        opt.update_dev(ode, d_p.data_ptr(), d_g.data_ptr(), d_m.data_ptr(), d_v.data_ptr(), grad_scale=1.0)
        ode.lib.kanode_sync(ode.h)
        p = d_p.cpu().numpy()
        ode.set_params(p)
        ptr = ode.solve(u0, TSPAN, t[:35]).array; pte = ode.solve(u0, (0.0, 14.0), t).array
        host[k] = [r["loss"], np.mean((ptr.astype(np.float64) - Xtr) ** 2), np.mean((pte.astype(np.float64) - Xte) ** 2)]
    # device-resident loop
    ode2 = K.KanOde(chain, dtype=np.float32); ode2.set_params(p0)
    tr = K.DeviceTrainer(ode2, u0, Xtr, TSPAN, t[:35], target_test=Xte, tspan_test=(0.0, 14.0), saveat_test=t, eta=eta)
    for _ in range(iters):
        tr.step()
    dev = tr.losses()
    assert dev.shape == (iters, 3) and np.isfinite(dev).all()
    assert np.abs(dev / host - 1).max() < 1e-6, np.abs(dev / host - 1).max(axis=0)
    assert np.abs(tr.params() - p).max() < 1e-6 * np.abs(p).max()
    assert dev[-1, 1] < dev[0, 1]                                # it trains
    # the handle serves the host-pointer entry points with the trained parameters afterwards (host copy refreshed lazily)
    assert np.abs(ode2.rhs(u0) - ode.rhs(u0)).max() < 1e-6
    ode.close(); ode2.close()


def test_packed_sums_and_packed_adam_for_data_parallel_training(lv_saveat):
    """kanode_pack_sums_dev ([gradient sum | loss sum | count] in one fp64 buffer = ONE all-reduce per step) and
    kanode_train_apply_packed_dev (Adam with g = packed gradient / packed count read on the device) against the unpacked path."""
    import ctypes as C
    import torch
    from kan_odes_b200.dist import packed_all_reduce, unpack
    chain = lv_chain(); p = glorot_params(chain)
    B = 64
    u0 = np.random.default_rng(9).uniform(0.5, 2.0, (B, 2)).astype(np.float32); tg = lv_targets(u0[:1], lv_saveat).repeat(B, axis=0).astype(np.float32)
    outs = []
    for packed_path in (False, True):
        ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
        lib = ode.lib
        d_u0, d_tg = torch.tensor(u0, device="cuda"), torch.tensor(tg, device="cuda")
        d_g = torch.zeros(240, device="cuda"); d_l = torch.zeros(1, dtype=torch.float64, device="cuda")
        sa = np.ascontiguousarray(lv_saveat)
        f = lib.kanode_loss_grad_dev
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_float,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.kanode_train_begin.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        K.abi.check(lib, ode.h, lib.kanode_train_begin(ode.h, 1e-3, 0.9, 0.999, 1e-8), "train_begin")
        for _ in range(3):
            K.abi.check(lib, ode.h, f(ode.h, d_u0.data_ptr(), B, 0.0, 3.5, sa.ctypes.data, sa.size, d_tg.data_ptr(), 1e-6, 1e-3, d_l.data_ptr(),
                                      d_g.data_ptr(), None, None, None), "loss_grad_dev")
            lib.kanode_sync(ode.h)                                 # torch's stream and the handle's own stream meet on the host in this test
            if packed_path:
                buf = packed_all_reduce(ode, d_l, d_g, B)          # world size 1: the pack kernel alone
                lib.kanode_sync(ode.h)
                loss, grad, cnt = unpack(buf, lv_saveat.size, 2)
                assert int(cnt.item()) == B and torch.allclose(grad, d_g / B, rtol=1e-6, atol=0)
                assert abs(loss.item() - d_l.item() / (B * lv_saveat.size * 2)) < 1e-15
                lib.kanode_train_apply_packed_dev.argtypes = [C.c_void_p, C.c_void_p]
                K.abi.check(lib, ode.h, lib.kanode_train_apply_packed_dev(ode.h, buf.data_ptr()), "apply_packed")
            else:
                lib.kanode_train_apply_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
                K.abi.check(lib, ode.h, lib.kanode_train_apply_dev(ode.h, d_g.data_ptr(), 1.0 / B), "apply")
            lib.kanode_sync(ode.h)
        pt = np.empty(240, np.float32)
        K.abi.check(lib, ode.h, lib.kanode_train_params(ode.h, pt.ctypes.data_as(C.c_void_p)), "train_params")
        outs.append(pt); ode.close()
    assert np.abs(outs[0] - outs[1]).max() < 2e-7 * np.abs(outs[0]).max() and np.abs(outs[0] - p).max() > 1e-4
