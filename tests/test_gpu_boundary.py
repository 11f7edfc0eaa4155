"""Round-2 boundary entry points on the GPU against the oracle: the pullback of the solve for an arbitrary loss
(kanode_solve_adjoint: what Zygote.gradient(loss, p) runs for the reference's own loss, LV_driver_KANODE.jl:197-203,284 and
Burgers_Surrogate.jl:105-107,191), the direct layer call (kdense.jl:109-130), per-edge activations / prune
(LV/Activation_getter.jl, LV_driver_KANODE.jl:52-108), the sparsity regulariser (:187-201) and the failed-solve report."""
import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import glorot_params, lv_chain, lv_targets, source_chain, surrogate_chain
from kan_odes_b200 import abi
from oracle import Oracle

pytestmark = pytest.mark.gpu

TSPAN = (0.0, 3.5)


def _relmax(a, b):
    return np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max()


def _nonmse(out, w, v):
    """loss(pred) = sum(w * pred^3 + v * pred) and its cotangent: not expressible through the fused MSE entry point."""
    return float((w * out ** 3 + v * out).sum()), 3 * w * out ** 2 + v


@pytest.mark.parametrize("model", ["lv", "generic_rswaf", "wide256", "source"])
def test_solve_adjoint_matches_oracle_for_a_non_mse_loss(model, lv_saveat):
    rng = np.random.default_rng(11)
    kw = {}
    if model == "lv":
        chain = lv_chain(); u0 = rng.uniform(0.5, 2.0, (64, 2)); sa, ts = lv_saveat, TSPAN
    elif model == "generic_rswaf":
        chain = K.Chain(K.KDense(3, 6, 4, basis_func=K.rswaf, normalizer=K.softsign), K.KDense(6, 3, 4, basis_func=K.rswaf, normalizer=K.softsign))
        u0 = rng.uniform(-0.5, 0.5, (5, 3)); sa = np.array([0.0, 0.3, 0.7, 1.0]); ts = (0.0, 1.0)
    elif model == "wide256":
        n = 256; chain = surrogate_chain(n); x = np.linspace(-1, 1, n)
        u0 = -rng.uniform(0.5, 1.5, (4, 1)) * np.sin(np.pi * x)[None, :]; sa = np.array([0.0, 0.1, 0.5, 0.9]); ts = (0.0, 1.0)
    else:
        n = 41; chain = source_chain(); x = np.linspace(-1, 1, n)
        u0 = rng.uniform(0.8, 1.2, (3, 1)) * (x ** 2 * np.cos(np.pi * x))[None, :]; sa = np.linspace(0, 0.2, 5); ts = (0.0, 0.2)
        kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=1e-4, dx=2.0 / (n - 1))
    p = glorot_params(chain, seed=1).astype(np.float64)
    desc = chain.desc(kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0))
    n_state = int(desc.n_state)
    w = rng.normal(size=(len(u0), len(sa), n_state)); v = rng.normal(size=(len(u0), len(sa), n_state))
    orc = Oracle(desc, np.float64)
    out_ref, _ = orc.solve(p, u0, ts, sa)
    _, cot = _nonmse(out_ref, w, v)
    ref = orc.adjoint(p, u0, ts, sa, cot)
    ode = K.KanOde(chain, kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0), dtype=np.float64)
    ode.set_params(p)
    r = ode.solve_adjoint(u0, ts, sa, cot)
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all() and (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert _relmax(r["out"], ref["out"]) < 2e-8
    assert _relmax(r["grad"], ref["grad"]) < 1e-6 and _relmax(r["du0"], ref["du0"]) < 1e-6
    ode.close()


def test_neural_ode_gradient_of_the_reference_loss_shapes(lv_saveat):
    """`Zygote.gradient(loss, p)` for loss(p) = mean(abs2, X - pred) + reg_loss(p, 5e-4, 0) (LV_driver_KANODE.jl:197-201 with
    sparse_on = 1) differentiated by the caller: the solve's pullback (NeuralODE.gradient) equals the fused MSE entry point plus
    the regulariser, and both equal the oracle."""
    chain = lv_chain(); p = glorot_params(chain).astype(np.float64)
    u0 = np.random.default_rng(2).uniform(0.5, 2.0, (16, 2)); X = lv_targets(u0, lv_saveat)
    node = K.NeuralODE(chain, TSPAN, K.Tsit5(), saveat=lv_saveat, dtype=np.float64)

    def loss_and_cot(pred):                                      # the data term of the reference loss, differentiated by hand
        e = pred - X
        return float((e ** 2).mean()), 2.0 * e / e.size
    val, grad = node.gradient(loss_and_cot, u0, p)
    orc = Oracle(chain.desc(), np.float64)
    ref = orc.loss_grad(p, u0, TSPAN, lv_saveat, X)
    assert abs(val - ref["loss"]) < 1e-10 * ref["loss"] and _relmax(grad, ref["grad"]) < 1e-8
    # the same with the sparsity term through the fused call
    rl, rg = orc.reg_loss(p, 5e-4, 0.0)
    node.ode.set_params(p); node.ode.set_regularizer(5e-4, 0.0)
    r = node.ode.loss_grad(u0, TSPAN, lv_saveat, X)
    assert abs(r["loss"] - (ref["loss"] + rl)) < 1e-10 * (ref["loss"] + rl)
    assert _relmax(r["grad"], ref["grad"] + rg) < 1e-8
    node.ode.set_regularizer(0.0, 0.0)
    # entropy term too, fp32 handle (reg evaluated on the device in fp64 from the fp32 parameters)
    ode32 = K.KanOde(chain, dtype=np.float32); ode32.set_params(p)
    l32, g32 = ode32.reg_loss(1.0, 1.0)
    l64, g64 = orc.reg_loss(p.astype(np.float32).astype(np.float64), 1.0, 1.0)
    assert abs(l32 - l64) < 1e-9 * abs(l64) and _relmax(g32, g64) < 1e-6
    ode32.close()


def test_direct_layer_call_and_edge_activations():
    """(l::KDense)(x, p, st) -> (y, st) and LV/Activation_getter.jl (incl. its 1e-10 identity, :33-36) for the LV model."""
    chain = lv_chain(); p = glorot_params(chain).astype(np.float64)
    ps = K.unflatten_params(chain, p)
    X = np.random.default_rng(3).uniform(0.0, 3.0, (50, 2))
    o = Oracle(chain.desc(abi.RHS_MAP), np.float64)
    l1, l2 = chain.layers
    y1, st = l1(X, ps["layer_1"], {"grid": None})
    assert st == {"grid": None} and y1.shape == (50, 10)
    o1 = Oracle(K.Chain(l1).desc(abi.RHS_MAP), np.float64)
    assert np.abs(y1 - o1.map(p[:l1.parameterlength()], X)).max() < 1e-12
    y2, _ = l2(y1, ps["layer_2"])
    assert np.abs(y2 - o.map(p, X)).max() < 1e-12
    assert l1(X[0].astype(np.float32), ps["layer_1"])[0].shape == (10,)            # one sample, fp32 path
    acts, inputs = K.activation_getter(chain, p, X)
    assert np.abs(acts[0] - o.edge_activations(p, 0, X)).max() < 1e-12
    assert np.abs(acts[1] - o.edge_activations(p, 1, y1)).max() < 1e-12
    assert np.abs(acts[0].sum(axis=1) - y1).max() < 1e-10 and np.abs(acts[1].sum(axis=1) - y2).max() < 1e-10
    # reference shapes: activations_x / activations_y are [K, O]; activations_second rows 2(i-1)+o
    activations_x, activations_y = acts[0][:, 0, :], acts[0][:, 1, :]
    assert activations_x.shape == (50, 10) and activations_y.shape == (50, 10)


def test_prune_drops_dead_hidden_nodes_and_keeps_the_function():
    chain = lv_chain(); p = glorot_params(chain).astype(np.float64)
    ps = K.unflatten_params(chain, p)
    dead = [2, 7]
    for j in dead:                                               # silence nodes 2 and 7 on the output side
        ps["layer_2"]["C"][:, j * 5:(j + 1) * 5] = 0.0; ps["layer_2"]["W"][:, j] = 0.0
    p = np.concatenate([ps[n][k].reshape(-1, order="F") for n in ("layer_1", "layer_2") for k in ("C", "W")])
    X = np.random.default_rng(4).uniform(0.0, 3.0, (200, 2))
    nc, pn, keep = K.prune(chain, p, X, theta=1e-2)
    assert keep == [j for j in range(10) if j not in dead] and nc.layers[0].out_dims == 8 and pn.size == nc.parameterlength()
    full = Oracle(chain.desc(abi.RHS_MAP), np.float64).map(p, X)
    small = Oracle(nc.desc(abi.RHS_MAP), np.float64).map(pn, X)
    assert np.abs(full - small).max() < 1e-12
    ode = K.KanOde(nc, dtype=np.float64); ode.set_params(pn)      # the pruned model is a valid ODE right-hand side
    assert np.abs(ode.rhs(X) - small).max() < 1e-12
    ode.close()


def test_failed_solves_are_reported_not_averaged_in():
    """A trajectory that fails (here: blow-up of an unstable explicit diffusion) makes kanode_loss_grad return
    KANODE_ERR_SOLVER; allow_failed=True hands back the result with the retcodes."""
    n = 4096; chain = source_chain(); x = np.linspace(-1, 1, n)   # the reference's anti-diffusive sign at this resolution (tests/test_gpu_pde.py)
    u0 = np.stack([x ** 2 * np.cos(np.pi * x), 0.5 * x ** 2 * np.cos(np.pi * x)])
    sa = np.linspace(0, 1.0, 5)
    ode = K.KanOde(chain, abi.RHS_SOURCE_LAPLACIAN, n, -1e-4, 2.0 / (n - 1), dtype=np.float64)
    ode.set_params(glorot_params(chain, seed=3).astype(np.float64))
    tg = np.zeros((2, sa.size, n))
    sol = ode.solve(u0, (0.0, 1.0), sa)
    if (sol.stats.retcode == 0).all():
        pytest.skip("the field did not fail on this device")
    with pytest.raises(K.KanodeError, match="SOLVER"):
        ode.loss_grad(u0, (0.0, 1.0), sa, tg)
    r = ode.loss_grad(u0, (0.0, 1.0), sa, tg, allow_failed=True)
    assert r["solver_failed"] and (r["fwd_stats"].retcode != 0).any()
    ode.close()


def test_mlp_node_baseline_matches_oracle(lv_saveat):
    """The MLP-NODE baseline `Lux.Chain(Lux.Dense(2 => 50, tanh), Lux.Dense(50 => 2))` (LV_driver_MLP.jl:61-76) as a second RHS kind:
    RHS, VJP, solve and the adjoint gradient against the oracle (which states the Dense layer in its natural act(Wx + b) form)."""
    mlp = K.Chain(K.Dense(2, 50, K.tanh), K.Dense(50, 2))
    assert mlp.parameterlength() == 252                                   # LV_driver_MLP.jl prints "parameter size: 252"
    rng = np.random.default_rng(4)
    ps, _ = K.setup(np.random.default_rng(0), mlp)
    ps["layer_1"]["bias"] = rng.normal(size=50).astype(np.float32) * 0.3; ps["layer_2"]["bias"] = np.array([0.1, -0.2], np.float32)
    p = K.flatten_params(ps).astype(np.float64)
    assert np.array_equal(K.unflatten_params(mlp, p)["layer_2"]["bias"], p[-2:])
    orc = Oracle(mlp.desc(), np.float64)
    u = rng.uniform(0.5, 2.0, (9, 2)); lam = rng.normal(size=(9, 2))
    ode = K.KanOde(mlp, dtype=np.float64); ode.set_params(p)
    assert np.abs(ode.rhs(u) - orc.rhs(p, u)).max() < 1e-13
    ub, pb = ode.vjp(u, lam); ub0, pb0 = orc.vjp(p, u, lam)
    assert np.abs(ub - ub0).max() < 1e-12 and np.abs(pb - pb0).max() < 1e-11
    u0 = rng.uniform(0.5, 2.0, (6, 2)); tg = lv_targets(u0, lv_saveat)
    ref = orc.loss_grad(p, u0, TSPAN, lv_saveat, tg, want_out=True)
    r = ode.loss_grad(u0, TSPAN, lv_saveat, tg)
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all() and (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert (r["bwd_stats"].nf == ref["bwd_stats"][:, 2]).all()
    assert abs(r["loss"] - ref["loss"]) < 1e-10 * ref["loss"] and _relmax(r["grad"], ref["grad"]) < 1e-7
    assert _relmax(ode.solve(u0, TSPAN, lv_saveat).array, ref["out"]) < 1e-9
    ode.close()
    ode32 = K.KanOde(mlp, dtype=np.float32); ode32.set_params(p)
    r32 = ode32.loss_grad(u0, TSPAN, lv_saveat, tg)
    assert _relmax(r32["grad"], ref["grad"]) < 5e-3 and abs(r32["loss"] - ref["loss"]) < 1e-3 * ref["loss"]
    ode32.close()
    # a mixed chain (KDense -> Dense) runs on the same kernels
    mix = K.Chain(K.KDense(2, 6, 4, normalizer=K.softsign), K.Dense(6, 2))
    pm = glorot_params(mix, seed=5).astype(np.float64); pm[-2:] = [0.05, -0.03]
    om = Oracle(mix.desc(), np.float64); odm = K.KanOde(mix, dtype=np.float64); odm.set_params(pm)
    rm = odm.loss_grad(u0, (0.0, 1.0), [0.5, 1.0], tg[:, :2]); rfm = om.loss_grad(pm, u0, (0.0, 1.0), [0.5, 1.0], tg[:, :2])
    assert _relmax(rm["grad"], rfm["grad"]) < 1e-7
    odm.close()
