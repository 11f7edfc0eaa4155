"""CPU checks of the oracle counterparts of the round-2 boundary entry points (test infrastructure checked against
independent statements): cotangent-input adjoint, the chain as a map, per-edge activations, reg_loss."""
import numpy as np

from conftest import glorot_params, lv_chain, lv_targets
from kan_odes_b200 import abi
from kan_odes_b200.layers import Chain, KDense, reg_loss, rswaf, softsign
from oracle import Oracle

TSPAN = (0.0, 3.5)


def test_adjoint_with_mse_cotangent_equals_loss_grad(lv_saveat):
    """kanode_oracle_adjoint fed dL/dpred of the MSE loss reproduces loss_grad (LV_driver_KANODE.jl:197-203,284)."""
    chain = lv_chain(); p = glorot_params(chain).astype(np.float64)
    u0 = np.random.default_rng(1).uniform(0.5, 2.0, (5, 2)); tg = lv_targets(u0, lv_saveat)
    o = Oracle(chain.desc(), np.float64)
    ref = o.loss_grad(p, u0, TSPAN, lv_saveat, tg, want_out=True)
    cot = 2.0 * (ref["out"] - tg) / (2 * lv_saveat.size) / len(u0)          # d mean(abs2) / d pred, 1/B of the batch mean
    r = o.adjoint(p, u0, TSPAN, lv_saveat, cot)
    assert np.allclose(r["out"], ref["out"], rtol=0, atol=0)
    assert np.abs(r["grad"] - ref["grad"]).max() < 1e-12 * np.abs(ref["grad"]).max()
    assert (r["bwd_stats"][:, :3] == ref["bwd_stats"][:, :3]).all()


def test_adjoint_of_a_non_mse_loss_matches_finite_differences(lv_saveat):
    """loss = sum(w * pred^3) + sum(v * pred): its gradient through the oracle adjoint vs central differences of the solve."""
    chain = lv_chain(4, 5); p = glorot_params(chain, seed=3).astype(np.float64)
    rng = np.random.default_rng(5)
    u0 = rng.uniform(0.7, 1.5, (2, 2)); sa = np.array([0.0, 0.4, 1.1, 2.0]); ts = (0.0, 2.0)
    w = rng.normal(size=(2, sa.size, 2)); v = rng.normal(size=(2, sa.size, 2))
    o = Oracle(chain.desc(), np.float64)

    def loss(pp):
        out, _ = o.solve(pp, u0, ts, sa, abstol=1e-12, reltol=1e-10)
        return float((w * out ** 3 + v * out).sum())
    out, _ = o.solve(p, u0, ts, sa, abstol=1e-12, reltol=1e-10)
    r = o.adjoint(p, u0, ts, sa, 3 * w * out ** 2 + v, abstol=1e-12, reltol=1e-10)
    for j in rng.choice(p.size, 6, replace=False):
        e = np.zeros_like(p); e[j] = 1e-6
        fd = (loss(p + e) - loss(p - e)) / 2e-6
        assert abs(fd - r["grad"][j]) < 2e-6 * max(1.0, abs(fd)), (j, fd, r["grad"][j])


def test_edge_activations_sum_to_the_layer_output_and_map_matches_chain():
    """Activation_getter.jl:33-36 (the commented 1e-10 identity) and the direct layer call."""
    for norm, basis in ((softsign, rswaf), (softsign, None)):
        kw = dict(normalizer=norm) | (dict(basis_func=basis) if basis else {})
        chain = Chain(KDense(3, 7, 6, **kw), KDense(7, 2, 6, **kw))
        p = glorot_params(chain, seed=2).astype(np.float64)
        x = np.random.default_rng(0).normal(size=(11, 3))
        o = Oracle(chain.desc(abi.RHS_MAP), np.float64)
        a1 = o.edge_activations(p, 0, x)
        h = a1.sum(axis=1)
        a2 = o.edge_activations(p, 1, h)
        y = o.map(p, x)
        assert np.abs(a2.sum(axis=1) - y).max() < 1e-10
        # single layer as a map: its flat parameters are the first block of the chain's
        l1 = Chain(chain.layers[0])
        o1 = Oracle(l1.desc(abi.RHS_MAP), np.float64)
        assert np.abs(o1.map(p[:l1.parameterlength()], x) - h).max() < 1e-12


def test_reg_loss_value_and_gradient():
    """reg_loss(p, act_reg, entropy_reg) (LV_driver_KANODE.jl:187-194) against the numpy statement and finite differences."""
    p = np.random.default_rng(7).normal(size=240) * 0.3
    o = Oracle(lv_chain().desc(), np.float64)
    for ar, er in ((5e-4, 0.0), (1.0, 1.0), (0.0, 2.0)):
        val, g = o.reg_loss(p, ar, er)
        assert abs(val - reg_loss(p, ar, er)) < 1e-12 * max(1.0, abs(val))
        for j in (0, 17, 239):
            e = np.zeros_like(p); e[j] = 1e-6
            fd = (reg_loss(p + e, ar, er) - reg_loss(p - e, ar, er)) / 2e-6
            assert abs(fd - g[j]) < 1e-7 * max(1.0, abs(fd))
