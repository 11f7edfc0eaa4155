"""Golden fixtures (tests/golden/*.npz, made by scripts/make_golden.py with the fp64 oracle; the reference ships none
— SURVEY.md §4).  CPU: the oracle still reproduces them (drift detection).  GPU: the CUDA fp64 path matches them
through the C ABI without the oracle in the loop."""
from pathlib import Path

import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import lv_chain, source_chain, surrogate_chain
from kan_odes_b200 import abi

GOLD = Path(__file__).resolve().parent / "golden"
CASES = {
    "lv_cfg1_p_dyn": (lv_chain, {}),
    "lv_cfg1_p_init": (lv_chain, {}),
    "lv_ensemble16": (lv_chain, {}),
    "burgers41": (lambda: surrogate_chain(41, 10, 5), {}),
    "allen_cahn_source41": (lambda: source_chain(10), dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=41, lap_coef=-1e-4, dx=0.05)),
}


def _rel(a, b):
    return np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    from oracle import Oracle
    mk, kw = CASES[name]
    chain = mk()
    g = np.load(GOLD / f"{name}.npz")
    orc = Oracle(chain.desc(**kw), np.float64)
    r = orc.loss_grad(g["p"], g["u0"], tuple(g["tspan"]), g["saveat"], g["target"], want_out=True)
    assert (r["fwd_stats"] == g["fwd_stats"]).all() and (r["bwd_stats"] == g["bwd_stats"]).all()
    assert _rel(r["out"], g["out"]) < 1e-12 and _rel(r["grad"], g["grad"]) < 1e-11
    assert abs(r["loss"] - float(g["loss"])) < 1e-13 * abs(float(g["loss"]))
    assert _rel(orc.rhs(g["p"], g["u0"]), g["rhs"]) < 1e-14
    ubar, pbar = orc.vjp(g["p"], g["u0"], g["lam"])
    assert _rel(ubar, g["ubar"]) < 1e-13 and _rel(pbar, g["pbar"]) < 1e-13


def test_golden_anchors_of_the_reference_setup():
    """Structural facts of the reference driver that the fixtures must carry (LV_driver_KANODE.jl:116-125,139-143)."""
    g = np.load(GOLD / "lv_cfg1_p_dyn.npz")
    assert g["p"].size == 240 and g["saveat"].size == 35 and np.allclose(g["saveat"][[0, -1]], [0.0, 3.4])
    assert tuple(g["tspan"]) == (0.0, 3.5) and np.allclose(g["u0"], 1.0)
    assert np.allclose(g["target"][0, 0], [1.0, 1.0])                     # X[:, 1] = u0
    # backward solve: 34 interior save times + t0 force >= 35 steps, plus the 1e-6 start-up ramp
    assert g["bwd_stats"][0, 0] >= 35 + 4


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_fp64_matches_golden(name):
    mk, kw = CASES[name]
    chain = mk()
    g = np.load(GOLD / f"{name}.npz")
    ode = K.KanOde(chain, kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0),
                   dtype=np.float64)
    ode.set_params(g["p"])
    assert _rel(ode.rhs(g["u0"]), g["rhs"]) < 1e-12
    ubar, pbar = ode.vjp(g["u0"], g["lam"])
    assert _rel(ubar, g["ubar"]) < 1e-11 and _rel(pbar, g["pbar"]) < 1e-11
    sol = ode.solve(g["u0"], tuple(g["tspan"]), g["saveat"])
    assert (sol.stats.naccept == g["fwd_stats"][:, 0]).all() and (sol.stats.nf == g["fwd_stats"][:, 2]).all()
    assert _rel(sol.array, g["out"]) < 1e-8
    r = ode.loss_grad(g["u0"], tuple(g["tspan"]), g["saveat"], g["target"])
    assert (r["bwd_stats"].naccept == g["bwd_stats"][:, 0]).all() and (r["bwd_stats"].nf == g["bwd_stats"][:, 2]).all()
    assert abs(r["loss"] - float(g["loss"])) < 1e-9 * abs(float(g["loss"]))
    assert _rel(r["grad"], g["grad"]) < 1e-7 and _rel(r["du0"], g["du0"]) < 1e-7
