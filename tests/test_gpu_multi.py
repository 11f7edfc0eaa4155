"""Several GPUs behind one handle (kanode_create_multi; SURVEY.md §8b/§8e: single-process surface, batch sharded inside the
library, the only cross-device step is the peer-memory sum of the gradient partials).  Needs >= 2 visible GPUs."""
import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import glorot_params, lv_chain, lv_targets, surrogate_chain
from oracle import Oracle

pytestmark = pytest.mark.gpu

TSPAN = (0.0, 3.5)


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _relmax(a, b):
    return np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max()


@pytest.mark.parametrize("B", [257, 3])
def test_one_handle_two_devices_matches_single_device_and_oracle(B, lv_saveat):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    chain = lv_chain(); p = glorot_params(chain).astype(np.float64)
    rng = np.random.default_rng(8)
    u0 = rng.uniform(0.5, 2.0, (B, 2)); tg = lv_targets(u0[:1], lv_saveat).repeat(B, axis=0) * rng.uniform(0.9, 1.1, (B, 1, 1))
    one = K.KanOde(chain, dtype=np.float64, device=0); one.set_params(p)
    two = K.KanOde(chain, dtype=np.float64, devices=[0, 1]); two.set_params(p)
    assert two.lib.kanode_device_count(two.h) == 2 and one.lib.kanode_device_count(one.h) == 1
    a = one.loss_grad(u0, TSPAN, lv_saveat, tg); b = two.loss_grad(u0, TSPAN, lv_saveat, tg)
    assert np.array_equal(a["fwd_stats"].naccept, b["fwd_stats"].naccept) and np.array_equal(a["bwd_stats"].nf, b["bwd_stats"].nf)
    assert abs(a["loss"] - b["loss"]) < 1e-13 * a["loss"] and _relmax(b["grad"], a["grad"]) < 1e-12 and _relmax(b["du0"], a["du0"]) < 1e-14
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg)
    assert _relmax(b["grad"], ref["grad"]) < 1e-7
    sa, sb = one.solve(u0, TSPAN, lv_saveat), two.solve(u0, TSPAN, lv_saveat)
    assert np.array_equal(sa.array, sb.array)
    # pullback with caller-supplied cotangents through the same sharding
    cot = rng.normal(size=(B, lv_saveat.size, 2))
    ra, rb = one.solve_adjoint(u0, TSPAN, lv_saveat, cot), two.solve_adjoint(u0, TSPAN, lv_saveat, cot)
    assert _relmax(rb["grad"], ra["grad"]) < 1e-12 and np.array_equal(ra["out"], rb["out"])
    # device-pointer entry points belong to one GPU
    with pytest.raises(K.KanodeError, match="UNSUPPORTED"):
        K.abi.check(two.lib, two.h, two.lib.kanode_solve_dev(two.h, None, 1, 0.0, 1.0, None, 0, 0, 0, None, None), "kanode_solve_dev")
    one.close(); two.close()


def test_wide_model_on_two_devices_fp32():
    if _ngpu() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    n = 256; chain = surrogate_chain(n); p = glorot_params(chain)
    x = np.linspace(-1, 1, n); rng = np.random.default_rng(3)
    u0 = -rng.uniform(0.5, 1.5, (6, 1)) * np.sin(np.pi * x)[None, :]
    sa = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9]); tg = u0[:, None, :] * np.exp(-sa)[None, :, None]
    one = K.KanOde(chain, dtype=np.float32, device=0); one.set_params(p)
    two = K.KanOde(chain, dtype=np.float32, devices=[0, 1]); two.set_params(p)
    a = one.loss_grad(u0, (0.0, 1.0), sa, tg); b = two.loss_grad(u0, (0.0, 1.0), sa, tg)
    assert _relmax(b["grad"], a["grad"].astype(np.float64)) < 1e-5 and abs(a["loss"] - b["loss"]) < 1e-6 * a["loss"]
    one.close(); two.close()
