"""Pins the CPU oracle to the REAL reference when a Julia dump is present (closes "parity unpinned").

`julia --project=<reference>/Lotka-Volterra julia/dump_reference.jl <reference root>` runs the unmodified reference (LV/src,
LV_driver_KANODE.jl:111-203,284 and Burgers_Surrogate.jl:82-107,191) at the fixed inputs of tests/golden/julia_inputs.json and
writes tests/golden/julia_{lv_init,lv_dyn,burgers}.json.  With those files in place this module compares the fp64 oracle field
by field; without them (no Julia in the build image) the comparison tests skip and only the harness self-check runs.

What each field verifies (SURVEY.md §8a / §8c checklist, the items recalled from un-vendored packages):
  sol_u                      a4-a8 KDense forward / rbf / tanh_fast(::Float64) / swish, a10 Tsit5 tableau, a15 dense output
  naccept, nreject, nf       a11 error norm, a12 PI controller + FastPower.fastpower, a13 initial dt, nf accounting (A.3)
  step_t (free solve)        a12-a14 step by step: every accepted step time, tstop clipping at the end of the span
  zgrad                      a6 rrule(_rbf), a16 InterpolatingAdjoint (jump / FSAL order at the save times, norm over [lambda; g])
  loss                       a17
"""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import lv_chain, surrogate_chain
from oracle import Oracle

GOLDEN = Path(__file__).resolve().parent / "golden"
TSPAN = (0.0, 3.5)
SAVEAT = np.arange(35) * 0.1


def _inputs():
    return {k: np.asarray(v, dtype=np.float64) for k, v in json.loads((GOLDEN / "julia_inputs.json").read_text()).items()}


def compare_lv(dump: dict, p: np.ndarray, rtol_state=1e-9, rtol_grad=1e-6):
    """Oracle vs one LV dump; returns the measured discrepancies (asserts inside)."""
    chain = lv_chain()
    orc = Oracle(chain.desc(), np.float64)
    u0 = np.array([[1.0, 1.0]])
    sol_u = np.asarray(dump["sol_u"], dtype=np.float64).reshape(2, -1, order="F")            # Array(sol): [n, nsave]
    tg = np.asarray(dump["target"], dtype=np.float64).reshape(2, -1, order="F").T[None]      # [1, nsave, n]
    assert np.allclose(np.asarray(dump["sol_t"]), SAVEAT, rtol=0, atol=1e-12)
    out, st, step_t = orc.solve(p, u0, TSPAN, SAVEAT, step_cap=256)
    e_state = np.abs(out[0].T - sol_u).max() / np.abs(sol_u).max()
    assert e_state < rtol_state, f"states differ: {e_state:.3e}"
    assert (int(st[0, 0]), int(st[0, 1]), int(st[0, 2])) == (int(dump["naccept"]), int(dump["nreject"]), int(dump["nf"])), \
        f"forward statistics differ: oracle {st[0, :3]}, reference {(dump['naccept'], dump['nreject'], dump['nf'])}"
    ref_steps = np.asarray(dump["step_t"], dtype=np.float64)[1:]                              # sol.t of the free solve minus t0
    mine = step_t[0][~np.isnan(step_t[0])]
    assert mine.size == ref_steps.size and np.abs(mine - ref_steps).max() < 1e-9 * TSPAN[1], "accepted-step times differ"
    r = orc.loss_grad(p, u0, TSPAN, SAVEAT, tg)
    assert abs(r["loss"] - float(dump["loss"])) < 1e-9 * max(float(dump["loss"]), 1e-300)
    zg = np.asarray(dump["zgrad"], dtype=np.float64)
    e_grad = np.abs(r["grad"] - zg).max() / np.abs(zg).max()
    assert e_grad < rtol_grad, f"gradient differs: {e_grad:.3e}"
    return dict(states=e_state, grad=e_grad)


@pytest.mark.parametrize("tag,key", [("lv_init", "lv_p_init"), ("lv_dyn", "lv_p_dyn")])
def test_oracle_matches_julia_dump_lv(tag, key):
    f = GOLDEN / f"julia_{tag}.json"
    if not f.exists():
        pytest.skip(f"{f.name} absent: run julia/dump_reference.jl on a machine with Julia (parity stays unpinned until then)")
    print(compare_lv(json.loads(f.read_text()), _inputs()[key]))


def test_oracle_matches_julia_dump_burgers():
    f = GOLDEN / "julia_burgers.json"
    if not f.exists():
        pytest.skip(f"{f.name} absent: run julia/dump_reference.jl on a machine with Julia")
    d = json.loads(f.read_text()); inp = _inputs()
    n = 41; chain = surrogate_chain(n)
    sa = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9])
    p = inp["burgers_p"].astype(np.float32).astype(np.float64)                                 # Float32 parameters (Burgers_Surrogate.jl:159)
    u0 = inp["burgers_u0"][None]
    tg = inp["burgers_target"].reshape(n, sa.size, order="F").T[None]
    orc = Oracle(chain.desc(), np.float64)
    out, st = orc.solve(p, u0, (0.0, 1.0), sa)
    sol_u = np.asarray(d["sol_u"], dtype=np.float64).reshape(n, -1, order="F")
    assert np.abs(out[0].T - sol_u).max() < 1e-6 * np.abs(sol_u).max()                        # Float32 parameters promote inside the reference
    assert (int(st[0, 0]), int(st[0, 1]), int(st[0, 2])) == (int(d["naccept"]), int(d["nreject"]), int(d["nf"]))
    r = orc.loss_grad(p, u0, (0.0, 1.0), sa, tg)
    zg = np.asarray(d["zgrad"], dtype=np.float64)
    assert abs(r["loss"] - float(d["loss"])) < 1e-6 * float(d["loss"])
    assert np.abs(r["grad"] - zg).max() < 1e-5 * np.abs(zg).max()


def test_harness_self_check_with_an_oracle_generated_dump():
    """The comparison code itself, exercised on a dump in the Julia script's format written from the oracle (so that a typo in
    the harness cannot hide behind the skip).  Not a parity statement."""
    inp = _inputs()
    p = inp["lv_p_dyn"]
    chain = lv_chain(); orc = Oracle(chain.desc(), np.float64)
    u0 = np.array([[1.0, 1.0]])
    tg = np.stack([1.0 + 0.3 * np.sin(SAVEAT), 1.0 + 0.2 * np.cos(SAVEAT)], axis=1)[None]
    out, st, step_t = orc.solve(p, u0, TSPAN, SAVEAT, step_cap=256)
    r = orc.loss_grad(p, u0, TSPAN, SAVEAT, tg)
    steps = step_t[0][~np.isnan(step_t[0])]
    dump = {"sol_t": SAVEAT.tolist(), "sol_u": out[0].T.reshape(-1, order="F").tolist(), "naccept": int(st[0, 0]), "nreject": int(st[0, 1]),
            "nf": int(st[0, 2]), "step_t": [0.0] + steps.tolist(), "target": tg[0].T.reshape(-1, order="F").tolist(),
            "loss": r["loss"], "zgrad": r["grad"].tolist()}
    res = compare_lv(json.loads(json.dumps(dump)), p)
    assert res["states"] == 0.0 and res["grad"] == 0.0
    dump["naccept"] += 1
    with pytest.raises(AssertionError, match="statistics"):
        compare_lv(dump, p)
