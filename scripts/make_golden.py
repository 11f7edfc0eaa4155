"""Generates tests/golden/*.npz with the fp64 CPU oracle (oracle/kanode_oracle.cpp).

The reference ships no golden vectors and Julia is not available here, so these fixtures pin the
ORACLE's behaviour (drift detection + a GPU check that does not need the oracle at run time); they do NOT pin the
oracle to the reference ("parity unpinned").  Inputs follow SURVEY.md §8(d).  Run from the repo root:
    python scripts/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

from conftest import glorot_params, lv_chain, lv_targets, source_chain, surrogate_chain  # noqa: E402
from kan_odes_b200 import abi  # noqa: E402
from oracle import Oracle  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def dump(name, chain, kw, p, u0, tspan, saveat, tg):
    orc = Oracle(chain.desc(**kw), np.float64)
    r = orc.loss_grad(p, u0, tspan, saveat, tg, want_out=True)
    lam = np.random.default_rng(11).normal(size=np.asarray(u0).shape)
    ubar, pbar = orc.vjp(p, u0, lam)
    np.savez_compressed(OUT / f"{name}.npz", p=p, u0=u0, tspan=np.array(tspan), saveat=saveat, target=tg, out=r["out"],
                        loss=r["loss"], grad=r["grad"], du0=r["du0"], fwd_stats=r["fwd_stats"], bwd_stats=r["bwd_stats"],
                        rhs=orc.rhs(p, u0), lam=lam, ubar=ubar, pbar=pbar)
    print(name, "loss", r["loss"], "fwd", r["fwd_stats"][0], "bwd", r["bwd_stats"][0])


sa = np.arange(35) * 0.1
chain = lv_chain()
u0 = np.array([[1.0, 1.0]])
tg = lv_targets(u0, sa)
p = glorot_params(chain, seed=0)
dump("lv_cfg1_p_dyn", chain, {}, p, u0, (0.0, 3.5), sa, tg)                       # BASELINE configs[0], glorot seed 0
dump("lv_cfg1_p_init", chain, {}, (p * np.float32(1e-5)).astype(np.float32), u0, (0.0, 3.5), sa, tg)  # ./1e5, LV_driver:175
u0e = np.random.default_rng(1234).uniform(0.5, 2.0, (16, 2))
dump("lv_ensemble16", chain, {}, p, u0e, (0.0, 3.5), sa, lv_targets(u0e, sa))      # slice of configs[1]

n = 41
chain = surrogate_chain(n, 10, 5)
x = np.linspace(-1, 1, n)
u0 = -np.sin(np.pi * x)[None, :]
sb = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9])
dump("burgers41", chain, {}, glorot_params(chain, seed=0), u0, (0.0, 1.0), sb, u0[:, None, :] * np.exp(-sb)[None, :, None])

chain = source_chain(10)
kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=-1e-4, dx=0.05)
u0 = (x**2 * np.cos(np.pi * x))[None, :]
sc = np.linspace(0, 1, 11)
dump("allen_cahn_source41", chain, kw, glorot_params(chain, seed=3), u0, (0.0, 1.0), sc,
     u0[:, None, :] * np.exp(0.5 * sc)[None, :, None])
