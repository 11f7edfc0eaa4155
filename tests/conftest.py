import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# ---- shared model builders (shapes of BASELINE.json configs at oracle-friendly sizes) ----------
from kan_odes_b200 import abi  # noqa: E402
from kan_odes_b200.layers import (Chain, KDense, flatten_params, rbf, setup, softsign,  # noqa: E402
                                  tanh_fast)


def lv_chain(width=10, grid=5):
    """Lotka-Volterra/LV_driver_KANODE.jl:130-142"""
    return Chain(KDense(2, width, grid, use_base_act=True, basis_func=rbf, normalizer=tanh_fast),
                 KDense(width, 2, grid, use_base_act=True, basis_func=rbf, normalizer=tanh_fast))


def surrogate_chain(n, hidden=10, grid=5):
    """PDE examples/Burgers_Surrogate.jl:82-88"""
    return Chain(KDense(n, hidden, grid, use_base_act=True, basis_func=rbf, normalizer=softsign),
                 KDense(hidden, n, grid, use_base_act=True, basis_func=rbf, normalizer=softsign))


def source_chain(grid=10):
    """PDE examples/Allen-Cahn_Source.jl:76-81"""
    return Chain(KDense(1, 1, grid, use_base_act=True, basis_func=rbf, normalizer=softsign))


def glorot_params(chain, seed=0, scale=1.0):
    ps, _ = setup(np.random.default_rng(seed), chain)
    return (flatten_params(ps) * np.float32(scale)).astype(np.float32)


def lv_true_rhs(t, u, a=1.5, b=1.0, g=1.0, d=3.0):
    """lotka! Lotka-Volterra/LV_driver_KANODE.jl:46-50 with p_ = [1.5, 1, 1, 3] (:118)"""
    return [a * u[0] - b * u[1] * u[0], g * u[0] * u[1] - d * u[1]]


def lv_targets(u0s, saveat):
    from scipy.integrate import solve_ivp
    out = np.empty((len(u0s), len(saveat), 2))
    for i, u0 in enumerate(u0s):
        s = solve_ivp(lv_true_rhs, (0.0, float(saveat[-1]) + 1e-9), u0, method="DOP853", t_eval=saveat,
                      rtol=1e-12, atol=1e-12)
        out[i] = s.y.T
    return out


@pytest.fixture(scope="session")
def lv_saveat():
    return np.arange(35) * 0.1   # t_train, LV_driver_KANODE.jl:116,123-125
