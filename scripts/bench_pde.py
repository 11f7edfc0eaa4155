"""Throughput of the generic (block-per-trajectory) path on the BASELINE.json PDE configs (configs[2..4]).
Not the driver's bench contract (that is bench.py on configs[1]); prints one JSON line per config with the device
times of the three kernels of one fwd+adjoint step (CUDA events inside the library).
usage: python scripts/bench_pde.py [burgers1024 ac4096 schrodinger16384 source4096] [--batch B] [--dtype f32|f64]"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import kan_odes_b200 as K  # noqa: E402
from kan_odes_b200 import abi  # noqa: E402


def surrogate(n, G):
    return K.Chain(K.KDense(n, 10, G, normalizer=K.softsign), K.KDense(10, n, G, normalizer=K.softsign))


def make(name, batch, rng):
    if name == "burgers1024":                                   # configs[2]; Burgers_Surrogate.jl:43,68,82-88
        n = 1024; chain = surrogate(n, 5); x = np.linspace(-1, 1, n)
        u0 = -rng.uniform(0.5, 1.5, (batch, 1)) * np.sin(np.pi * x)[None, :]
        sa = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9]); ts = (0.0, 1.0); kw = {}
        tg = u0[:, None, :] * np.exp(-sa)[None, :, None]
    elif name == "ac4096":                                      # configs[3] wide layer; Allen-Cahn_Surrogate.jl:80-87
        n = 4096; chain = surrogate(n, 10); x = np.linspace(-1, 1, n)
        u0 = rng.uniform(0.8, 1.2, (batch, 1)) * (x**2 * np.cos(np.pi * x))[None, :]
        sa = np.array([0.1, 0.3, 0.5, 0.7, 0.9]); ts = (0.0, 1.0); kw = {}
        tg = u0[:, None, :] * (1 - 0.5 * sa)[None, :, None]
    elif name == "schrodinger16384":                            # configs[4]; Schrodinger_Surrogate.jl:68,73,89-96
        n = 32768; chain = surrogate(n, 10); x = np.linspace(-5, 5, 16384)
        u0 = rng.uniform(0.8, 1.2, (batch, 1)) * np.concatenate([2 / np.cosh(x), np.zeros_like(x)])[None, :]
        sa = np.array([0.1, 0.3, 0.5, 0.7, 0.9, 1.1, 1.3, 1.5]); ts = (0.0, np.pi / 2); kw = {}
        tg = u0[:, None, :] * np.cos(sa)[None, :, None]
    elif name == "source4096":                                  # configs[3] hidden source; Allen-Cahn_Source.jl:34-54,76-99
        n = 4096; chain = K.Chain(K.KDense(1, 1, 10, normalizer=K.softsign)); x = np.linspace(-1, 1, n)
        u0 = rng.uniform(0.8, 1.2, (batch, 1)) * (x**2 * np.cos(np.pi * x))[None, :]
        sa = np.linspace(0, 1, 101); ts = (0.0, 1.0)
        kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=-1e-4, dx=2.0 / (n - 1))
        tg = u0[:, None, :] * np.exp(0.5 * sa)[None, :, None]
    else:
        raise SystemExit(f"unknown config {name}")
    ps, _ = K.setup(np.random.default_rng(0), chain)
    return chain, kw, K.flatten_params(ps), u0, ts, sa, tg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["burgers1024"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    dt = np.float32 if a.dtype == "f32" else np.float64
    for name in a.configs:
        chain, kw, p, u0, ts, sa, tg = make(name, a.batch, np.random.default_rng(3))
        ode = K.KanOde(chain, kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0), dtype=dt)
        ode.set_params(p)
        ms = np.zeros((a.steps, 3)); m3 = (C.c_float * 3)()
        r = ode.loss_grad(u0, ts, sa, tg)                         # warm-up (allocations, record growth)
        wall = []
        for i in range(a.steps):
            t0 = time.perf_counter(); r = ode.loss_grad(u0, ts, sa, tg); wall.append(time.perf_counter() - t0)
            ode.lib.kanode_last_timing(ode.h, m3); ms[i] = list(m3)
        nf_f, nf_b = int(r["fwd_stats"].nf.sum()), int(r["bwd_stats"].nf.sum())
        k = ms.mean(0)
        print(json.dumps({"config": name, "batch": a.batch, "dtype": a.dtype, "n": ode.n, "np": ode.np_,
                          "kernel_ms": {"forward": k[0], "backward": k[1], "grad_reduce": k[2]},
                          "ics_per_s_device": a.batch / (k.sum() / 1e3), "e2e_ics_per_s": a.batch / np.mean(wall),
                          "fwd_steps": [int(r["fwd_stats"].naccept.min()), int(r["fwd_stats"].naccept.max())],
                          "bwd_steps": [int(r["bwd_stats"].naccept.min()), int(r["bwd_stats"].naccept.max())],
                          "rhs_evals": {"forward": nf_f, "backward": nf_b},
                          "failed": int((r["fwd_stats"].retcode != 0).sum() + (r["bwd_stats"].retcode != 0).sum()),
                          "loss": r["loss"]}), flush=True)
        ode.close()


if __name__ == "__main__":
    main()
