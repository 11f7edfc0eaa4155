"""Writes tests/golden/julia_inputs.json: the fixed parameters / initial conditions / targets julia/dump_reference.jl feeds to the
unmodified reference, so that both sides start from bit-identical numbers (the Julia RNG stream is never needed).
    python scripts/make_julia_inputs.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import kan_odes_b200 as K  # noqa: E402


def main():
    lv = K.Chain(K.KDense(2, 10, 5, normalizer=K.tanh_fast), K.KDense(10, 2, 5, normalizer=K.tanh_fast))
    ps, _ = K.setup(np.random.default_rng(0), lv)
    p = K.flatten_params(ps).astype(np.float64)
    n = 41
    bg = K.Chain(K.KDense(n, 10, 5, normalizer=K.softsign), K.KDense(10, n, 5, normalizer=K.softsign))
    psb, _ = K.setup(np.random.default_rng(0), bg)
    x = np.linspace(-1, 1, n)
    u0 = -np.sin(np.pi * x); u0[0] = u0[-1] = 0.0               # Burgers_Surrogate.jl:70: [0; prob.u0; 0]
    sa = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9])
    tg = u0[None, :] * np.exp(-sa)[:, None]                     # [nsave, n]; synthetic stand-in for the MethodOfLines data
    out = {"lv_p_init": (p / 1e5).tolist(), "lv_p_dyn": p.tolist(),
           "burgers_p": K.flatten_params(psb).astype(np.float64).tolist(), "burgers_u0": u0.tolist(),
           "burgers_target": tg.T.reshape(-1, order="F").tolist()}   # column-major [n, nsave]
    (ROOT / "tests" / "golden" / "julia_inputs.json").write_text(json.dumps(out))
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
