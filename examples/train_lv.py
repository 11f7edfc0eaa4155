"""Lotka-Volterra KAN-ODE training loop on the B200 path — the body of Lotka-Volterra/LV_driver_KANODE.jl:111-305
(data generation, model, loss, Zygote gradient, Adam update, train/test loss) with the reference's call shapes.

    python examples/train_lv.py [--iters 200] [--dtype f64]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import kan_odes_b200 as K  # noqa: E402


def lv_data(tspan=(0.0, 14.0), timestep=0.1):
    """solve(ODEProblem(lotka!, [1,1], tspan, [1.5,1,1,3]), Tsit5(), abstol=reltol=1e-12, saveat=0.1)  (:111-127)"""
    from scipy.integrate import solve_ivp
    t = np.arange(0.0, tspan[1] + 1e-9, timestep)
    f = lambda _, u: [1.5 * u[0] - u[1] * u[0], u[0] * u[1] - 3.0 * u[1]]
    s = solve_ivp(f, tspan, [1.0, 1.0], method="DOP853", t_eval=t, rtol=1e-12, atol=1e-12)
    return t, s.y                                                   # X: [2, 141]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--lr", type=float, default=5e-4)                # Flux.Adam(5e-4) :219
    a = ap.parse_args()
    dt = np.float64 if a.dtype == "f64" else np.float32
    t, X = lv_data()
    end_index = int(np.floor(len(t) * 3.5 / 14))                     # :123
    t_train = t[:end_index]
    kan1 = K.Chain(K.KDense(2, 10, 5, use_base_act=True, basis_func=K.rbf, normalizer=K.tanh_fast),
                   K.KDense(10, 2, 5, use_base_act=True, basis_func=K.rbf, normalizer=K.tanh_fast))
    pM, stM = K.setup(np.random.default_rng(0), kan1)
    p = (K.flatten_params(pM).astype(np.float64) / 1e5).astype(dt)  # :175
    u0 = np.array([[1.0, 1.0]])
    train_node = K.NeuralODE(kan1, (0.0, 3.5), K.Tsit5(), saveat=t_train, dtype=dt)
    test_node = K.NeuralODE(kan1, (0.0, 14.0), K.Tsit5(), saveat=t, dtype=dt)
    Xtr = X[:, :end_index].T[None]                                   # target[1][nsave][n]
    opt = K.Adam(a.lr)
    for i in range(1, a.iters + 1):
        loss, grad, _ = train_node.loss_and_grad(u0, p, Xtr)         # grad = Zgrad(loss, p)[1]   :284
        opt.update(p, grad)                                          # update!(opt, p, grad)       :287
        if i % 20 == 0 or i == 1:
            pred_test = np.asarray(test_node(u0, p)[0])              # loss_test(p)                :212-214
            print(f"iter {i:5d}  loss_train {loss:.6e}  loss_test {np.mean((X - pred_test)**2):.6e}")
    return loss


if __name__ == "__main__":
    main()
