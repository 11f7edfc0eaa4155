"""The reference's `.mat` training checkpoint (LV_driver_KANODE.jl:251-272 written, :149-159 and prune :59-68 read back):
    p_list     [iterations, np, 1]  parameter history (row j = flat p after iteration j)
    loss       [iterations]         loss_train history
    loss_test  [iterations]         loss_test history
    kan_pred_t, kan_pred_u1, kan_pred_u2   test-span prediction of the current model (times, the two state components)
    size_KAN   [num_layers, layer_width, grid_size]
Written as MAT v5 with scipy.io (MAT.jl's `matread` reads v5 and v7.3 alike), so the reference's plotting scripts
(plotter_*.py / Plotting_*.jl) and its restart / prune path load checkpoints produced here unchanged.  Host-side only: nothing
of the hot path lives here.
"""
from __future__ import annotations

import numpy as np


def save_checkpoint(path, p_list, loss, loss_test, pred_t, pred_u, size_kan) -> None:
    """pred_u: [nsave, 2] (or [2, nsave]) test-span prediction; p_list: sequence of flat parameter vectors."""
    from scipy.io import savemat
    P = np.asarray(p_list, dtype=np.float64)
    P = P.reshape(P.shape[0], -1, 1)
    n = P.shape[0]
    l = np.zeros(n); l[:len(loss)] = np.asarray(loss, dtype=np.float64)[:n]                    # :257-264 (zero padded to len(p_list))
    lt = np.zeros(n); lt[:len(loss_test)] = np.asarray(loss_test, dtype=np.float64)[:n]
    u = np.asarray(pred_u, dtype=np.float64)
    if u.shape[0] == 2 and u.shape[-1] != 2:
        u = u.T
    savemat(str(path), {"p_list": P, "loss": l, "loss_test": lt, "kan_pred_t": np.asarray(pred_t, dtype=np.float64),
                        "kan_pred_u1": u[:, 0].copy(), "kan_pred_u2": u[:, 1].copy(),
                        "size_KAN": np.asarray(size_kan, dtype=np.float64)}, oned_as="column")


def load_checkpoint(path) -> dict:
    """The restart path of the driver (:149-159): p = p_list[end, :, 1], the loss histories, the model size."""
    from scipy.io import loadmat
    m = loadmat(str(path))
    P = np.asarray(m["p_list"], dtype=np.float64)
    if P.ndim == 2:
        P = P[:, :, None]
    return {"p_list": [P[j, :, 0].copy() for j in range(P.shape[0])], "p": P[-1, :, 0].copy(),
            "loss": np.ravel(m["loss"]).tolist(), "loss_test": np.ravel(m["loss_test"]).tolist(),
            "kan_pred_t": np.ravel(m["kan_pred_t"]), "kan_pred_u1": np.ravel(m["kan_pred_u1"]), "kan_pred_u2": np.ravel(m["kan_pred_u2"]),
            "size_KAN": [int(v) for v in np.ravel(m["size_KAN"])]}
