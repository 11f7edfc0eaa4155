"""Independent array-form statement of the KDense layer in torch (fp64), written from the Julia source the way the
Julia source writes it (LV/src/kdense.jl:109-130, LV/src/utils.jl:8-13): x_norm -> basis [G, I*K] -> reshape
[G*I, K] -> C * basis + W * swish.(x).  Used to cross-check the scalar-loop oracle and, through autograd, the
hand-written reverse pass.  Not the oracle, not the product: test-only."""
import numpy as np
import torch

from kan_odes_b200 import abi


def _norm(kind, x):
    if kind == abi.NORM_TANH:
        return torch.tanh(x)
    if kind == abi.NORM_SOFTSIGN:
        return x / (1 + x.abs())
    return torch.sigmoid(x)


def _basis(kind, y):
    if kind == abi.BASIS_RBF:
        return torch.exp(-y**2)
    if kind == abi.BASIS_RSWAF:
        return 1 - torch.tanh(y)**2
    return 1 / (1 + y**2)


def chain_torch(chain, p, x):
    """x: [I, K] (Julia layout), p: flat torch vector.  Returns [O, K]."""
    off = 0
    for l in chain.layers:
        I, O, G = l.in_dims, l.out_dims, l.grid_len
        C = p[off:off + O * G * I].reshape(G * I, O).T; off += O * G * I       # column-major [O, G*I]
        grid = torch.tensor(l.initialstates()["grid"].astype(np.float64))
        inv_h = float(np.float32(1.0) / np.float32(l.denominator))
        K = x.shape[1]
        xn = _norm(l.normalizer.code, x)                                        # [I, K]
        # Julia reshape(x_norm,1,:) is column-major: element order (i fastest, then k).  torch is row-major,
        # so build [K, I] explicitly:
        xr = xn.T.reshape(-1)                                                    # index k*I + i  == Julia (i,k)
        basis = _basis(l.basis_func.code, (xr[None, :] - grid[:, None]) * inv_h)  # [G, I*K]
        # Julia reshape(basis, G*I, K): column-major => row index g + G*i for column k
        basis = basis.T.reshape(K, I * G).T                                     # [(i*G+g), K]
        y = C @ basis
        if l.use_base_act:
            W = p[off:off + O * I].reshape(I, O).T; off += O * I
            y = y + W @ (x * torch.sigmoid(x))
        x = y
    return x


def rhs_torch(chain, desc, p, u):
    """u: [B, n] -> du [B, n] for either rhs_kind."""
    if desc.rhs_kind == abi.RHS_CHAIN:
        return chain_torch(chain, p, u.T).T
    n = u.shape[1]
    lap = torch.zeros(n, n, dtype=u.dtype)
    idx = torch.arange(n)
    lap[idx, idx] = -2.0
    lap[idx[:-1], idx[1:]] = 1.0
    lap[idx[1:], idx[:-1]] = 1.0
    lap[0, -1] = 1.0; lap[-1, 0] = 1.0                                           # AC_Source:50-54
    lap = lap / desc.dx**2
    kan = chain_torch(chain, p, u.reshape(1, -1)).reshape(u.shape)              # kan1_.(u)
    return (desc.lap_coef * lap @ u.T).T + kan
