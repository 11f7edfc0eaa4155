"""Host-side mirror of the reference's layer surface (Python stands in for Julia here:
the Julia toolchain is absent in the build image; julia/KANODEsB200.jl is the
same surface over the same C ABI).

Mirrors, name for name:
  KDense(in_dims, out_dims, grid_len; normalizer, grid_lims, denominator, basis_func,
         base_act, use_base_act, init_C, init_W, allow_fast_activation)   LV/src/kdense.jl:20-68
  initialparameters / initialstates / parameterlength / statelength       LV/src/kdense.jl:70-107
  Chain(layers...), setup(rng, chain) -> (ps, st)                          LV/LV_driver_KANODE.jl:139-143
  flat parameter vector == ComponentArray(pM) data                        LV/LV_driver_KANODE.jl:173-175

The layer itself holds no compute: evaluation goes through the CUDA library (node.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import numpy as np

from . import abi

# ---- names the reference scripts pass as kwargs (utils.jl:8-62, NNlib activations) ----------


class _Named:
    def __init__(self, name: str, code: int):
        self.name, self.code = name, code

    def __repr__(self) -> str:  # pragma: no cover
        return self.name


rbf = _Named("rbf", abi.BASIS_RBF)
rswaf = _Named("rswaf", abi.BASIS_RSWAF)
iqf = _Named("iqf", abi.BASIS_IQF)

tanh = _Named("tanh", abi.NORM_TANH)
tanh_fast = _Named("tanh_fast", abi.NORM_TANH)          # NNlib.fast_act(tanh) == tanh_fast (kdense.jl:57-61)
softsign = _Named("softsign", abi.NORM_SOFTSIGN)
sigmoid = _Named("sigmoid", abi.NORM_SIGMOID)
sigmoid_fast = _Named("sigmoid_fast", abi.NORM_SIGMOID)
swish = _Named("swish", 0)


def glorot_uniform(rng: np.random.Generator, *dims: int) -> np.ndarray:
    """[EXT WeightInitializers 1.0.4] (rand(Float32, dims) - 0.5) * sqrt(24 / (fan_in + fan_out)).

    Returns a column-major (Fortran-order) float32 array like Julia's.  The Julia Xoshiro stream is
    not reproduced (parity tests always pass `p` explicitly).
    """
    fan_out, fan_in = dims[0], dims[1]
    scale = np.float32(np.sqrt(np.float32(24.0) / np.float32(fan_in + fan_out)))
    r = rng.random(dims, dtype=np.float32)
    return np.asfortranarray((r - np.float32(0.5)) * scale)


@dataclass
class KDense:
    in_dims: int
    out_dims: int
    grid_len: int
    normalizer: _Named = tanh
    grid_lims: tuple = (-1.0, 1.0)
    denominator: float | None = None
    basis_func: _Named = rbf
    base_act: _Named = swish
    use_base_act: bool = True
    init_C: Callable = glorot_uniform
    init_W: Callable = glorot_uniform
    allow_fast_activation: bool = True

    def __post_init__(self):
        if self.grid_len < 2:
            raise ValueError("grid_len must be >= 2")
        if not (self.grid_lims[1] > self.grid_lims[0]):          # kdense.jl:50-51
            raise AssertionError("grid_span > 0")
        if self.denominator is None:                              # kdense.jl:27
            self.denominator = float(np.float32(2.0 / (self.grid_len - 1)))
        if self.base_act is not swish:
            raise NotImplementedError("only base_act=swish is on the hot path (kdense.jl:31)")
        if self.allow_fast_activation and self.normalizer is tanh:
            self.normalizer = tanh_fast
        if self.allow_fast_activation and self.normalizer is sigmoid:
            self.normalizer = sigmoid_fast

    # LuxCore protocol ------------------------------------------------------------------
    def initialparameters(self, rng: np.random.Generator) -> dict:
        p = {"C": self.init_C(rng, self.out_dims, self.grid_len * self.in_dims)}   # [O, G*I]  kdense.jl:75
        if self.use_base_act:
            p["W"] = self.init_W(rng, self.out_dims, self.in_dims)                 # kdense.jl:81
        return p

    def initialstates(self, rng=None) -> dict:
        g = np.arange(self.grid_len, dtype=np.float64) / (self.grid_len - 1)
        grid = ((1.0 - g) * np.float32(self.grid_lims[0]) + g * np.float32(self.grid_lims[1])).astype(np.float32)
        return {"grid": grid}                                                      # kdense.jl:88-92

    def parameterlength(self) -> int:                                              # kdense.jl:98-107
        n = self.in_dims * self.grid_len * self.out_dims
        if self.use_base_act:
            n += self.in_dims * self.out_dims
        return n

    def statelength(self) -> int:                                                  # kdense.jl:94-96
        return self.grid_len

    def __call__(self, x, p, st=None, *, device: int = 0, dtype=None):
        """(l::KDense)(x, p, st) -> (y, st)  (kdense.jl:109-130; direct calls at Activation_getter.jl:39,
        Allen-Cahn_Source.jl:91).  x: [in_dims] or [K, in_dims] (one sample per row; Julia's columns), p: {"C": [O, G*I],
        "W": [O, I]} or the layer's flat [vec(C); vec(W)].  Evaluated by the CUDA library through a KANODE_RHS_MAP handle
        (cached on the layer per device / dtype)."""
        from .node import KanOde
        x = np.asarray(x)
        dt = np.dtype(dtype if dtype is not None else (np.float64 if x.dtype == np.float64 else np.float32))
        key = (device, dt.str)
        cache = self.__dict__.setdefault("_map_handles", {})
        if key not in cache:
            cache[key] = KanOde(Chain(self), rhs_kind=abi.RHS_MAP, device=device, dtype=dt)
        ode = cache[key]
        flat = flatten_params({"layer_1": p}) if isinstance(p, dict) else np.asarray(p)
        ode.set_params(flat.astype(dt))
        y = ode.rhs(x.reshape(-1, self.in_dims))
        return (y[0] if x.ndim == 1 else y), st

    def _fill(self, d: abi.LayerDesc) -> None:
        d.in_dims, d.out_dims, d.grid_len = self.in_dims, self.out_dims, self.grid_len
        d.normalizer, d.basis = self.normalizer.code, self.basis_func.code
        d.use_base_act = int(self.use_base_act)
        d.grid_lo, d.grid_hi = float(self.grid_lims[0]), float(self.grid_lims[1])
        d.denominator = float(self.denominator)
        d.kind, d.dense_act = abi.LAYER_KDENSE, 0


identity = _Named("identity", abi.ACT_IDENTITY)


@dataclass
class Dense:
    """`Lux.Dense(in => out, act)`: y = act(W x + b) — the layers of the MLP-NODE baseline
    `Lux.Chain(Lux.Dense(2 => 50, tanh), Lux.Dense(50 => 2))` (Lotka-Volterra/LV_driver_MLP.jl:61).  Parameters
    `{"weight": [out, in], "bias": [out]}`; flat order [vec(weight); bias] like `getdata(ComponentArray(p_))` (:65-67).
    `act` is `tanh` or `identity` (the last layer of a chain must be `identity`)."""
    in_dims: int
    out_dims: int
    act: _Named = None
    use_base_act: bool = True      # Chain.parameterlength / unflatten use the KDense protocol names

    def __post_init__(self):
        if self.act is None:
            self.act = identity
        if self.act in (tanh, tanh_fast):
            self.act = _Named("tanh", abi.ACT_TANH)
        elif self.act is not identity and self.act.name != "tanh":
            raise NotImplementedError("Dense activations on the hot path: tanh, identity")
        self.grid_len = 0

    def initialparameters(self, rng: np.random.Generator) -> dict:
        return {"weight": glorot_uniform(rng, self.out_dims, self.in_dims), "bias": np.zeros(self.out_dims, np.float32)}

    def initialstates(self, rng=None) -> dict:
        return {}

    def parameterlength(self) -> int:
        return (self.in_dims + 1) * self.out_dims

    def statelength(self) -> int:
        return 0

    def _fill(self, d: abi.LayerDesc) -> None:
        d.in_dims, d.out_dims, d.grid_len = self.in_dims, self.out_dims, 0
        d.normalizer = d.basis = d.use_base_act = 0
        d.grid_lo, d.grid_hi, d.denominator = -1.0, 1.0, 1.0
        d.kind, d.dense_act = abi.LAYER_DENSE, self.act.code


@dataclass
class Chain:
    """Lux.Chain of KDense layers; layer names are layer_1, layer_2, ... (LV_driver_KANODE.jl:71)."""
    layers: Sequence[KDense] = field(default_factory=list)

    def __init__(self, *layers: KDense):
        self.layers = list(layers)
        if not 1 <= len(self.layers) <= abi.KANODE_MAX_LAYERS:
            raise ValueError("1..8 layers")
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            if a.out_dims != b.in_dims:
                raise ValueError("layer dims do not chain")

    def parameterlength(self) -> int:
        return sum(l.parameterlength() for l in self.layers)

    def desc(self, rhs_kind: int = abi.RHS_CHAIN, n_state: int | None = None,
             lap_coef: float = 0.0, dx: float = 1.0) -> abi.Desc:
        d = abi.Desc()
        d.n_layers = len(self.layers)
        for i, l in enumerate(self.layers):
            l._fill(d.layers[i])
        d.rhs_kind = rhs_kind
        d.n_state = self.layers[0].in_dims if n_state is None else int(n_state)
        d.lap_coef, d.dx = float(lap_coef), float(dx)
        return d


def activation_getter(chain: Chain, p, X, *, device: int = 0, dtype=np.float64):
    """LV/Activation_getter.jl:3-63 for any chain: the per-edge activations of every layer at the inputs the samples X [K, I_1]
    drive through the chain.  Returns (acts, layer_inputs): acts[l][k, i, o] (kanode_edge_activations), layer_inputs[l] [K, I_l].
    For the reference's two-layer LV model: activations_x = acts[0][:, 0, :], activations_y = acts[0][:, 1, :] ([K, O]), and
    its activations_second[2*(i-1)+o, k] = acts[1][k, i, o]."""
    from .node import KanOde
    dt = np.dtype(dtype)
    ode = KanOde(chain, rhs_kind=abi.RHS_MAP, device=device, dtype=dt)
    ode.set_params(np.asarray(p, dtype=dt))
    x = np.asarray(X, dtype=dt).reshape(-1, chain.layers[0].in_dims)
    acts, inputs = [], []
    for l in range(len(chain.layers)):
        inputs.append(x)
        a = ode.edge_activations(l, x)
        acts.append(a)
        x = a.sum(axis=1)                                             # the layer output (Activation_getter.jl:33-36,39)
    ode.close()
    return acts, inputs


def prune(chain: Chain, p, X, theta: float = 1e-2, *, device: int = 0):
    """prune (LV_driver_KANODE.jl:52-108) for a two-layer chain [I, H, O]: a hidden node is kept when both its largest incoming
    and its largest outgoing activation magnitude over the samples X exceed theta (gamma_pr = 1e-2 in the manuscript).
    Returns (new_chain, new_flat_params, nodes_to_keep).  (The reference fills the pruned layer-2 W from layer_2.C columns
    (:103) — an indexing slip; here W2 keeps its own columns.)"""
    if len(chain.layers) != 2:
        raise ValueError("prune handles the reference's two-layer KAN")
    l1, l2 = chain.layers
    acts, _ = activation_getter(chain, p, X, device=device)
    keep = []
    for j in range(l1.out_dims):
        input_score = np.abs(acts[0][:, :, j]).max()                  # over samples and inputs (:78)
        output_score = np.abs(acts[1][:, j, :]).max()                 # over samples and outputs (:79)
        if min(input_score, output_score) > theta:
            keep.append(j)
    if not keep:
        raise ValueError("every hidden node would be pruned")
    ps = unflatten_params(chain, np.asarray(p))
    G = l1.grid_len
    c2cols = np.concatenate([np.arange(j * G, (j + 1) * G) for j in keep])
    new = {"layer_1": {"C": ps["layer_1"]["C"][keep, :]}, "layer_2": {"C": ps["layer_2"]["C"][:, c2cols]}}
    if l1.use_base_act:
        new["layer_1"]["W"] = ps["layer_1"]["W"][keep, :]
    if l2.use_base_act:
        new["layer_2"]["W"] = ps["layer_2"]["W"][:, keep]
    import dataclasses
    nc = Chain(dataclasses.replace(l1, out_dims=len(keep)), dataclasses.replace(l2, in_dims=len(keep)))
    flat = np.concatenate([np.asarray(new[n][k]).reshape(-1, order="F") for n in ("layer_1", "layer_2") for k in ("C", "W") if k in new[n]])
    return nc, flat.astype(np.asarray(p).dtype), keep


def reg_loss(p, act_reg: float = 1.0, entropy_reg: float = 1.0) -> float:
    """reg_loss (LV_driver_KANODE.jl:187-194) on the host — a reference for KanOde.reg_loss / set_regularizer."""
    a = np.abs(np.asarray(p, dtype=np.float64)); S = a.sum(); e = a[a > 0] / S
    return float(S * act_reg - (e * np.log(e)).sum() * entropy_reg)


def setup(rng: np.random.Generator, chain: Chain):
    """Lux.setup(rng, chain) -> (ps, st): nested dicts keyed layer_1.. like the NamedTuple."""
    ps = {f"layer_{i + 1}": l.initialparameters(rng) for i, l in enumerate(chain.layers)}
    st = {f"layer_{i + 1}": l.initialstates(rng) for i, l in enumerate(chain.layers)}
    return ps, st


def flatten_params(ps: dict) -> np.ndarray:
    """getdata(ComponentArray(pM)): [vec(C1); vec(W1); vec(C2); vec(W2); ...] column-major vec."""
    parts = []
    for name in sorted(ps, key=lambda s: int(s.split("_")[1])):
        for key in ("C", "W", "weight", "bias"):
            if key in ps[name]:
                parts.append(np.asarray(ps[name][key], dtype=np.float32).reshape(-1, order="F"))
    return np.concatenate(parts)


def unflatten_params(chain: Chain, p: np.ndarray) -> dict:
    """ComponentArray(p, pM_axis): inverse of flatten_params."""
    p = np.asarray(p)
    out, off = {}, 0
    for i, l in enumerate(chain.layers):
        d = {}
        if isinstance(l, Dense):
            n = l.out_dims * l.in_dims
            d["weight"] = p[off:off + n].reshape((l.out_dims, l.in_dims), order="F"); off += n
            d["bias"] = p[off:off + l.out_dims]; off += l.out_dims
            out[f"layer_{i + 1}"] = d
            continue
        n = l.out_dims * l.grid_len * l.in_dims
        d["C"] = p[off:off + n].reshape((l.out_dims, l.grid_len * l.in_dims), order="F"); off += n
        if l.use_base_act:
            n = l.out_dims * l.in_dims
            d["W"] = p[off:off + n].reshape((l.out_dims, l.in_dims), order="F"); off += n
        out[f"layer_{i + 1}"] = d
    if off != p.size:
        raise ValueError(f"expected {off} parameters, got {p.size}")
    return out
