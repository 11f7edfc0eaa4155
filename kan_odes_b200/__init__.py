"""kan_odes_b200 — B200-native (sm_100a) KAN-ODE hot path: KDense RHS + Tsit5 + interpolating adjoint.

Host-side mirror of the reference's driver surface over the C ABI in include/kanode.h.  (The directory is named
with an underscore so that it is importable; the project name is kan-odes_b200.)
"""
from .abi import KanodeError, load_library  # noqa: F401
from .layers import (Chain, Dense, KDense, identity, activation_getter, flatten_params, glorot_uniform, iqf, prune, rbf, reg_loss, rswaf, setup,  # noqa: F401
                     sigmoid, sigmoid_fast, softsign, swish, tanh, tanh_fast, unflatten_params)
from .checkpoint import load_checkpoint, save_checkpoint  # noqa: F401
from .optim import Adam, DeviceTrainer  # noqa: F401
from .node import KanOde, NeuralODE, ODESolution, SourceODE, Stats, Tsit5  # noqa: F401

__version__ = "0.1.0"
