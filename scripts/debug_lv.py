import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import kan_odes_b200 as K
from conftest import lv_chain, glorot_params, lv_targets
from oracle import Oracle
chain = lv_chain(); p = glorot_params(chain, 0)
sa = np.arange(35) * 0.1
u0 = np.random.default_rng(1234).uniform(0.5, 2.0, (4, 2))
tg = lv_targets(u0, sa)
ref = Oracle(chain.desc()).loss_grad(p, u0, (0, 3.5), sa, tg)
for dt in (np.float64, np.float32):
    node = K.NeuralODE(chain, (0, 3.5), K.Tsit5(), saveat=sa, dtype=dt)
    loss, grad, info = node.loss_and_grad(u0, p, tg)
    print(dt.__name__, 'loss', loss, ref['loss'])
    print(' fwd', info['fwd_stats'], ref['fwd_stats'].T)
    print(' bwd', info['bwd_stats'], ref['bwd_stats'].T)
    print(' du0', info['du0'], ref['du0'])
    print(' grad', grad[:6], ref['grad'][:6], np.abs(grad).max())
