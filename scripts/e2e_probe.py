"""Diagnostic: where the end-to-end time of one kanode_loss_grad call on the LV ensemble goes (C call with / without the optional
outputs, Python wrapper, parameter upload, raw pinned H2D of the targets).  usage: python scripts/e2e_probe.py  (needs a GPU)"""
import sys, time, ctypes as C
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
import numpy as np, torch
import kan_odes_b200 as K
from kan_odes_b200 import abi
import bench
chain, p, u0, tg = bench.make_workload(65536, 1234)
ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
h_u0 = torch.tensor(u0, dtype=torch.float32).pin_memory().numpy()
h_tg = torch.tensor(tg, dtype=torch.float32).pin_memory().numpy()
h_p = np.ascontiguousarray(p, np.float32)
sa = np.ascontiguousarray(bench.SAVEAT)
lib = ode.lib
B = 65536
loss = C.c_float(); grad = np.empty(ode.np_, np.float32); du0 = np.empty((B, 2), np.float32)
fst = (abi.Stats * B)(); bst = (abi.Stats * B)()
def call(full):
    rc = lib.kanode_loss_grad(ode.h, h_u0.ctypes.data, B, 0.0, 3.5, sa.ctypes.data, sa.size, h_tg.ctypes.data, C.c_float(1e-6), C.c_float(1e-3),
                              C.byref(loss), grad.ctypes.data, du0.ctypes.data if full else None, fst if full else None, bst if full else None)
    assert rc == 0
for full in (True, False):
    for _ in range(3): call(full)
    t = time.perf_counter()
    for _ in range(10): call(full)
    print("C call full=%s: %.3f ms" % (full, (time.perf_counter() - t) * 100))
for _ in range(3): ode.loss_grad(h_u0, bench.TSPAN, bench.SAVEAT, h_tg)
t = time.perf_counter()
for _ in range(10): ode.loss_grad(h_u0, bench.TSPAN, bench.SAVEAT, h_tg)
print("python loss_grad: %.3f ms" % ((time.perf_counter() - t) * 100))
t = time.perf_counter()
for _ in range(10): ode.set_params(h_p)
print("set_params: %.3f ms" % ((time.perf_counter() - t) * 100))
# raw H2D
d = torch.empty(h_tg.size, dtype=torch.float32, device='cuda'); src = torch.from_numpy(h_tg)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): d.copy_(src.view(-1), non_blocking=True); torch.cuda.synchronize()
print("H2D 18.4MB pinned: %.3f ms" % ((time.perf_counter() - t) * 100))
