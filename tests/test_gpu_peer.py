"""The data-parallel step's collective done by the library itself over peer memory (kanode_peer.cu):
kanode_peer_export / kanode_peer_attach / kanode_pack_allreduce_dev.  One kernel packs [gradient sum | loss sum | count],
stores it into every rank's mailbox, waits for the world's epoch flags and sums in rank order.

* world = 1 in-process: the fused kernel equals kanode_pack_sums_dev bit for bit, over several epochs (both parities);
* world = 2, two processes (gloo carries only the 64-byte IPC handles): rank r runs on GPU r when the box has two GPUs,
  otherwise both ranks share cuda:0 (the mailboxes are then same-device IPC mappings and the two kernels time-slice).
  Every rank must see the bit-identical sum of both ranks' sums, equal to the oracle on the whole batch."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

pytestmark = pytest.mark.gpu

TSPAN = (0.0, 3.5)


def _step(ode, torch, u0, tg, sa, dtype):
    """kanode_loss_grad_dev on device tensors -> (loss_sum, grad_sum) tensors (un-normalised sums)."""
    import ctypes as C
    dev = torch.device("cuda", ode.device)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    d_u0 = torch.tensor(u0, dtype=tdt, device=dev); d_tg = torch.tensor(tg, dtype=tdt, device=dev)
    d_grad = torch.zeros(ode.np_, dtype=tdt, device=dev); d_loss = torch.zeros(1, dtype=torch.float64, device=dev)
    creal = C.c_double if dtype == np.float64 else C.c_float
    fn = ode.lib.kanode_loss_grad_dev_f64 if dtype == np.float64 else ode.lib.kanode_loss_grad_dev
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                   creal, creal, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    sa = np.ascontiguousarray(sa, dtype=np.float64)
    torch.cuda.synchronize(dev)                                              # the handle runs on its own stream
    rc = fn(ode.h, d_u0.data_ptr(), u0.shape[0], TSPAN[0], TSPAN[1], sa.ctypes.data, sa.size, d_tg.data_ptr(), 1e-6, 1e-3,
            d_loss.data_ptr(), d_grad.data_ptr(), None, None, None)
    assert rc == 0, rc
    ode.sync()
    return d_loss, d_grad


def test_world_of_one_equals_pack_sums(lv_saveat):
    import torch
    import kan_odes_b200 as K
    from kan_odes_b200 import dist as kd
    from conftest import glorot_params, lv_chain, lv_targets
    chain = lv_chain(); p = glorot_params(chain)
    u0 = np.random.default_rng(2).uniform(0.5, 2.0, (64, 2)); tg = lv_targets(u0, lv_saveat)
    ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
    d_loss, d_grad = _step(ode, torch, u0, tg, lv_saveat, np.float32)
    ref = kd.packed_all_reduce(ode, d_loss, d_grad, 64)                     # not attached yet: kanode_pack_sums_dev
    ode.sync(); ref = ref.clone(); torch.cuda.synchronize()
    assert kd.peer_setup(ode)
    for _ in range(5):                                                       # epochs 1..5: both parities of the mailbox
        got = kd.packed_all_reduce(ode, d_loss, d_grad, 64)
        ode.sync()
        assert torch.equal(got, ref)
    assert ode.lib.kanode_peer_status(ode.h) == 0
    assert float(ref[-1]) == 64.0 and float(ref[-2]) == float(d_loss)
    ode.close()


def _worker(rank, world, port, batch, ngpu, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    import torch
    import torch.distributed as dist
    import kan_odes_b200 as K
    from kan_odes_b200 import dist as kd
    from conftest import glorot_params, lv_chain, lv_targets
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank if ngpu >= world else 0
    torch.cuda.set_device(dev)
    chain = lv_chain(); p = glorot_params(chain).astype(np.float64)
    sa = np.arange(35) * 0.1
    u0 = np.random.default_rng(1234).uniform(0.5, 2.0, (batch, 2)); tg = lv_targets(u0, sa)
    lo, hi = kd.shard_bounds(batch, rank, world)
    ode = K.KanOde(chain, dtype=np.float64, device=dev); ode.set_params(p)
    ok = kd.peer_setup(ode)
    outs = []
    if ok:
        d_loss, d_grad = _step(ode, torch, u0[lo:hi], tg[lo:hi], sa, np.float64)
        for _ in range(4):
            buf = kd.packed_all_reduce(ode, d_loss, d_grad, hi - lo)
            ode.sync()
            outs.append(buf.cpu().numpy().copy())
        st = ode.lib.kanode_peer_status(ode.h)
    else:
        st = -1
    q.put((rank, ok, st, outs))
    dist.barrier()
    ode.close()
    dist.destroy_process_group()


def test_two_ranks_exchange_through_peer_mailboxes(lv_saveat):
    import torch
    import torch.multiprocessing as mp
    from conftest import glorot_params, lv_chain, lv_targets
    from oracle import Oracle
    batch, world = 37, 2
    ngpu = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 11
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, ngpu, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=240) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(ok for _, ok, _, _ in res), "peer mailboxes could not be mapped"
    assert all(st == 0 for _, _, st, _ in res), "a rank timed out waiting for its peer"
    a, b = res[0][3], res[1][3]
    for x, y in zip(a, b):
        assert np.array_equal(x, y) and np.array_equal(x, a[0])              # bit-identical on both ranks, every epoch
    chain = lv_chain(); p = glorot_params(chain).astype(np.float64)
    u0 = np.random.default_rng(1234).uniform(0.5, 2.0, (batch, 2)); tg = lv_targets(u0, lv_saveat)
    ref = Oracle(chain.desc(), np.float64).loss_grad(p, u0, TSPAN, lv_saveat, tg)
    npar = a[0].size - 2
    assert a[0][npar + 1] == batch
    loss = a[0][npar] / (batch * lv_saveat.size * 2); grad = a[0][:npar] / batch
    assert abs(loss - ref["loss"]) < 1e-9 * ref["loss"]
    assert np.abs(grad - ref["grad"]).max() < 1e-7 * np.abs(ref["grad"]).max()
