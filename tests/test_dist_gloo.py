"""world_size-2 gloo test of the data-parallel host logic (kan_odes_b200/dist.py): contiguous sharding of the
trajectories, all-reduce of the unnormalised loss / gradient sums, one global normalisation.  The per-shard compute
is the CPU oracle here (test-only); on the GPU box the same functions wrap kanode_loss_grad_dev (bench.py)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

from kan_odes_b200.dist import combine_loss_grad, shard_bounds  # noqa: E402


def test_shard_bounds_cover_batch_exactly():
    for batch in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import glorot_params, lv_chain, lv_targets
    from oracle import Oracle
    chain = lv_chain(); p = glorot_params(chain, 0)
    sa = np.arange(35) * 0.1
    u0 = np.random.default_rng(1234).uniform(0.5, 2.0, (batch, 2))
    tg = lv_targets(u0, sa)
    lo, hi = shard_bounds(batch, rank, world)
    r = Oracle(chain.desc()).loss_grad(p, u0[lo:hi], (0.0, 3.5), sa, tg[lo:hi])
    nloc = hi - lo
    loss_sum = torch.tensor([r["loss"] * nloc * sa.size * 2], dtype=torch.float64)     # unnormalised, as the C ABI returns
    grad_sum = torch.tensor(r["grad"] * nloc, dtype=torch.float64)
    loss2, grad2, cnt2 = combine_loss_grad(loss_sum.clone(), grad_sum.clone(), nloc, sa.size, 2, sync=False)   # device-only variant
    loss, grad, total = combine_loss_grad(loss_sum, grad_sum, nloc, sa.size, 2)
    assert torch.equal(loss2, loss) and torch.equal(grad2, grad) and int(cnt2.item()) == total
    if rank == 0:
        q.put((float(loss), grad.numpy(), total))
    dist.barrier()
    dist.destroy_process_group()


class _StubLib:
    """Stands in for libkanode_b200.so in the collective decision of peer_setup (no GPU here): rank `bad` fails at `stage`."""
    def __init__(self, rank, bad, stage):
        self.rank, self.bad, self.stage, self.attached = rank, bad, stage, False

    def kanode_peer_export(self, h, buf):
        buf[0] = 1 + self.rank
        return -5 if (self.rank == self.bad and self.stage == "export") else 0

    def kanode_peer_attach(self, h, rank, world, blob):
        assert bytes(blob)[0] == 1 and bytes(blob)[64] == 2      # handles arrive in rank order
        if self.rank == self.bad and self.stage == "attach":
            return -5
        self.attached = True
        return 0


class _StubOde:
    def __init__(self, rank, bad, stage, np_=240):
        self.lib, self.h, self.np_ = _StubLib(rank, bad, stage), None, np_


def _peer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kan_odes_b200.dist import peer_setup
    res = []
    for bad, stage, np_ in ((-1, "", 240), (1, "export", 240), (0, "attach", 240), (-1, "", 5000)):
        ode = _StubOde(rank, bad, stage, np_)
        res.append((peer_setup(ode), ode._peer_ok))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_setup_is_all_or_none_and_never_leaves_a_rank_waiting():
    """peer_setup's collective decision (kan_odes_b200/dist.py): a rank whose GPU cannot export or map a mailbox, or a model too
    large for it, sends EVERY rank to the NCCL path — and every rank still takes part in every gather (no hang)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 23
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = [(True, True), (False, False), (False, False), (False, False)]
    assert got[0] == want and got[1] == want


@pytest.mark.parametrize("batch", [8, 7])      # even and ragged shards
def test_two_rank_gloo_matches_single_process(batch):
    from conftest import glorot_params, lv_chain, lv_targets
    from oracle import Oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + batch
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    [p.start() for p in procs]
    loss, grad, total = q.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs) and total == batch
    chain = lv_chain(); p = glorot_params(chain, 0)
    sa = np.arange(35) * 0.1
    u0 = np.random.default_rng(1234).uniform(0.5, 2.0, (batch, 2))
    ref = Oracle(chain.desc()).loss_grad(p, u0, (0.0, 3.5), sa, lv_targets(u0, sa))
    assert abs(loss - ref["loss"]) < 1e-12 * abs(ref["loss"])
    assert np.abs(grad - ref["grad"]).max() < 1e-12 * np.abs(ref["grad"]).max()
