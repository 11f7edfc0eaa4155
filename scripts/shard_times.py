"""Diagnostic: device time of the LV backward kernel for the 8 synthetic shards the 8-GPU weak-scaling bench uses (seed 1234 + rank):
the step time of the job is the maximum over ranks, so the shard with the longest adaptive solve sets it.
usage: python scripts/shard_times.py  (needs a GPU)"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import kan_odes_b200 as K  # noqa: E402

for r in range(8):
    chain, p, u0, tg = bench.make_workload(65536, 1234 + r)
    ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
    for _ in range(3):
        res = ode.loss_grad(u0, bench.TSPAN, bench.SAVEAT, tg)
    m3 = (C.c_float * 3)(); ode.lib.kanode_last_timing(ode.h, m3)
    att = res["bwd_stats"].naccept + res["bwd_stats"].nreject
    print(f"shard {r}: fwd {m3[0]:.3f} ms  bwd {m3[1]:.3f} ms  reduce {m3[2]:.3f} ms   backward attempts mean {att.mean():.1f} max {att.max()} "
          f"(> 100: {(att > 100).sum()})", flush=True)
    ode.close()
