"""Throughput of the PDE-surrogate configs of BASELINE.json (configs[2..4]) through the wide lockstep engine, on N GPUs.

Thin command line over bench.run_pde_workload (the same code bench.py folds into its JSON line under "workloads"):
W >= 3 warm-up steps, K timed steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks,
256 MiB L2 flush between timed steps, SM clocks / throttle reasons sampled during the timed region, inputs resident in HBM
(device-pointer C-ABI call), weak scaling (`--batch` ICs per GPU); the ONLY collective is the all-reduce of the gradient /
loss sums.  One JSON line per workload (rank 0).

  python scripts/bench_pde.py [burgers1024 ac4096 schrodinger16384 source4096] [--batch B] [--dtype f32|f64] [--steps K]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_pde.py schrodinger16384 --batch 32
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["burgers1024"])
    ap.add_argument("--batch", type=int, default=0, help="initial conditions per GPU (default: the bench.py figure of the workload)")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu", action="store_true", help="add the CPU-oracle baseline (N = 1)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for name in a.workloads:
        line = bench.run_pde_workload(name, a.batch or bench.PDE_WORKLOADS[name], a.dtype, a.steps, a.warmup, world, rank, local,
                                      bench.read_peaks(), with_cpu=a.cpu)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
