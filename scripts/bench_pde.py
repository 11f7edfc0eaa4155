"""Throughput of the PDE-surrogate configs of BASELINE.json (configs[2..4]) through the wide lockstep engine, on N GPUs.

Not the driver's bench contract (that is bench.py on configs[1]); same measurement rules: W >= 3 warm-up steps, K timed
steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks, 256 MiB L2 flush between
timed steps, inputs resident in HBM (device-pointer C-ABI call), weak scaling (`--batch` ICs per GPU), and the ONLY
collective is the all-reduce of the gradient / loss sums.  One JSON line per config (rank 0):
  value        initial conditions trained per second (fwd Tsit5 + interpolating adjoint + gradient), whole job
  roofline     the step-end pass over the per-IC gradient state g (wide_gp1/wide_gp2, HBM-bound): algorithmic bytes =
               2 * np * sizeof(T) per IC per step attempt (read g_old, write g_new), time from CUDA events inside the library
  kernel_ms    forward / backward / gradient reduction (CUDA events inside the library)

  python scripts/bench_pde.py [burgers1024 ac4096 schrodinger16384 source4096] [--batch B] [--dtype f32|f64] [--steps K]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_pde.py schrodinger16384 --batch 32
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import kan_odes_b200 as K  # noqa: E402
from kan_odes_b200 import abi  # noqa: E402


def surrogate(n, G):
    return K.Chain(K.KDense(n, 10, G, normalizer=K.softsign), K.KDense(10, n, G, normalizer=K.softsign))


def make(name, batch, rng):
    if name == "burgers1024":                                   # configs[2]; Burgers_Surrogate.jl:43,68,82-88
        n = 1024; chain = surrogate(n, 5); x = np.linspace(-1, 1, n)
        u0 = -rng.uniform(0.5, 1.5, (batch, 1)) * np.sin(np.pi * x)[None, :]
        sa = np.array([0.0, 0.1, 0.3, 0.5, 0.7, 0.9]); ts = (0.0, 1.0); kw = {}
        tg = u0[:, None, :] * np.exp(-sa)[None, :, None]
    elif name == "ac4096":                                      # configs[3] wide layer; Allen-Cahn_Surrogate.jl:80-87
        n = 4096; chain = surrogate(n, 10); x = np.linspace(-1, 1, n)
        u0 = rng.uniform(0.8, 1.2, (batch, 1)) * (x**2 * np.cos(np.pi * x))[None, :]
        sa = np.array([0.1, 0.3, 0.5, 0.7, 0.9]); ts = (0.0, 1.0); kw = {}
        tg = u0[:, None, :] * (1 - 0.5 * sa)[None, :, None]
    elif name == "schrodinger16384":                            # configs[4]; Schrodinger_Surrogate.jl:68,73,89-96
        n = 32768; chain = surrogate(n, 10); x = np.linspace(-5, 5, 16384)
        u0 = rng.uniform(0.8, 1.2, (batch, 1)) * np.concatenate([2 / np.cosh(x), np.zeros_like(x)])[None, :]
        sa = np.array([0.1, 0.3, 0.5, 0.7, 0.9, 1.1, 1.3, 1.5]); ts = (0.0, np.pi / 2); kw = {}
        tg = u0[:, None, :] * np.cos(sa)[None, :, None]
    elif name == "source4096":                                  # configs[3] hidden source; Allen-Cahn_Source.jl:34-54,76-99
        n = 4096; chain = K.Chain(K.KDense(1, 1, 10, normalizer=K.softsign)); x = np.linspace(-1, 1, n)
        u0 = rng.uniform(0.8, 1.2, (batch, 1)) * (x**2 * np.cos(np.pi * x))[None, :]
        sa = np.linspace(0, 0.2, 21); ts = (0.0, 0.2)           # stable Fisher-KPP sign at this resolution (tests/test_gpu_pde.py)
        kw = dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=n, lap_coef=1e-4, dx=2.0 / (n - 1))
        tg = u0[:, None, :] * np.exp(0.5 * sa)[None, :, None]
    else:
        raise SystemExit(f"unknown config {name}")
    ps, _ = K.setup(np.random.default_rng(0), chain)
    return chain, kw, K.flatten_params(ps), u0, ts, sa, tg


def main():
    import torch
    import torch.distributed as dist

    from kan_odes_b200.dist import combine_loss_grad
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["burgers1024"])
    ap.add_argument("--batch", type=int, default=64, help="initial conditions per GPU")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    f64 = a.dtype == "f64"
    tdt, ndt, creal, esz = (torch.float64, np.float64, C.c_double, 8) if f64 else (torch.float32, np.float32, C.c_float, 4)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    for name in a.configs:
        B = a.batch
        chain, kw, p, u0, ts, sa, tg = make(name, B, np.random.default_rng(3 + rank))      # every rank its own shard
        stream = torch.cuda.Stream()
        ode = K.KanOde(chain, kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0),
                       device=local, stream=stream.cuda_stream, dtype=ndt)
        ode.set_params(p)
        lib = ode.lib
        lib.kanode_set_record_capacity(ode.h, 512 if name == "source4096" else 64)   # *_dev entry points do not regrow the dense record
        with torch.cuda.stream(stream):
            d_u0 = torch.tensor(u0, dtype=tdt, device=dev); d_tg = torch.tensor(tg, dtype=tdt, device=dev)
            d_grad = torch.zeros(ode.np_, dtype=tdt, device=dev); d_loss = torch.zeros(1, dtype=torch.float64, device=dev)
            d_fst = torch.zeros(B * 4, dtype=torch.int32, device=dev); d_bst = torch.zeros(B * 4, dtype=torch.int32, device=dev)
            flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        sac = np.ascontiguousarray(sa, dtype=np.float64)
        fn = lib.kanode_loss_grad_dev_f64 if f64 else lib.kanode_loss_grad_dev
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                       creal, creal, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

        last = [None]

        def step():
            rc = fn(ode.h, d_u0.data_ptr(), B, ts[0], ts[1], sac.ctypes.data, sac.size, d_tg.data_ptr(), 1e-6, 1e-3,
                    d_loss.data_ptr(), d_grad.data_ptr(), None, d_fst.data_ptr(), d_bst.data_ptr())
            abi.check(lib, ode.h, rc, "kanode_loss_grad_dev")
            if world > 1:
                last[0] = combine_loss_grad(d_loss, d_grad, B, sa.size, ode.n, sync=False)

        m3 = (C.c_float * 3)(); gms = C.c_float(); gpasses = C.c_int32()
        with torch.cuda.stream(stream):
            for _ in range(max(a.warmup, 3)):
                step()
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            launches0 = ode.launch_count()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
            kms = np.zeros((a.steps, 3)); gp = np.zeros(a.steps); gpn = np.zeros(a.steps); wide = True
            for i in range(a.steps):
                flush.zero_()
                evs[i][0].record(stream); step(); evs[i][1].record(stream)
                lib.kanode_last_timing(ode.h, m3); kms[i] = list(m3)
                if lib.kanode_last_gpass_timing(ode.h, C.byref(gms), C.byref(gpasses)) == 0:
                    gp[i], gpn[i] = gms.value, gpasses.value
                else:
                    wide = False
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            launches = ode.launch_count() - launches0
            total_ms = torch.tensor([sum(x.elapsed_time(y) for x, y in evs)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
            total_ms = float(total_ms.item())
            fst = d_fst.cpu().numpy().reshape(B, 4); bst = d_bst.cpu().numpy().reshape(B, 4)
            loss = float(last[0][0].item()) if world > 1 else float(d_loss.item()) / (B * sa.size * ode.n)
        if rank == 0:
            k = kms.mean(0)
            attempts = int((bst[:, 0] + bst[:, 1]).sum())              # per-IC step attempts of the adjoint = g passes per IC
            line = {"metric": "kan_ode_fwd_adjoint_ic_train_steps_per_s", "config": {"workload": name, "batch_per_gpu": B, "global_batch": world * B,
                                                                                     "n": ode.n, "np": ode.np_, "l2": "256 MiB flush between timed steps",
                                                                                     "parallelism": f"dp{world} (ICs sharded, gradient all-reduce only)"},
                    "value": world * B * a.steps / (total_ms / 1e3), "unit": "ICs/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                    "ms_per_step": total_ms / a.steps, "scaling": "weak", "dtype": a.dtype, "data": "synthetic",
                    "kernel_ms": {"forward": k[0], "backward": k[1], "grad_reduce": k[2]},
                    "rhs_evals_per_s": world * int(fst[:, 2].sum() + bst[:, 2].sum()) * a.steps / (total_ms / 1e3),
                    "fwd_steps": [int(fst[:, 0].min()), int(fst[:, 0].max())], "bwd_steps": [int(bst[:, 0].min()), int(bst[:, 0].max())],
                    "failed": int((fst[:, 3] != 0).sum() + (bst[:, 3] != 0).sum()), "loss": loss, "gpu_launches": int(launches)}
            if wide and gp.mean() > 0:
                alg = attempts * 2 * ode.np_ * esz
                traffic = None                                  # DRAM bytes of one gp1+gp2 launch pair from the committed ncu capture
                try:
                    tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
                    if name == "schrodinger16384" and B == 32 and not f64:
                        traffic = tj["wide_gp1_kernel_bytes_schrodinger16384_b32"] + tj["wide_gp2_kernel_bytes_schrodinger16384_b32"]
                except Exception:
                    pass
                ach = alg / (gp.mean() / 1e3) / 1e9
                line["roofline"] = {"kernel": "wide_gp1_kernel+wide_gp2_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": ach / hbm_peak, "algorithmic_bytes": alg, "ms": float(gp.mean()), "passes": int(gpn.mean()),
                                    "share_of_backward": float(gp.mean() / k[1]), "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                                    "traffic": traffic}
            print(json.dumps(line), flush=True)
        ode.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
