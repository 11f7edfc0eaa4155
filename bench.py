#!/usr/bin/env python
"""bench.py — KAN-ODE fwd+adjoint training-step throughput (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[1]): Lotka-Volterra KAN-ODE [2,10,2] grid=5 (tanh_fast, RBF, SiLU base branch),
ensemble of 65,536 synthetic initial conditions PER GPU (weak scaling), u0 ~ U[0.5,2]^2 seed 1234, tspan (0,3.5),
saveat 0:0.1:3.4, targets = true Lotka-Volterra (1.5,1,1,3) trajectories, abstol=1e-6, reltol=1e-3 (the reference's
defaults), parameters = glorot_uniform seed 0 (non-trivial field).  One "step" = one kanode_loss_grad over the whole
ensemble: dense forward Tsit5 solve + interpolating-adjoint backward solve + gradient reduction (+ NCCL all-reduce
of the 240-float gradient when N>1).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dtype f32|f64]

Prints ONE JSON line (rank 0).  `value` = trajectories trained per second with inputs resident in HBM (CUDA events on
the launching stream, max over ranks); `e2e` = the same through the host-pointer C-ABI call (pinned host buffers,
H2D/D2H inside the timed region); `roofline` describes the dominant kernel (the adjoint/backward kernel);
`cpu_baseline` is the CPU oracle (a C++ port of the reference algorithm, NOT Julia) on the box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TSPAN = (0.0, 3.5)
SAVEAT = np.arange(35) * 0.1                     # t_train (LV_driver_KANODE.jl:116,123-125)
FLOP_FWD_EVAL = 816                              # FLOP per KAN RHS evaluation per sample
FLOP_BWD_EVAL = 3 * FLOP_FWD_EVAL                # a fused forward+VJP evaluation counts 3x
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # derived fp32 FFMA peak at 1965 MHz (not in MEASURED_PEAKS.json)


def lv_chain():
    import kan_odes_b200 as K
    return K.Chain(K.KDense(2, 10, 5, use_base_act=True, basis_func=K.rbf, normalizer=K.tanh_fast),
                   K.KDense(10, 2, 5, use_base_act=True, basis_func=K.rbf, normalizer=K.tanh_fast))


def make_workload(batch: int, seed: int):
    """Synthetic ICs + true-LV targets (vectorised classical RK4, h=0.0025 => ~1e-11 accurate)."""
    import kan_odes_b200 as K
    chain = lv_chain()
    ps, _ = K.setup(np.random.default_rng(0), chain)
    p = K.flatten_params(ps)
    u0 = np.random.default_rng(seed).uniform(0.5, 2.0, (batch, 2))

    def f(u):
        return np.stack([1.5 * u[:, 0] - u[:, 1] * u[:, 0], u[:, 0] * u[:, 1] - 3.0 * u[:, 1]], axis=1)
    tg = np.empty((batch, SAVEAT.size, 2))
    u, h, sub = u0.copy(), 0.0025, 40
    for s in range(SAVEAT.size):
        tg[:, s] = u
        for _ in range(sub):
            k1 = f(u); k2 = f(u + 0.5 * h * k1); k3 = f(u + 0.5 * h * k2); k4 = f(u + h * k3)
            u = u + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return chain, p, u0, tg


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, n, 0): lab for n, lab in [
            ("nvmlClocksEventReasonHwSlowdown", "hw_slowdown"),
            ("nvmlClocksEventReasonHwThermalSlowdown", "hw_thermal_slowdown"),
            ("nvmlClocksEventReasonSwThermalSlowdown", "sw_thermal_slowdown"),
            ("nvmlClocksEventReasonSwPowerCap", "sw_power_cap"),
            ("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"),
            ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
            ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"),
            ("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap")] if getattr(nv, n, 0)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, lab in names.items():
                    if r & bit:
                        self.reasons.add(lab)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(chain, p, u0, tg, budget_s: float = 12.0):
    """CPU oracle (C++ port of the reference algorithm, fp64, OpenMP over trajectories) on a bounded sample."""
    from oracle import Oracle
    orc = Oracle(chain.desc(), np.float64)
    cores = os.cpu_count() or 1
    n0 = min(1024, u0.shape[0])
    t = time.perf_counter(); orc.loss_grad(p, u0[:n0], TSPAN, SAVEAT, tg[:n0]); dt0 = time.perf_counter() - t
    n = int(min(u0.shape[0], max(n0, n0 * budget_s / max(dt0, 1e-3))))
    t = time.perf_counter(); r = orc.loss_grad(p, u0[:n], TSPAN, SAVEAT, tg[:n]); dt = time.perf_counter() - t
    nf = int(r["fwd_stats"][:, 2].sum() + r["bwd_stats"][:, 2].sum())
    return {"value": n / dt, "unit": "trajectories/s", "cores": cores, "kind": "port",
            "sample": f"{n} of the workload's trajectories, one fwd+adjoint step, fp64 C++ oracle with OpenMP "
                      f"({cores} threads); not the Julia reference (Julia absent)",
            "rhs_evals_per_s": nf / dt}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chain, p, u0, tg = make_workload(8192, 1234)
    from oracle import Oracle
    orc = Oracle(chain.desc(), np.float64)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        orc.loss_grad(p, u0[:1024], TSPAN, SAVEAT, tg[:1024])
    t = time.perf_counter()
    for _ in range(args.steps):
        orc.loss_grad(p, u0, TSPAN, SAVEAT, tg)
    dt = time.perf_counter() - t
    val = u0.shape[0] * args.steps / dt
    line = {"impl": "reference", "metric": "kan_ode_fwd_adjoint_trajectory_train_steps_per_s", "value": val,
            "unit": "trajectories/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "lotka_volterra_kan_ode_2_10_2_g5_ensemble", "batch_per_step": int(u0.shape[0]),
                       "note": "each step is a bounded 8192-trajectory sample of the 65,536-trajectory workload"},
            "cpu_baseline": {"value": val, "unit": "trajectories/s", "cores": cores, "kind": "port",
                             "sample": "8192 trajectories per step, fp64 C++ oracle (OpenMP); Julia absent"},
            "e2e": {"value": val, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import kan_odes_b200 as K
    from kan_odes_b200 import abi
    from kan_odes_b200.dist import combine_loss_grad

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the KAN-ODE hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    f64 = args.dtype == "f64"
    tdt, ndt, creal = (torch.float64, np.float64, C.c_double) if f64 else (torch.float32, np.float32, C.c_float)

    B = args.batch
    chain, p, u0, tg = make_workload(B, 1234 + rank)              # every rank trains its own shard (weak scaling)
    lib = abi.load_library()
    stream = torch.cuda.Stream()
    ode = K.KanOde(chain, device=local, stream=stream.cuda_stream, dtype=ndt)
    ode.set_params(p)
    npar = ode.np_
    dev = torch.device("cuda", local)
    with torch.cuda.stream(stream):
        d_u0 = torch.tensor(u0, dtype=tdt, device=dev)
        d_tg = torch.tensor(tg, dtype=tdt, device=dev)
        d_grad = torch.zeros(npar, dtype=tdt, device=dev)
        d_loss = torch.zeros(1, dtype=torch.float64, device=dev)
        d_fst = torch.zeros(B * 4, dtype=torch.int32, device=dev)
        d_bst = torch.zeros(B * 4, dtype=torch.int32, device=dev)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    sa = np.ascontiguousarray(SAVEAT)
    fn = lib.kanode_loss_grad_dev_f64 if f64 else lib.kanode_loss_grad_dev
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                   creal, creal, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    def step(with_stats: bool):
        rc = fn(ode.h, d_u0.data_ptr(), B, TSPAN[0], TSPAN[1], sa.ctypes.data, sa.size, d_tg.data_ptr(), 1e-6, 1e-3,
                d_loss.data_ptr(), d_grad.data_ptr(), None,
                d_fst.data_ptr() if with_stats else None, d_bst.data_ptr() if with_stats else None)
        abi.check(lib, ode.h, rc, "kanode_loss_grad_dev")
        if world > 1 and not os.environ.get("KANODE_BENCH_NO_ALLREDUCE"):  # the only collective: gradient + loss sums
            combine_loss_grad(d_loss, d_grad, B, SAVEAT.size, 2, sync=False)   # no host read-back inside the timed loop
            # (KANODE_BENCH_NO_ALLREDUCE=1 is a diagnostic: it isolates the collective's share of the step at N > 1; not a bench mode)

    ms3 = (C.c_float * 3)()
    with torch.cuda.stream(stream):
        step(True)                                                 # untimed: per-trajectory statistics
        stream.synchronize()
        fst = d_fst.cpu().numpy().reshape(B, 4); bst = d_bst.cpu().numpy().reshape(B, 4)
        nf_f, nf_b = int(fst[:, 2].sum()), int(bst[:, 2].sum())
        failed = int((fst[:, 3] != 0).sum() + (bst[:, 3] != 0).sum())
        for _ in range(max(args.warmup, 3)):
            step(False)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local); sampler.start()
        launches0 = ode.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        k_ms = np.zeros((args.steps, 3))
        for i in range(args.steps):
            flush.zero_()                                          # L2 flush between timed iterations (untimed)
            evs[i][0].record(stream)
            step(False)
            evs[i][1].record(stream)
            if world == 1:
                lib.kanode_last_timing(ode.h, ms3)                 # per-kernel CUDA-event times of this step (blocks until it finished)
                k_ms[i] = [ms3[0], ms3[1], ms3[2]]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler.stop_flag = True; sampler.join()
        launches = ode.launch_count() - launches0
        if world > 1:
            # N > 1: the timed loop above never blocks the host (8 ranks share the box's cores; a per-step read-back exposes the
            # launch latency of every rank to the all-reduce).  Per-kernel times come from extra, untimed steps.
            for i in range(args.steps):
                step(False)
                lib.kanode_last_timing(ode.h, ms3)
                k_ms[i] = [ms3[0], ms3[1], ms3[2]]
            torch.cuda.synchronize()
            dist.barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * B * args.steps / (total_ms / 1e3)

    # ---- end to end through the host-pointer C-ABI call (pinned host buffers) ----
    h_u0 = torch.tensor(u0, dtype=tdt).pin_memory().numpy()
    h_tg = torch.tensor(tg, dtype=tdt).pin_memory().numpy()
    h_p = np.ascontiguousarray(p, dtype=ndt)
    e2e_steps = max(2, min(args.steps, 5))
    # the reference's call is Zygote.gradient(loss, p): loss and gradient only (no d loss/d u0, no solver statistics)
    for _ in range(2):
        ode.set_params(h_p); r = ode.loss_grad(h_u0, TSPAN, SAVEAT, h_tg, want_du0=False, want_stats=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ode.set_params(h_p)
        r = ode.loss_grad(h_u0, TSPAN, SAVEAT, h_tg, want_du0=False, want_stats=False)
        if world > 1:
            g = torch.tensor(r["grad"], device=dev); dist.all_reduce(g); g.cpu()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * B * e2e_steps / float(te.item())
    esz = 8 if f64 else 4
    h2d = h_u0.nbytes + h_tg.nbytes + SAVEAT.nbytes + npar * esz
    d2h = 8 + npar * esz + B * 16                                  # loss sum, gradient, forward retcodes (dense-record overflow check)

    # per-rank device times and SM clocks: the job's step time is the max over ranks, so a slower GPU (clock / power) shows here
    per_rank = None
    if world > 1:
        mine = sampler.result()
        loc = torch.tensor([float(k_ms[:, 0].mean()), float(k_ms[:, 1].mean()), float(k_ms[:, 2].mean()),
                            float(mine["sm_mhz"] or 0.0)], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(loc) for _ in range(world)]
        dist.all_gather(allv, loc)
        per_rank = {"forward_ms": [round(float(v[0]), 4) for v in allv], "backward_ms": [round(float(v[1]), 4) for v in allv],
                    "sm_mhz": [float(v[3]) for v in allv]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = sampler.result()
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bwd_ms = float(k_ms[:, 1].mean()); fwd_ms = float(k_ms[:, 0].mean()); red_ms = float(k_ms[:, 2].mean())
    bwd_flop = nf_b * FLOP_BWD_EVAL
    achieved_tf = bwd_flop / (bwd_ms / 1e3) / 1e12
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get("small_backward_kernel_bytes")
    except Exception:
        pass
    # algorithmic HBM bytes of the backward kernel: dense record read once, dL/du read once,
    # final per-trajectory gradient written once
    rec_bytes = int(fst[:, 0].sum()) * (8 + 17 * esz)
    alg_bytes = rec_bytes + B * SAVEAT.size * 2 * esz + B * npar * esz
    line = {
        "metric": "kan_ode_fwd_adjoint_trajectory_train_steps_per_s", "value": value, "unit": "trajectories/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "lotka_volterra_kan_ode_2_10_2_g5_ensemble", "batch_per_gpu": B,
                   "global_batch": world * B, "tspan": list(TSPAN), "nsave": int(SAVEAT.size), "abstol": 1e-6,
                   "reltol": 1e-3, "params": "glorot_uniform seed 0", "l2": "256 MiB flush between timed steps",
                   "parallelism": f"dp{world} (trajectories sharded, gradient all-reduce only)"},
        "train_steps_per_s": args.steps / (total_ms / 1e3),
        "rhs_evals_per_s": world * (nf_f + nf_b) * args.steps / (total_ms / 1e3),
        "rhs_evals_per_step_per_gpu": {"forward": nf_f, "backward_fused_fwd_vjp": nf_b},
        "failed_trajectories": failed,
        "kernel_ms": {"forward": fwd_ms, "backward": bwd_ms, "grad_reduce": red_ms},
        "roofline": {"kernel": "small_backward_kernel", "bound": "ffma", "achieved": achieved_tf,
                     "peak": FFMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved_tf / FFMA_PEAK_TFLOPS,
                     "peak_source": "derived 148 SM x 128 lanes x 2 x 1.965 GHz (fp32 FFMA; not in MEASURED_PEAKS.json)",
                     "flop_per_unit": FLOP_BWD_EVAL, "units_per_launch": nf_b, "traffic": traffic,
                     "hbm": {"algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (bwd_ms / 1e3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (bwd_ms / 1e3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "e2e": {"value": e2e_val, "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if per_rank is not None:
        line["per_rank"] = per_rank
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(chain, p, u0, tg)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was sent to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)          # keep stdout for the JSON line only
    os.dup2(2, 1)                     # native-library prints to fd 1 (e.g. "NCCL version ...") go to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="trajectories per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
