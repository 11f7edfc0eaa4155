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
H2D/D2H inside the timed region); `roofline` describes the dominant kernel (the adjoint/backward kernel) against the
FFMA, MUFU and HBM peaks; `cpu_baseline` is the CPU oracle (a C++ port of the reference algorithm, NOT Julia) on the
box's host cores.  Beside the headline (N = 1 only, skipped with --lean):
  `parity`     the timed dtype against the fp64 oracle on a sample of the same workload: adaptive run (states, gradient,
               % identical accepted-step counts) and dt-replay run (arithmetic parity on the oracle's step sequence);
  `f64`        the parity-exact fp64 instantiation timed the same way (`value_f64`, `ms_per_step_f64`);
  `cfg1`       BASELINE configs[0]: latency of one single-trajectory fwd+adjoint call (GPU, and the oracle on one core);
  `workloads`  BASELINE configs[2..4] (Burgers-1024, Allen-Cahn-4096 wide layer, hidden source 4096, Schrodinger-16384)
               through the wide lockstep engine, each with value, ms_per_step, roofline, cpu_baseline and its own clocks.
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
MUFU_PER_EVAL = 96                               # 72 ex2 + 24 rcp per RHS evaluation
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


def read_peaks():
    """HBM peak from MEASURED_PEAKS.json (driver-written); FFMA / MUFU peaks from the committed microbenchmark
    (scripts/peaks_ffma_mufu.cu -> profiles/peaks_ffma_mufu.json), else the derived figures."""
    peaks = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)", "ffma_tflops": FFMA_PEAK_TFLOPS,
             "ffma_source": "derived 148 SM x 128 lanes x 2 x 1.965 GHz", "mufu_tops": 148 * 16 * 1.965e9 / 1e12,
             "mufu_source": "derived 148 SM x 16 lanes x 1.965 GHz"}
    try:
        m = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        peaks["hbm_gbs"] = float(m["hbm_gbs"]); peaks["hbm_source"] = "MEASURED_PEAKS.json"
    except Exception:
        pass
    try:
        m = json.loads((ROOT / "profiles" / "peaks_ffma_mufu.json").read_text())
        peaks["ffma_tflops"] = float(max(m["ffma_tflops"], m.get("ffma2_tflops", 0.0)))
        peaks["mufu_tops"] = float(m["mufu_ex2_tops"])
        peaks["ffma_source"] = peaks["mufu_source"] = "measured on B200: profiles/peaks_ffma_mufu.json (scripts/peaks_ffma_mufu.cu)"
    except Exception:
        pass
    return peaks


def surrogate(n, G):
    import kan_odes_b200 as K
    return K.Chain(K.KDense(n, 10, G, normalizer=K.softsign), K.KDense(10, n, G, normalizer=K.softsign))


PDE_WORKLOADS = {"burgers1024": 64, "ac4096": 64, "source4096": 64, "schrodinger16384": 32}   # name -> ICs per GPU


def pde_make(name, batch, rng):
    """Synthetic inputs of BASELINE configs[2..4]: model, parameters (glorot seed 0), ICs, save times, targets."""
    import kan_odes_b200 as K
    from kan_odes_b200 import abi
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
        raise SystemExit(f"unknown workload {name}")
    ps, _ = K.setup(np.random.default_rng(0), chain)
    return chain, kw, K.flatten_params(ps), u0, ts, sa, tg


def pde_desc(chain, kw):
    from kan_odes_b200 import abi
    return chain.desc(kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0))


def pde_cpu_baseline(name, threads, budget_s=8.0):
    """fp64 oracle on a bounded sample of a PDE workload: ICs are added until ~budget_s of CPU time is spent."""
    from oracle import Oracle
    nb = max(1, min(threads, 8))
    chain, kw, p, u0, ts, sa, tg = pde_make(name, nb, np.random.default_rng(3))
    orc = Oracle(pde_desc(chain, kw), np.float64)
    orc.set_threads(threads)
    t = time.perf_counter(); done = 0
    while True:
        orc.loss_grad(p, u0, ts, sa, tg); done += nb
        dt = time.perf_counter() - t
        if dt > budget_s or done >= 8 * nb:
            break
    return {"value": done / dt, "unit": "ICs/s", "cores": threads, "kind": "port",
            "sample": f"{done} ICs of the workload in batches of {nb}, one fwd+adjoint step each, fp64 C++ oracle (OpenMP over ICs, {threads} threads); Julia absent"}


def run_pde_workload(name, B, dtype, steps, warmup, world, rank, local, peaks, with_cpu=True, seed=3):
    """One PDE workload through the device-pointer C-ABI call; returns the result dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    import kan_odes_b200 as K
    from kan_odes_b200 import abi
    from kan_odes_b200.dist import combine_loss_grad
    dev = torch.device("cuda", local)
    f64 = dtype == "f64"
    tdt, ndt, creal, esz = (torch.float64, np.float64, C.c_double, 8) if f64 else (torch.float32, np.float32, C.c_float, 4)
    chain, kw, p, u0, ts, sa, tg = pde_make(name, B, np.random.default_rng(seed + rank))      # every rank its own shard
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
        for _ in range(max(warmup, 3)):
            step()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local); sampler.start()
        launches0 = ode.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kms = np.zeros((steps, 3)); gp = np.zeros(steps); gpn = np.zeros(steps); wide = True
        align_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        for i in range(steps):
            flush.zero_()
            if world > 1:
                dist.all_reduce(align_buf)                         # untimed device-side rendezvous (see run_ours): launch / flush jitter of
                                                                   # one rank must not be waited for inside the others' timed step
            evs[i][0].record(stream); step(); evs[i][1].record(stream)
            lib.kanode_last_timing(ode.h, m3); kms[i] = list(m3)
            if lib.kanode_last_gpass_timing(ode.h, C.byref(gms), C.byref(gpasses)) == 0:
                gp[i], gpn[i] = gms.value, gpasses.value
            else:
                wide = False
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler.stop_flag = True; sampler.join()
        launches = ode.launch_count() - launches0
        total_ms = torch.tensor([sum(x.elapsed_time(y) for x, y in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_ms = float(total_ms.item())
        fst = d_fst.cpu().numpy().reshape(B, 4); bst = d_bst.cpu().numpy().reshape(B, 4)
        loss = float(last[0][0].item()) if world > 1 else float(d_loss.item()) / (B * sa.size * ode.n)
    # ---- end to end: the host-pointer call (pinned host buffers, H2D / D2H inside the timed region) ----
    h_u0 = torch.tensor(u0, dtype=tdt).pin_memory().numpy(); h_tg = torch.tensor(tg, dtype=tdt).pin_memory().numpy()
    h_p = np.ascontiguousarray(p, dtype=ndt)
    for _ in range(2):
        ode.set_params(h_p); ode.loss_grad(h_u0, ts, sa, h_tg, want_du0=False, want_stats=False)
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ode.set_params(h_p); r = ode.loss_grad(h_u0, ts, sa, h_tg, want_du0=False, want_stats=False)
        if world > 1:
            g = torch.tensor(r["grad"], device=dev); dist.all_reduce(g); g.cpu()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * B * e2e_steps / float(te.item())
    n_state, npar = ode.n, ode.np_
    ode.close()
    if rank != 0:
        return None
    k = kms.mean(0)
    attempts = int((bst[:, 0] + bst[:, 1]).sum())              # per-IC step attempts of the adjoint = g passes per IC
    line = {"metric": "kan_ode_fwd_adjoint_ic_train_steps_per_s",
            "config": {"workload": name, "batch_per_gpu": B, "global_batch": world * B, "n": n_state, "np": npar,
                       "l2": "256 MiB flush between timed steps" + (", then an untimed device-side rendezvous of the ranks" if world > 1 else ""),
                       "parallelism": f"dp{world} (ICs sharded, gradient all-reduce only)"},
            "value": world * B * steps / (total_ms / 1e3), "unit": "ICs/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": total_ms / steps, "scaling": "weak", "dtype": dtype, "data": "synthetic",
            "kernel_ms": {"forward": float(k[0]), "backward": float(k[1]), "grad_reduce": float(k[2])},
            "rhs_evals_per_s": world * int(fst[:, 2].sum() + bst[:, 2].sum()) * steps / (total_ms / 1e3),
            "fwd_steps": [int(fst[:, 0].min()), int(fst[:, 0].max())], "bwd_steps": [int(bst[:, 0].min()), int(bst[:, 0].max())],
            "failed": int((fst[:, 3] != 0).sum() + (bst[:, 3] != 0).sum()), "loss": loss, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "ICs/s", "h2d_bytes_per_step": int(h_u0.nbytes + h_tg.nbytes + sac.nbytes + npar * esz),
                    "d2h_bytes_per_step": int(8 + npar * esz + B * 16), "steps": e2e_steps},
            "clocks": sampler.result()}
    hbm_peak = peaks["hbm_gbs"]
    if name != "source4096":
        # surrogates: the step-end pass over the per-IC gradient state g is the dominant traffic: 2 * np * sizeof(T) bytes per IC
        # per adjoint step attempt (read g_old, write g_new).  Per-kernel CUDA events exist when the attempts are launched
        # directly (n > 8192); under CUDA-graph replay (n <= 8192) the time base is the whole adjoint solve.
        alg = attempts * 2 * npar * esz
        traffic = None                                  # DRAM bytes of one gp1+gp2 launch pair from the committed ncu capture
        try:
            tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
            if name == "schrodinger16384" and B == 32 and not f64:
                traffic = tj["wide_gp1_kernel_bytes_schrodinger16384_b32"] + tj["wide_gp2_kernel_bytes_schrodinger16384_b32"]
        except Exception:
            pass
        per_kernel = bool(wide and gp.mean() > 0)
        gp_ms, gp_n = (float(gp.mean()), int(gpn.mean())) if per_kernel else (0.0, 0)
        if not per_kernel and n_state > 8192:
            # the timed steps replay CUDA graphs (no events inside an attempt): time the g passes in a separate, untimed profiling
            # step with direct launches (KANODE_WIDE_GRAPH_MAXN=8192 keeps the event pairs around every pass)
            old_env = os.environ.get("KANODE_WIDE_GRAPH_MAXN"); os.environ["KANODE_WIDE_GRAPH_MAXN"] = "8192"
            try:
                prof = K.KanOde(chain, kw.get("rhs_kind", abi.RHS_CHAIN), kw.get("n_state"), kw.get("lap_coef", 0.0), kw.get("dx", 1.0),
                                device=local, stream=stream.cuda_stream, dtype=ndt)
            finally:
                if old_env is None:
                    os.environ.pop("KANODE_WIDE_GRAPH_MAXN", None)
                else:
                    os.environ["KANODE_WIDE_GRAPH_MAXN"] = old_env
            prof.set_params(p); lib.kanode_set_record_capacity(prof.h, 64)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    rc = fn(prof.h, d_u0.data_ptr(), B, ts[0], ts[1], sac.ctypes.data, sac.size, d_tg.data_ptr(), 1e-6, 1e-3,
                            d_loss.data_ptr(), d_grad.data_ptr(), None, d_fst.data_ptr(), d_bst.data_ptr())
                    abi.check(lib, prof.h, rc, "kanode_loss_grad_dev")
                if lib.kanode_last_gpass_timing(prof.h, C.byref(gms), C.byref(gpasses)) == 0:
                    gp_ms, gp_n, per_kernel = float(gms.value), int(gpasses.value), True
            prof.close()
        t_ms = gp_ms if per_kernel else float(k[1])
        ach = alg / (t_ms / 1e3) / 1e9
        line["roofline"] = {"kernel": "wide_gp1_kernel+wide_gp2_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "algorithmic_bytes": alg, "ms": t_ms,
                            "time_base": "g-pass kernels (CUDA events around every pass, direct-launch profiling step)" if per_kernel else
                                         "whole adjoint solve, all its kernels (CUDA-graph replay: no per-kernel events)",
                            "passes": gp_n if per_kernel else None,
                            "share_of_backward": float(gp_ms / k[1]) if per_kernel else None,
                            "peak_source": peaks["hbm_source"], "traffic": traffic,
                            "whole_step_frac": alg / (total_ms / steps / 1e3) / 1e9 / hbm_peak}
    else:
        # hidden-source model: elementwise stencil + pointwise KAN, 2 * n * sizeof(T) bytes per RHS evaluation per IC
        evals = int(fst[:, 2].sum() + bst[:, 2].sum())
        alg = evals * 2 * n_state * esz
        ach = alg / (total_ms / steps / 1e3) / 1e9
        line["roofline"] = {"kernel": "wsrc stage kernels (whole step)", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "algorithmic_bytes": alg, "peak_source": peaks["hbm_source"], "traffic": None,
                            "note": "launch-latency regime: ~100 step attempts of 7 elementwise stage kernels over 1 MB of state"}
    if with_cpu and world == 1:
        line["cpu_baseline"] = pde_cpu_baseline(name, os.cpu_count() or 1)
    return line


def cpu_baseline(chain, p, u0, tg, budget_s: float = 12.0):
    """CPU oracle (C++ port of the reference algorithm, fp64, OpenMP over trajectories) on a bounded sample."""
    from oracle import Oracle
    orc = Oracle(chain.desc(), np.float64)
    cores = orc.set_threads(os.cpu_count() or 1)
    n0 = min(1024, u0.shape[0])
    t = time.perf_counter(); orc.loss_grad(p, u0[:n0], TSPAN, SAVEAT, tg[:n0]); dt0 = time.perf_counter() - t
    n = int(min(u0.shape[0], max(n0, n0 * budget_s / max(dt0, 1e-3))))
    t = time.perf_counter(); r = orc.loss_grad(p, u0[:n], TSPAN, SAVEAT, tg[:n]); dt = time.perf_counter() - t
    nf = int(r["fwd_stats"][:, 2].sum() + r["bwd_stats"][:, 2].sum())
    return {"value": n / dt, "unit": "trajectories/s", "cores": cores, "kind": "port",
            "sample": f"{n} of the workload's trajectories, one fwd+adjoint step, fp64 C++ oracle with OpenMP "
                      f"({cores} threads); not the Julia reference (Julia absent)",
            "rhs_evals_per_s": nf / dt}


def cfg1_inputs():
    """BASELINE configs[0]: the reference driver's single trajectory (LV_driver_KANODE.jl:111-127): u0 = (1,1), tspan (0,3.5),
    35 save times, target = true Lotka-Volterra; parameters = glorot seed 0 (non-trivial field)."""
    chain, p, _, _ = make_workload(1, 0)
    u0 = np.array([[1.0, 1.0]])

    def f(u):
        return np.stack([1.5 * u[:, 0] - u[:, 1] * u[:, 0], u[:, 0] * u[:, 1] - 3.0 * u[:, 1]], axis=1)
    tg = np.empty((1, SAVEAT.size, 2)); u, h = u0.copy(), 0.0025
    for s_ in range(SAVEAT.size):
        tg[:, s_] = u
        for _ in range(40):
            k1 = f(u); k2 = f(u + 0.5 * h * k1); k3 = f(u + 0.5 * h * k2); k4 = f(u + h * k3)
            u = u + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return chain, p, u0, tg


def cfg1_cpu_latency(reps: int = 200):
    """One fwd+adjoint call of the single-trajectory config on ONE host core (the reference driver is serial)."""
    from oracle import Oracle
    chain, p, u0, tg = cfg1_inputs()
    orc = Oracle(chain.desc(), np.float64)
    prev = orc.set_threads(0); orc.set_threads(1)
    orc.loss_grad(p, u0, TSPAN, SAVEAT, tg)
    t = time.perf_counter()
    for _ in range(reps):
        orc.loss_grad(p, u0, TSPAN, SAVEAT, tg)
    dt = (time.perf_counter() - t) / reps
    orc.set_threads(prev)
    return {"ms_per_call": 1e3 * dt, "cores": 1, "kind": "port", "sample": f"mean of {reps} calls, fp64 C++ oracle, 1 thread"}


def cfg1_gpu_latency(local: int, reps: int = 50):
    """Latency of one host-pointer kanode_loss_grad call for the single trajectory (replicas only: nothing to shard)."""
    import kan_odes_b200 as K
    chain, p, u0, tg = cfg1_inputs()
    out = {"workload": "lotka_volterra_kan_ode_2_10_2_g5_single_trajectory", "note": "latency-bound: absolute time only (SURVEY.md 8d)"}
    for name, dt_ in (("f64", np.float64), ("f32", np.float32)):
        ode = K.KanOde(chain, device=local, dtype=dt_); ode.set_params(p)
        for _ in range(5):
            ode.loss_grad(u0, TSPAN, SAVEAT, tg, want_du0=False, want_stats=False)
        ts_ = []
        for _ in range(reps):
            t = time.perf_counter(); ode.loss_grad(u0, TSPAN, SAVEAT, tg, want_du0=False, want_stats=False)
            ts_.append(time.perf_counter() - t)
        out[f"ms_per_call_{name}"] = 1e3 * float(np.median(ts_))
        ode.close()
    return out


def lv_parity(chain, p, u0, tg, ndt, local: int, n: int = 4096, step_cap: int = 128):
    """The timed dtype against the fp64 oracle on the first n trajectories of the bench workload ('accepted-step
    count and max rel. error vs oracle reported beside' the throughput).  adaptive = the run that is timed; replay = the same
    kernels on the oracle's accepted-step sequence (arithmetic parity, controller taken out)."""
    import kan_odes_b200 as K
    from oracle import Oracle
    n = min(n, u0.shape[0])
    orc = Oracle(chain.desc(), np.float64); orc.set_threads(os.cpu_count() or 1)
    ref = orc.loss_grad(p, u0[:n], TSPAN, SAVEAT, tg[:n], want_out=True, step_cap=step_cap)

    def rel(a, b_):
        return float(np.abs(np.asarray(a, np.float64) - b_).max() / np.abs(b_).max())
    ode = K.KanOde(chain, device=local, dtype=ndt); ode.set_params(p)
    r = ode.loss_grad(u0[:n], TSPAN, SAVEAT, tg[:n])
    sol = ode.solve(u0[:n], TSPAN, SAVEAT)
    same_f = r["fwd_stats"].naccept == ref["fwd_stats"][:, 0]; same_b = r["bwd_stats"].naccept == ref["bwd_stats"][:, 0]
    out = {"sample": f"first {n} trajectories of the bench workload vs the fp64 C++ oracle", "dtype": np.dtype(ndt).name,
           "adaptive": {"states_max_rel_err": rel(sol.array, ref["out"]), "grad_max_rel_err": rel(r["grad"], ref["grad"]),
                        "loss_rel_err": abs(r["loss"] - ref["loss"]) / abs(ref["loss"]),
                        "identical_naccept_forward_pct": 100.0 * float(same_f.mean()),
                        "identical_naccept_adjoint_pct": 100.0 * float(same_b.mean()),
                        "naccept_forward": [int(r["fwd_stats"].naccept.min()), int(r["fwd_stats"].naccept.max())],
                        "naccept_adjoint": [int(r["bwd_stats"].naccept.min()), int(r["bwd_stats"].naccept.max())]}}
    if np.isnan(ref["bwd_t"][:, -1]).all():
        q = ode.loss_grad_replay(u0[:n], TSPAN, SAVEAT, tg[:n], ref["fwd_t"], ref["bwd_t"])
        out["replay"] = {"states_max_rel_err": rel(q["out"], ref["out"]), "grad_max_rel_err": rel(q["grad"], ref["grad"]),
                         "du0_max_rel_err": rel(q["du0"], ref["du0"]), "loss_rel_err": abs(q["loss"] - ref["loss"]) / abs(ref["loss"]),
                         "identical_naccept_pct": 100.0 * float(((q["fwd_stats"].naccept == ref["fwd_stats"][:, 0]) &
                                                                  (q["bwd_stats"].naccept == ref["bwd_stats"][:, 0])).mean())}
    ode.close()
    return out


def lv_device_leg(chain, p, u0, tg, f64: bool, local: int, steps: int, warmup: int, schedule: bool = True):
    """Device-timed LV ensemble steps in one dtype on one GPU (inputs resident, L2 flushed between steps).
    schedule=False: the adjoint launch order does not use the previous call's step counts (KANODE_SCHEDULE=0)."""
    import torch

    import kan_odes_b200 as K
    from kan_odes_b200 import abi
    tdt, ndt, creal = (torch.float64, np.float64, C.c_double) if f64 else (torch.float32, np.float32, C.c_float)
    B = u0.shape[0]
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream()
    old = os.environ.get("KANODE_SCHEDULE")
    if not schedule:
        os.environ["KANODE_SCHEDULE"] = "0"                        # read once by kanode_create
    try:
        ode = K.KanOde(chain, device=local, stream=stream.cuda_stream, dtype=ndt); ode.set_params(p)
    finally:
        if not schedule:
            if old is None:
                os.environ.pop("KANODE_SCHEDULE", None)
            else:
                os.environ["KANODE_SCHEDULE"] = old
    lib = ode.lib
    fn = lib.kanode_loss_grad_dev_f64 if f64 else lib.kanode_loss_grad_dev
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                   creal, creal, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    sa = np.ascontiguousarray(SAVEAT)
    with torch.cuda.stream(stream):
        d_u0 = torch.tensor(u0, dtype=tdt, device=dev); d_tg = torch.tensor(tg, dtype=tdt, device=dev)
        d_grad = torch.zeros(ode.np_, dtype=tdt, device=dev); d_loss = torch.zeros(1, dtype=torch.float64, device=dev)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

        def step():
            rc = fn(ode.h, d_u0.data_ptr(), B, TSPAN[0], TSPAN[1], sa.ctypes.data, sa.size, d_tg.data_ptr(), 1e-6, 1e-3,
                    d_loss.data_ptr(), d_grad.data_ptr(), None, None, None)
            abi.check(lib, ode.h, rc, "kanode_loss_grad_dev")
        for _ in range(max(warmup, 3)):
            step()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.zero_(); evs[i][0].record(stream); step(); evs[i][1].record(stream)
        torch.cuda.synchronize()
    ms = sum(a_.elapsed_time(b_) for a_, b_ in evs) / steps
    ode.close()
    return {"ms_per_step": ms, "value": B / (ms / 1e3), "unit": "trajectories/s", "steps": steps, "batch_per_gpu": B}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; Julia is absent from the image) on all host
    cores, on the bench arm's own workload: every step is the full 65,536-trajectory ensemble."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chain, p, u0, tg = make_workload(args.batch, 1234)
    from oracle import Oracle
    orc = Oracle(chain.desc(), np.float64)
    cores = orc.set_threads(os.cpu_count() or 1)        # explicit: torchrun exports OMP_NUM_THREADS=1
    for _ in range(min(args.warmup, 2)):
        orc.loss_grad(p, u0[:4096], TSPAN, SAVEAT, tg[:4096])
    t = time.perf_counter()
    for _ in range(args.steps):
        orc.loss_grad(p, u0, TSPAN, SAVEAT, tg)
    dt = time.perf_counter() - t
    val = u0.shape[0] * args.steps / dt
    line = {"impl": "reference", "metric": "kan_ode_fwd_adjoint_trajectory_train_steps_per_s", "value": val,
            "unit": "trajectories/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "lotka_volterra_kan_ode_2_10_2_g5_ensemble", "batch_per_gpu": int(u0.shape[0]),
                       "global_batch": int(u0.shape[0]), "tspan": list(TSPAN), "nsave": int(SAVEAT.size), "abstol": 1e-6,
                       "reltol": 1e-3, "params": "glorot_uniform seed 0",
                       "note": "host CPU only: the whole ensemble of ONE GPU's shard per step, whatever --gpus says"},
            "cpu_baseline": {"value": val, "unit": "trajectories/s", "cores": cores, "kind": "port",
                             "sample": f"{u0.shape[0]} trajectories per step, fp64 C++ oracle (OpenMP, {cores} threads); Julia absent"},
            "e2e": {"value": val, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.lean:
        line["cfg1"] = cfg1_cpu_latency()
        line["workloads"] = {name: pde_cpu_baseline(name, cores) for name in PDE_WORKLOADS}
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import kan_odes_b200 as K
    from kan_odes_b200 import abi
    from kan_odes_b200.dist import packed_all_reduce

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
    esz = 8 if f64 else 4
    chain, p, u0, tg = make_workload(B, 1234 + rank)              # every rank trains its own shard (weak scaling)
    lib = abi.load_library()
    stream = torch.cuda.Stream()
    ode = K.KanOde(chain, device=local, stream=stream.cuda_stream, dtype=ndt)
    ode.set_params(p)
    npar = ode.np_
    dev = torch.device("cuda", local)
    # the step's collective: the library's own peer-memory kernel (pack + exchange over NVLink + sum in ONE launch, kanode_peer.cu);
    # KANODE_BENCH_NCCL=1 keeps the pack kernel + NCCL all-reduce for an A/B
    collective = "none"
    if world > 1:
        from kan_odes_b200.dist import peer_setup
        collective = "nccl all-reduce of one packed buffer"
        if not os.environ.get("KANODE_BENCH_NCCL") and peer_setup(ode):
            collective = "peer-memory kernel (CUDA IPC mailboxes over NVLink), no NCCL call in the step"
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
        if world > 1 and not os.environ.get("KANODE_BENCH_NO_ALLREDUCE"):
            # the only collective of a step: [gradient sum | loss sum | count] packed by one kernel, ONE all-reduce, no host read-back
            packed_all_reduce(ode, d_loss, d_grad, B)
            # (KANODE_BENCH_NO_ALLREDUCE=1 is a diagnostic: it isolates the collective's share of the step at N > 1; not a bench mode)

    align = world > 1 and not os.environ.get("KANODE_BENCH_NO_ALIGN") and not os.environ.get("KANODE_BENCH_NO_ALLREDUCE")
    ms3 = (C.c_float * 3)()
    prng = np.random.default_rng(99)                               # same on every rank: replicated parameters
    with torch.cuda.stream(stream):                                # resident like a device-side optimizer's output: the refresh is kernels only
        p_steps = [torch.tensor(p * (1.0 + 2e-3 * prng.standard_normal(p.shape)), dtype=torch.float32, device=dev) for _ in range(args.steps)]
    lib.kanode_set_params_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
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
            if not args.fixed_params:                              # training conditions: the parameters move between steps, so the
                # launch order predicted from the last step is not exact (untimed; device-side refresh, the host is not blocked)
                abi.check(lib, ode.h, lib.kanode_set_params_dev(ode.h, p_steps[i].data_ptr(), npar), "kanode_set_params_dev")
            flush.zero_()                                          # L2 flush between timed iterations (untimed)
            if align:
                # untimed device-side rendezvous: the refresh + flush above are not part of a step, but a rank that finishes
                # them late would make every other rank wait INSIDE its timed region (the step's collective is a rendezvous)
                packed_all_reduce(ode, d_loss, d_grad, B)
            evs[i][0].record(stream)
            step(False)
            evs[i][1].record(stream)
            if world == 1:
                lib.kanode_last_timing(ode.h, ms3)                 # per-kernel CUDA-event times of this step (blocks until it finished)
                k_ms[i] = [ms3[0], ms3[1], ms3[2]]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            abi.check(lib, ode.h, lib.kanode_peer_status(ode.h), "kanode_peer_status")   # a timed-out exchange is an error, not a number
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
        step_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = sum(step_ms)
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
    h2d = h_u0.nbytes + h_tg.nbytes + SAVEAT.nbytes + npar * esz
    d2h = 8 + npar * esz + 12                                      # loss sum, gradient sum, 3 failure counters (stats_scan_kernel)

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
    ode.close()
    workloads = None
    if not args.lean:
        workloads = {}
        for name, bpg in PDE_WORKLOADS.items():
            if world > 1:
                dist.barrier()
            workloads[name] = run_pde_workload(name, bpg, "f32", max(3, min(args.steps, 5)), 3, world, rank, local, read_peaks(),
                                               with_cpu=world == 1 and not args.no_cpu)   # CPU baselines: rank 0 at N = 1 only
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = sampler.result()
    peaks = read_peaks()
    hbm_peak = peaks["hbm_gbs"]; ffma_peak = peaks["ffma_tflops"]; mufu_peak = peaks["mufu_tops"]
    bwd_ms = float(k_ms[:, 1].mean()); fwd_ms = float(k_ms[:, 0].mean()); red_ms = float(k_ms[:, 2].mean())
    bwd_flop = nf_b * FLOP_BWD_EVAL
    achieved_tf = bwd_flop / (bwd_ms / 1e3) / 1e12
    # MUFU: 96 per RHS evaluation; the fused forward+VJP evaluation computes every activation once
    mufu_ops = nf_b * MUFU_PER_EVAL
    achieved_mufu = mufu_ops / (bwd_ms / 1e3) / 1e12
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get("small_backward_lg_kernel_bytes")
    except Exception:
        pass
    # algorithmic HBM bytes of the backward kernel: dense record read once, dL/du read once, per-warp gradient
    # partials written once (the per-trajectory gradient state never leaves the registers)
    rec_bytes = int(fst[:, 0].sum()) * 20 * esz
    alg_bytes = rec_bytes + B * SAVEAT.size * 2 * esz + ((B + 5) // 6) * npar * esz
    line = {
        "metric": "kan_ode_fwd_adjoint_trajectory_train_steps_per_s", "value": value, "unit": "trajectories/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "lotka_volterra_kan_ode_2_10_2_g5_ensemble", "batch_per_gpu": B,
                   "global_batch": world * B, "tspan": list(TSPAN), "nsave": int(SAVEAT.size), "abstol": 1e-6,
                   "reltol": 1e-3, "params": "glorot_uniform seed 0" + ("" if args.fixed_params else ", perturbed 0.2% (relative, seed 99) before every timed step"),
                   "l2": "256 MiB flush between timed steps" + (", then an untimed device-side rendezvous of the ranks" if align else ""),
                   "parallelism": f"dp{world} (trajectories sharded, gradient all-reduce only)", "collective": collective},
        "train_steps_per_s": args.steps / (total_ms / 1e3),
        "rhs_evals_per_s": world * (nf_f + nf_b) * args.steps / (total_ms / 1e3),
        "rhs_evals_per_step_per_gpu": {"forward": nf_f, "backward_fused_fwd_vjp": nf_b},
        "failed_trajectories": failed,
        "kernel_ms": {"forward": fwd_ms, "backward": bwd_ms, "grad_reduce": red_ms,
                      "backward_min": float(k_ms[:, 1].min()), "backward_max": float(k_ms[:, 1].max())},
        "step_ms_spread": {"min": float(min(step_ms)), "max": float(max(step_ms)), "note": "this rank's timed steps (CUDA events)"},
        "roofline": {"kernel": "small_backward_lg_kernel", "bound": "ffma", "achieved": achieved_tf,
                     "peak": ffma_peak, "unit": "TFLOP/s", "frac": achieved_tf / ffma_peak,
                     "peak_source": peaks["ffma_source"], "derived_peak": FFMA_PEAK_TFLOPS,
                     "flop_per_unit": FLOP_BWD_EVAL, "units_per_launch": nf_b, "traffic": traffic,
                     "mufu": {"ops_per_unit": MUFU_PER_EVAL, "achieved": achieved_mufu, "peak": mufu_peak, "unit": "Top/s",
                              "frac": achieved_mufu / mufu_peak, "peak_source": peaks["mufu_source"]},
                     "hbm": {"algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (bwd_ms / 1e3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (bwd_ms / 1e3) / 1e9 / hbm_peak,
                             "peak_source": peaks["hbm_source"]}},
        "e2e": {"value": e2e_val, "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if per_rank is not None:
        line["per_rank"] = per_rank
    if workloads is not None:
        line["workloads"] = workloads
    if world == 1 and not args.lean:
        other = "f32" if f64 else "f64"
        leg = lv_device_leg(chain, p, u0, tg, not f64, local, 3, 3)
        line[other] = leg; line[f"value_{other}"] = leg["value"]; line[f"ms_per_step_{other}"] = leg["ms_per_step"]
        nh = lv_device_leg(chain, p, u0, tg, f64, local, 3, 3, schedule=False)
        line["no_history"] = {"ms_per_step": nh["ms_per_step"], "value": nh["value"],
                              "note": "adjoint solves launched in natural order, NOT sorted by the previous call's step margins (KANODE_SCHEDULE=0)"}
        line["cfg1"] = cfg1_gpu_latency(local)
        if not args.no_cpu:
            line["parity"] = lv_parity(chain, p, u0, tg, ndt, local)
            line["cfg1"]["cpu_baseline"] = cfg1_cpu_latency()
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
    ap.add_argument("--no-cpu", action="store_true", help="skip every leg that runs the CPU oracle (cpu_baseline, parity)")
    ap.add_argument("--fixed-params", action="store_true", help="replay identical parameters every timed step (default: perturb them 0.2% per step)")
    ap.add_argument("--lean", action="store_true", help="headline workload only: no f64 leg, parity block, cfg1 latency, PDE workloads")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
