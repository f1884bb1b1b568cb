#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

metric   : accepted neighbour pair-evals/s of one 3D WCSPH step (update! + interact!), FP32
workload : config 3 of BASELINE.json at N = 1: 254^3 = 16 387 064 particles on a perturbed cubic
           lattice, search_radius = 3 * spacing, FullGridCellList, one step = update!(nhs, y, y;
           points_moving = (true, true)) on coordinates perturbed by sigma = 4e-4 * r
           (benchmarks/update.jl:28-49) + the WCSPH continuity + momentum interaction
           (benchmarks/smoothed_particle_hydrodynamics.jl:45-102).  At N > 1 every rank owns a
           slab of 254 lattice layers (weak scaling: 254 x 254 x 254N particles, ~config 5 at
           N = 8) and exchanges one ghost cell layer per step over NCCL.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm
                                                                (C restatement, all host threads)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(REPO, "pointneighbors.jl_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "neighbour pair-evals/s (3D WCSPH step incl. update!, FP32)"
UNIT = "pair-evals/s"
SM_COUNT = 148
FP32_LANES = 128


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# synthetic input (the distribution of test/point_cloud.jl + benchmarks/run_benchmarks.jl:81-89)
# ------------------------------------------------------------------------------------------------
def lattice_cloud_torch(dims, domain_n, z0, seed, device):
    """Perturbed lattice slab generated on the device: lattice indices 1..dims[0] x 1..dims[1] x
    (z0+1)..(z0+dims[2]), sigma = 0.05 applied twice, cell-sorted with dimension 1 most
    significant, Float32, normalised by (domain_n + 1)."""
    import torch
    nx, ny, nz = dims
    N = nx * ny * nz
    gen = torch.Generator(device=device).manual_seed(seed)
    k = torch.arange(N, device=device, dtype=torch.int64)
    ix = (k % nx).to(torch.float64) + 1.0
    iy = ((k // nx) % ny).to(torch.float64) + 1.0
    iz = (k // (nx * ny)).to(torch.float64) + 1.0 + z0
    c = torch.stack([ix, iy, iz], dim=1)
    c += 0.05 * torch.randn(N, 3, device=device, dtype=torch.float64, generator=gen)
    cell = torch.floor(c / 3.0).to(torch.int64)
    c += 0.05 * torch.randn(N, 3, device=device, dtype=torch.float64, generator=gen)
    key = (cell[:, 0] * (ny + 8) + cell[:, 1]) * (nz + z0 + 8) + cell[:, 2]
    perm = torch.sort(key, stable=True).indices
    c = c[perm]
    out = (c.to(torch.float32) / np.float32(domain_n + 1)).contiguous()
    return out


def wcsph_state_torch(N, r, seed, device):
    """benchmarks/smoothed_particle_hydrodynamics.jl:54-95: rho = 1000 + rand, v = 0,
    m = 0.1 * spacing, Cole EOS (exponent 1, c = 10), h = r / 2."""
    import torch
    gen = torch.Generator(device=device).manual_seed(seed)
    rho = (1000.0 + torch.rand(N, device=device, generator=gen, dtype=torch.float32))
    v = torch.zeros((N, 4), device=device, dtype=torch.float32)
    v[:, 3] = rho
    mass = torch.full((N,), float(np.float32(0.1) * (np.float32(r) / np.float32(3))),
                      device=device, dtype=torch.float32)
    pressure = (np.float32(100.0) * (rho - np.float32(1000.0))).contiguous()
    return v.contiguous(), mass, pressure


class ClockSampler(threading.Thread):
    """Sample SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU algorithm (C restatement in oracle/, the reference is pure Julia and
# cannot run here) on all host threads, on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(n_lattice, seed=1):
    from oracle import pn_oracle
    import pnb200
    T = np.float32
    c, r, mn, mx = pnb200.benchmark_cloud((n_lattice,) * 3, seed=seed)
    rng = np.random.default_rng(seed + 1)
    c2 = (c + (T(4e-4) * r) * rng.standard_normal(c.shape).astype(T)).astype(T)
    N = len(c)
    rho = (T(1000) + rng.random(N).astype(T)).astype(T)
    v = np.zeros((N, 4), T)
    v[:, 3] = rho
    mass = np.full(N, T(0.1) * (r / T(3)), T)
    pressure = (T(100) * (rho - T(1000))).astype(T)
    h = T(r / T(2))
    params = np.array([h, 10.0, 0.02, 0.0, 0.01, 0.1, pnb200.wendland_c2_norm(3, h)], T)
    g = pn_oracle.Grid(3, r, mn, mx)
    g.build(c)
    pairs = [int(g.count_neighbors(c, c).sum())]
    g.build(c2)
    pairs.append(int(g.count_neighbors(c2, c2).sum()))
    coords = [c, c2]

    def step(s):
        y = coords[(s + 1) % 2]
        g.build_dvov(y)                                   # update! (ParallelUpdate, DVoV layout)
        g.wcsph(y, y, v, v, mass, mass, pressure, pressure, params, use_dvov=True, parallel=True)
        return pairs[(s + 1) % 2]

    return step, N, pn_oracle.max_threads()


def run_cpu_arm(steps, warmup, budget_s=120.0):
    """Times `steps` steps of the CPU restatement on a sample sized to fit the budget."""
    step64, N64, threads = cpu_step_factory(64)
    t0 = time.perf_counter()
    step64(0)
    t64 = time.perf_counter() - t0
    per_point = t64 / N64
    n_s = 64
    for cand in (96, 128, 160, 200, 254):
        if per_point * cand ** 3 * (steps + warmup) <= budget_s:
            n_s = cand
    step, N, threads = (step64, N64, threads) if n_s == 64 else cpu_step_factory(n_s)
    for s in range(warmup):
        step(s)
    t0 = time.perf_counter()
    pairs = 0
    for s in range(steps):
        pairs += step(s)
    dt = time.perf_counter() - t0
    return {"value": pairs / dt, "ms_per_step": 1e3 * dt / steps, "cores": threads,
            "sample": f"{n_s}^3 = {N} particles of the same perturbed lattice (density, r = 3 spacings "
                      f"identical), {steps} steps of update! (DVoV atomic push) + WCSPH interact!",
            "n_sample": N}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = run_cpu_arm(args.steps, max(args.warmup, 1), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "WCSPH step 3D (config 3), bounded CPU sample: " + res["sample"]},
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "reference is pure Julia (no julia binary in this image): this is the C/OpenMP "
                "restatement of its CPU path (oracle/), reference data structures, all host threads",
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import pnb200 as pn
    from pnb200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pnb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 or args.total_layers > 0:
        import torch.distributed as dist
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        else:
            # one rank of the slab code path (strong-scaling base line of --total-layers)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        from pnb200 import slabs
        return slabs.bench_multi_gpu(args, rank, world, dev, METRIC, UNIT)

    n = args.lattice
    T = np.float32
    N = n ** 3
    r = T(3.0) / T(n + 1)
    mn, mx = np.zeros(3, T), np.ones(3, T)
    A = lattice_cloud_torch((n, n, n), n, 0, 1, dev)
    gen = torch.Generator(device=dev).manual_seed(2)
    B = (A + (T(4e-4) * r) * torch.randn(N, 3, device=dev, generator=gen)).contiguous()
    v, mass, pressure = wcsph_state_torch(N, r, 3, dev)
    dv = torch.zeros((N, 4), device=dev, dtype=torch.float32)
    nhs = pn.GridNeighborhoodSearch[3](
        search_radius=r, n_points=N,
        cell_list=pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=r))
    h = T(r / T(2))
    closure = pn.WCSPHInteract(dv, v, v, mass, mass, pressure, pressure, smoothing_length=h,
                               sound_speed=T(10.0), alpha=T(0.02), beta=T(0.0), delta=T(0.1))
    coords = [A, B]

    def step(s):
        y = coords[(s + 1) % 2]
        pn.update_(nhs, y, y, points_moving=(True, True))
        pn.foreach_point_neighbor(closure, y, y, nhs)

    # accepted pairs P and candidate tests K_ref of both clouds (device side, untimed)
    cnt = torch.zeros(N, dtype=torch.int64, device=dev)
    pairs, kref = [], []
    for y in coords:
        pn.initialize_(nhs, y, y)
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), y, y, nhs)
        pairs.append(int(cnt.sum()))
        cs, _ = nhs.export_csr()
        gs = nhs.cell_list.n_cells_per_dimension
        counts = (cs[1:] - cs[:-1]).to(torch.float64).reshape(gs[2], gs[1], gs[0])
        nb = torch.nn.functional.conv3d(counts[None, None], torch.ones(1, 1, 3, 3, 3, device=dev,
                                                                       dtype=torch.float64), padding=1)
        kref.append(int((counts * nb[0, 0]).sum().item()))
    C_cells = nhs.total_cells()

    for s in range(args.warmup):
        step(s)
    torch.cuda.synchronize()
    _lib.profile(enable=True, reset=True)
    _lib.profile(reset=True)
    launches0 = int(_lib.lib().pnb_launch_count())
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    upd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
    torch.cuda.synchronize()
    ev0.record()
    total_pairs = 0
    for s in range(args.steps):
        y = coords[(s + 1) % 2]
        upd_ev[s][0].record()
        pn.update_(nhs, y, y, points_moving=(True, True))
        upd_ev[s][1].record()
        pn.foreach_point_neighbor(closure, y, y, nhs)
        total_pairs += pairs[(s + 1) % 2]
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = int(_lib.lib().pnb_launch_count()) - launches0
    prof = _lib.profile(enable=False)
    update_ms = float(np.mean([a.elapsed_time(b) for a, b in upd_ev]))
    value = total_pairs / (ms_total * 1e-3)

    # ---- e2e: same step through the public API with HOST (pinned) buffers ---------------------
    # Every step copies its inputs (coordinates, state, pressure) from pinned host memory and
    # copies its result (dv) back.  Two measurements:
    #   serial    : H2D -> update! -> interact! -> D2H on one stream (latency of one step)
    #   pipelined : three streams, double-buffered device arrays: the H2D of step s+1 and the
    #               D2H of step s-1 overlap the kernels of step s (throughput; this is `value`)
    hA, hB = A.cpu().pin_memory(), B.cpu().pin_memory()
    hv, hp = v.cpu().pin_memory(), pressure.cpu().pin_memory()
    hdv = [torch.empty((N, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    dy = [torch.empty_like(A) for _ in range(2)]
    dvv = [v, torch.empty_like(v)]
    dpp = [pressure, torch.empty_like(pressure)]
    dvo = [dv, torch.empty_like(dv)]
    cls = [closure, pn.WCSPHInteract(dvo[1], dvv[1], dvv[1], mass, mass, dpp[1], dpp[1],
                                     smoothing_length=h, sound_speed=T(10.0), alpha=T(0.02),
                                     beta=T(0.0), delta=T(0.1))]
    hcoords = [hA, hB]

    def e2e_serial_step(s):
        dy[0].copy_(hcoords[(s + 1) % 2], non_blocking=True)
        dvv[0].copy_(hv, non_blocking=True)
        dpp[0].copy_(hp, non_blocking=True)
        pn.update_(nhs, dy[0], dy[0], points_moving=(True, True))
        pn.foreach_point_neighbor(cls[0], dy[0], dy[0], nhs)
        hdv[0].copy_(dvo[0], non_blocking=True)

    e2e_steps = max(4, min(args.steps, 10))
    for s in range(2):
        e2e_serial_step(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(e2e_steps):
        e2e_serial_step(s)
    e1.record()
    torch.cuda.synchronize()
    e2e_serial_ms = e0.elapsed_time(e1) / e2e_steps

    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def issue_h2d(s):
        bb = s % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp[bb])           # the kernels of step s-2 are done with these buffers
            dy[bb].copy_(hcoords[(s + 1) % 2], non_blocking=True)
            dvv[bb].copy_(hv, non_blocking=True)
            dpp[bb].copy_(hp, non_blocking=True)
            ev_in[bb].record(s_in)

    def run_pipelined(n_steps):
        pairs_done = 0
        issue_h2d(0)
        for s in range(n_steps):
            bb = s % 2
            if s + 1 < n_steps:
                issue_h2d(s + 1)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[bb])
                s_cmp.wait_event(ev_out[bb])      # dv buffer of step s-2 has reached the host
                pn.update_(nhs, dy[bb], dy[bb], points_moving=(True, True))
                pn.foreach_point_neighbor(cls[bb], dy[bb], dy[bb], nhs)
                ev_cmp[bb].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[bb])
                hdv[bb].copy_(dvo[bb], non_blocking=True)
                ev_out[bb].record(s_out)
            pairs_done += pairs[(s + 1) % 2]
        return pairs_done

    torch.cuda.synchronize()
    run_pipelined(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    e2e_pairs = run_pipelined(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)      # wall clock across the three streams
    h2d = int(hA.numel() * 4 + hv.numel() * 4 + hp.numel() * 4)
    d2h = int(hdv[0].numel() * 4)

    # ---- roofline -----------------------------------------------------------------------------
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    sweep_ms, sweep_n = prof["k_sweep_cells"]
    sweep_avg = sweep_ms / max(sweep_n, 1)
    bytes_sweep = 56 * N + 4 * (C_cells + 1)          # WCSPH interact
    ach = bytes_sweep / (sweep_avg * 1e-3) / 1e9
    # update! = one-pass bucket build (k_bucket_scatter) after the first build; the two-pass CSR
    # build (hist + scan + scatter) when a cell overflowed its bucket or PNB_BUILD_LAYOUT=0
    build_names = [k for k in ("k_bucket_scatter", "k_cell_hist", "k_scan_lookback", "k_scatter_points")
                   if prof.get(k, (0.0, 0))[1]]
    build_ms = sum(prof[k][0] for k in build_names) / args.steps
    bytes_update = 28 * N + 4 * (C_cells + 1)          # 16N + 4(C+1) + 12N (cell-ordered coordinates)
    ach_u = bytes_update / (build_ms * 1e-3) / 1e9
    P_avg = float(np.mean(pairs))
    K_avg = float(np.mean(kref))
    flops = 8.0 * K_avg + 70.0 * P_avg                 # non-FMA FP32 operations
    fp32_peak = SM_COUNT * FP32_LANES * sm_max_mhz * 1e6 / 1e12
    # the whole interaction's flops over ALL its sweep kernels: the tile kernel plus the overflow /
    # surplus-point kernels that finish the same sweep (k_sweep_overflow phase)
    ovf_ms, ovf_n = prof.get("k_sweep_overflow", (0.0, 0))
    sweep_all = sweep_avg + (ovf_ms / max(sweep_n, 1))
    fp32_ach = flops / (sweep_all * 1e-3) / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"WCSPH step 3D: {n}^3 = {N} particles (BASELINE config 3), "
                               "update!(points_moving=(true,true)) on sigma=4e-4*r perturbed "
                               "coordinates + continuity+momentum interact!",
                   "search_radius": float(r), "cells": C_cells, "pairs_per_step": P_avg,
                   "candidate_tests_per_step": K_avg, "velocities": "zero (reference benchmark)",
                   "l2": "inputs (197 MB coordinates + 328 MB state) larger than the 126 MB L2"},
        "update_ms": update_ms,
        "interact_ms": (ms_total / args.steps) - update_ms,
        "points_per_s": N / (ms_total / args.steps * 1e-3),
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": e2e_pairs / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                "mode": "pipelined: 3 streams, double-buffered device arrays, every step copies "
                        "its inputs from pinned host memory and its dv back (wall clock)",
                "serial_ms_per_step": e2e_serial_ms,
                "serial_value": float(np.mean(pairs)) / (e2e_serial_ms * 1e-3)},
        "roofline": {"bound": "hbm", "kernel": "k_sweep_tiles<3,false,WcsphClT<false>,4,true>", "achieved": ach,
                     "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this exact
                     # workload, from the ncu --set full capture summarised in
                     # profiles/r1_wcsph_sweep_v6_ncu_summary.txt (900.8 MB + 255.3 MB)
                     "traffic": 1156059136 if n == 254 else None,
                     "peak_source": peak_src, "launch_ms": sweep_avg,
                     "algorithmic_bytes": bytes_sweep,
                     "note": "the fused interaction is FP32-issue bound (~100 flop/B), not HBM bound "
                             "(SURVEY.md 8d); see fp32_pipe for the binding roofline"},
        "fp32_pipe": {"achieved_tflops": fp32_ach, "peak_tflops": fp32_peak,
                      "frac": fp32_ach / fp32_peak,
                      "sweep_ms": sweep_all,
                      "model": "8*K_ref + 70*P non-FMA operations over the tile kernel + the overflow / "
                               "surplus-point kernels of the same sweep; peak = 148 SM x 128 lanes x max clock"},
        "roofline_update": {"bound": "hbm", "kernels": build_names, "achieved": ach_u,
                            "peak": hbm_peak, "unit": "GB/s", "frac": ach_u / hbm_peak,
                            # one-pass bucket build: profiles/r1_update_v4_bucket_ncu_summary.txt
                            # (209.5 MB read + 223.6 MB written per launch at this workload)
                            "traffic": 433141504 if (n == 254 and "k_bucket_scatter" in build_names) else None,
                            "device_ms": build_ms, "algorithmic_bytes": bytes_update,
                            "per_kernel_ms": {k: prof[k][0] / max(prof[k][1], 1) for k in build_names},
                            "call_ms": update_ms,
                            "layout": "buckets (one pass)" if "k_bucket_scatter" in build_names
                                      else "CSR (two passes)"},
        "kernel_ms": {k: (v_[0] / v_[1] if v_[1] else 0.0) for k, v_ in prof.items()},
    }
    if not args.no_cpu_baseline:
        cb = run_cpu_arm(3, 1, budget_s=25.0)
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"],
                                "kind": "port", "sample": cb["sample"]}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lattice", type=int, default=254, help="lattice points per dimension (per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--total-layers", type=int, default=0,
                    help="STRONG scaling: fixed lattice 254 x 254 x TOTAL_LAYERS split over the GPUs "
                         "(2032 = the 131 M particle cloud of BASELINE config 5); 0 = weak scaling, "
                         "254 layers per GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
