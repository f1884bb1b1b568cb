#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

metric   : accepted neighbour pair-evals/s of one 3D WCSPH step (update! + interact!), FP32
workload : N = 1: config 3 of BASELINE.json: 254^3 = 16 387 064 particles on a perturbed cubic
           lattice, search_radius = 3 * spacing, FullGridCellList, one step = update!(nhs, y, y;
           points_moving = (true, true)) on coordinates perturbed by sigma = 4e-4 * r
           (benchmarks/update.jl:28-49) + the WCSPH continuity + momentum interaction
           (benchmarks/smoothed_particle_hydrodynamics.jl:45-102).
           N > 1: config 5: the 504^3 = 128 024 064 particle cloud (171^3 cells) slab-decomposed
           along the last cell dimension over the N GPUs (STRONG scaling), one migrant + ghost
           layer exchange per step over NCCL, overlapped with the sweep of the interior layers
           (pnb200/slabs.py).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm
                                                                (C restatement, all host threads)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
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
T = np.float32


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def git_head():
    try:
        return subprocess.run(["git", "-C", REPO, "rev-parse", "--short", "HEAD"], capture_output=True,
                              text=True, timeout=5).stdout.strip() or None
    except Exception:
        return None


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at config 3, from the
    ncu --set full capture recorded in profiles/r2_traffic.json (with the git hash and the summary
    file it came from) -- never a constant in this file."""
    path = os.path.join(REPO, "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f).get(kernel)
    except Exception:
        return None, None
    if not rec:
        return None, None
    return int(rec["dram_bytes"]), {k: rec.get(k) for k in ("git", "capture", "workload")}


def workload_config(n):
    """`config` of the JSON line: identical in the GPU arm and the reference arm."""
    N = n ** 3
    r = T(3.0) / T(n + 1)
    return {"workload": f"WCSPH step 3D: {n}^3 = {N} particles (BASELINE config 3), "
                        "update!(points_moving=(true,true)) on sigma=4e-4*r perturbed "
                        "coordinates + continuity+momentum interact!",
            "lattice": n, "particles": N, "search_radius": float(r),
            "velocities": "zero (reference benchmark)",
            "l2": "inputs (197 MB coordinates + 328 MB state) larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# synthetic input (the distribution of test/point_cloud.jl + benchmarks/run_benchmarks.jl:81-89)
# ------------------------------------------------------------------------------------------------
def lattice_cloud_torch(dims, domain_n, z0, seed, device):
    import pnb200
    return pnb200.benchmark_cloud_torch(dims, domain_n, z0=z0, seed=seed, device=device)[0]


def lattice_cloud_numpy(n, seed):
    """The same distribution on the host (reference arm: no CUDA on its path), with one combined
    sort key instead of a lexsort of three."""
    rng = np.random.default_rng(seed)
    k = np.arange(n ** 3, dtype=np.int64)
    c = np.empty((n ** 3, 3), np.float64)
    c[:, 0] = k % n + 1
    c[:, 1] = (k // n) % n + 1
    c[:, 2] = k // (n * n) + 1
    c += 0.05 * rng.standard_normal(c.shape)
    cell = np.floor(c / 3.0).astype(np.int64)
    key = (cell[:, 0] * (n + 8) + cell[:, 1]) * (n + 8) + cell[:, 2]
    c += 0.05 * rng.standard_normal(c.shape)
    c = c[np.argsort(key, kind="stable")]
    return np.ascontiguousarray((c.astype(T) / T(n + 1)).astype(T))


def wcsph_state_torch(N, r, seed, device, moving=False):
    """benchmarks/smoothed_particle_hydrodynamics.jl:54-95: rho = 1000 + rand, v = 0,
    m = 0.1 * spacing, Cole EOS (exponent 1, c = 10), h = r / 2."""
    import torch
    gen = torch.Generator(device=device).manual_seed(seed)
    rho = (1000.0 + torch.rand(N, device=device, generator=gen, dtype=torch.float32))
    v = torch.zeros((N, 4), device=device, dtype=torch.float32)
    if moving:
        v[:, :3] = 0.1 * torch.randn(N, 3, device=device, generator=gen, dtype=torch.float32)
    v[:, 3] = rho
    mass = torch.full((N,), float(T(0.1) * (T(r) / T(3))), device=device, dtype=torch.float32)
    pressure = (T(100.0) * (rho - T(1000.0))).contiguous()
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
# cannot run here) on ALL host threads, always on the headline cloud (254^3); a step = update! of
# the whole cloud + interact! over a bounded, contiguous sample of the points
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_cpu_arm(n, steps, warmup, budget_s):
    """update! (DVoV atomic push, the reference's ParallelUpdate, src/nhs_grid.jl:255-281) of the
    WHOLE n^3 cloud + WCSPH interact! over the first m points (a slab of the cloud, which is
    sorted with dimension 1 most significant), m chosen so that steps + warmup fit budget_s.
    torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly to all host cores."""
    from oracle import pn_oracle
    import pnb200
    threads = host_threads()
    pn_oracle.set_threads(threads)
    N = n ** 3
    r = T(3.0) / T(n + 1)
    mn, mx = np.zeros(3, T), np.ones(3, T)
    c = lattice_cloud_numpy(n, 1)
    rng = np.random.default_rng(2)
    c2 = (c + (T(4e-4) * r) * rng.standard_normal(c.shape).astype(T)).astype(T)
    rho = (T(1000) + rng.random(N, dtype=np.float32)).astype(T)
    v = np.zeros((N, 4), T)
    v[:, 3] = rho
    mass = np.full(N, T(0.1) * (r / T(3)), T)
    pressure = (T(100) * (rho - T(1000))).astype(T)
    h = T(r / T(2))
    params = np.array([h, 10.0, 0.02, 0.0, 0.01, 0.1, pnb200.wendland_c2_norm(3, h)], T)
    g = pn_oracle.Grid(3, r, mn, mx)
    coords = [c, c2]
    # calibration: one update! and the interaction of 1 % of the points
    t0 = time.perf_counter()
    g.build_dvov(c)
    t_upd = time.perf_counter() - t0
    probe = np.arange(0, N, 100, dtype=np.int64)
    t0 = time.perf_counter()
    g.wcsph(c, c, v, v, mass, mass, pressure, pressure, params, points=probe, use_dvov=True, parallel=True)
    t_pp = (time.perf_counter() - t0) / len(probe)
    per_step = budget_s / max(steps + warmup, 1)
    frac = min(1.0, max(0.02, (per_step - t_upd) / (t_pp * N)))
    m = N if frac >= 0.999 else int(frac * N)
    sample = None if m == N else np.arange(m, dtype=np.int64)
    pairs = []
    for y in coords:
        g.build_dvov(y)
        cn = g.count_neighbors(y, y, points=sample, use_dvov=True)
        pairs.append(int(cn.sum() if sample is None else cn[:m].sum()))
    t_u, t_i = [], []

    def step(s, timed):
        y = coords[(s + 1) % 2]
        a = time.perf_counter()
        g.build_dvov(y)                                   # update! (ParallelUpdate, DVoV layout)
        b = time.perf_counter()
        g.wcsph(y, y, v, v, mass, mass, pressure, pressure, params, points=sample, use_dvov=True,
                parallel=True)
        e = time.perf_counter()
        if timed:
            t_u.append(b - a)
            t_i.append(e - b)
        return pairs[(s + 1) % 2]

    for s in range(warmup):
        step(s, False)
    total_pairs = 0
    for s in range(steps):
        total_pairs += step(s, True)
    f = m / N
    # the sample's share of the (whole-cloud) update! is charged, its interact! in full
    dt = float(np.sum(t_i) + f * np.sum(t_u))
    what = (f"update! (DVoV atomic push) of all {N} particles + WCSPH interact! over "
            + ("all of them" if m == N else f"the first {m} ({100 * f:.1f} %, a slab of the cloud; its "
               f"share of the update! time is charged)"))
    return {"value": total_pairs / dt, "ms_per_step": 1e3 * dt / steps / f, "cores": threads,
            "sample": what, "n_sample": m, "update_ms": 1e3 * float(np.mean(t_u)),
            "interact_ms_sample": 1e3 * float(np.mean(t_i))}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.lattice
    res = run_cpu_arm(n, args.steps, min(args.warmup, 1), budget_s=170.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "sample": res["sample"], "update_ms": res["update_ms"],
                         "interact_ms_sample": res["interact_ms_sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "reference is pure Julia (no julia binary in this image): this is the C/OpenMP "
                "restatement of its CPU path (oracle/), reference data structures, all host threads; "
                "ms_per_step is extrapolated to the whole cloud from the sample",
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def timed_events(fn, reps, flush_buf=None, warm=3):
    import torch
    for _ in range(warm):
        fn()
    times = []
    for _ in range(reps):
        if flush_buf is not None:
            flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.min(times)), float(np.median(times))


def sub_configs(dev, hbm_peak, fp32_peak):
    """BASELINE configs 1, 2 and 4 at full size: device time (CUDA events around the API call, L2
    flushed between repetitions where the working set fits it), roofline figures, and a check of
    the result against the CPU oracle (all points for config 1, samples for 2 and 4)."""
    import torch
    import pnb200 as pn
    from oracle import pn_oracle
    pn_oracle.set_threads(host_threads())
    out = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    mn, mx = np.zeros(3, T), np.ones(3, T)

    def grid(n, A, box=None):
        N = A.shape[0]
        r = T(3.0) / T(n + 1)
        if box is None:
            cl = pn.FullGridCellList(min_corner=mn, max_corner=mx, search_radius=r)
            nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, cell_list=cl)
        else:
            cl = pn.FullGridCellList(min_corner=box[0], max_corner=box[1], search_radius=r)
            nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, cell_list=cl,
                                               periodic_box=pn.PeriodicBox(min_corner=box[0], max_corner=box[1]))
        pn.initialize_(nhs, A, A)
        return nhs, r

    # ---- config 1: count_neighbors, 64^3 -------------------------------------------------------
    n = 64
    A = lattice_cloud_torch((n, n, n), n, 0, 11, dev)
    nhs, r = grid(n, A)
    N, C = A.shape[0], nhs.total_cells()
    cnt = torch.zeros(N, dtype=torch.int64, device=dev)
    f = pn.CountNeighbors(cnt)
    pn.update_(nhs, A, A)
    t_min, t_med = timed_events(lambda: pn.foreach_point_neighbor(f, A, A, nhs), 20, flush)
    u_min, _ = timed_events(lambda: pn.update_(nhs, A, A, blocking=False), 20, flush)
    c = A.cpu().numpy()
    og = pn_oracle.Grid(3, r, mn, mx)
    og.build(c)
    ref = og.count_neighbors(c, c)
    P = int(ref.sum())
    K = og.candidate_tests(c)
    ok = bool((cnt.cpu().numpy() == ref).all())
    by = 24 * N + 4 * (C + 1)
    out["config1_count_64"] = {
        "workload": "count_neighbors, 64^3 = 262144 points (BASELINE config 1)", "pairs": P,
        "sweep_ms_min": t_min, "sweep_ms_median": t_med, "value": P / (t_min * 1e-3), "unit": UNIT,
        "update_ms_min": u_min, "l2": "flushed between repetitions (256 MB write)",
        "roofline": {"bound": "hbm", "achieved": by / (t_min * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": by / (t_min * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": by},
        "fp32_pipe_frac": 8.0 * K / (t_min * 1e-3) / 1e12 / fp32_peak,
        "oracle_checked": {"what": "neighbour counts of all points", "equal": ok}}
    del nhs, A, cnt

    # ---- config 2: n-body, 101^3 ---------------------------------------------------------------
    n = 101
    A = lattice_cloud_torch((n, n, n), n, 0, 12, dev)
    nhs, r = grid(n, A)
    N, C = A.shape[0], nhs.total_cells()
    gen = torch.Generator(device=dev).manual_seed(5)
    m_nb = (1e10 * (torch.rand(N, device=dev, generator=gen) + 1)).to(torch.float32)
    dv3 = torch.zeros((N, 3), device=dev)
    G = T(6.6743e-11)
    f = pn.NBodyGravity(dv3, m_nb, G)
    pn.update_(nhs, A, A)
    t_min, t_med = timed_events(lambda: pn.foreach_point_neighbor(f, A, A, nhs), 20, flush)
    c = A.cpu().numpy()
    og = pn_oracle.Grid(3, r, mn, mx)
    og.build(c)
    pts = np.arange(0, N, 37, dtype=np.int64)
    _, r64, rabs = og.nbody(c, c, m_nb.cpu().numpy(), G, points=pts, wide=True)
    got = dv3[torch.as_tensor(pts, device=dev)].cpu().numpy()
    ok = bool(np.all(np.abs(got - r64[pts]) <= 1e-5 * rabs[pts] + 1e-30))
    cntd = torch.zeros(N, dtype=torch.int64, device=dev)
    pn.foreach_point_neighbor(pn.CountNeighbors(cntd), A, A, nhs)
    P = int(cntd.sum())
    ok = ok and bool((cntd[torch.as_tensor(pts, device=dev)].cpu().numpy() ==
                      og.count_neighbors(c, c, points=pts)[pts]).all())
    K = og.candidate_tests(c)
    by = 32 * N + 4 * (C + 1)
    out["config2_nbody_101"] = {
        "workload": "n-body gravity, 101^3 = 1030301 points (BASELINE config 2)", "pairs": P,
        "sweep_ms_min": t_min, "sweep_ms_median": t_med, "value": P / (t_min * 1e-3), "unit": UNIT,
        "l2": "flushed between repetitions (256 MB write)",
        "roofline": {"bound": "hbm", "achieved": by / (t_min * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": by / (t_min * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": by},
        "fp32_pipe_frac": (8.0 * K + 14.0 * P) / (t_min * 1e-3) / 1e12 / fp32_peak,
        "oracle_checked": {"what": f"dv of {len(pts)} sampled points within 1e-5 * sum|term|, their counts exact",
                           "equal": ok}}
    del nhs, A, dv3, m_nb, cntd

    # ---- config 4: neighbour lists + TLSPH deformation gradient, 200^3 periodic ---------------
    n = 200
    s_ = T(1.0) / T(n + 1)
    A = lattice_cloud_torch((n, n, n), n, 0, 13, dev)
    bmn, bmx = np.full(3, s_ / T(2), T), np.full(3, (T(n) + T(0.5)) * s_, T)
    nhs, r = grid(n, A, box=(bmn, bmx))
    N = A.shape[0]
    pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=N, periodic_box=nhs.periodic_box,
                                              update_neighborhood_search=nhs, max_neighbors=160,
                                              transpose_backend=True)
    pn.initialize_(pre, A, A)
    b_min, b_med = timed_events(lambda: pn.update_(pre, A, A), 4, None, warm=1)
    off, ids = pre.export_csr()
    P = int(off[-1])
    gen = torch.Generator(device=dev).manual_seed(6)
    xcur = (A + (T(0.01) * r) * torch.sin(2 * np.pi * A)).contiguous()
    m0 = torch.full((N,), 0.1, device=dev)
    rho0 = torch.full((N,), 1000.0, device=dev)
    Lm = (torch.eye(3, device=dev).reshape(1, 9) + 0.05 * torch.randn(N, 9, device=dev, generator=gen)).contiguous()
    F = torch.zeros((N, 9), device=dev)
    h = T(r / T(2))
    fd = pn.TLSPHDeformationGradient(F, xcur, m0, rho0, Lm, smoothing_length=h, ndims_=3)
    s_min, s_med = timed_events(lambda: pn.foreach_point_neighbor(fd, A, A, pre), 10, None)
    c = A.cpu().numpy()
    og = pn_oracle.Grid(3, r, bmn, bmx, periodic_box=(bmn, bmx))
    og.build(c)
    pts = np.arange(0, N, 2003, dtype=np.int64)
    ro, ri = og.neighbor_lists(c[pts], c, sort=True)
    sel = torch.as_tensor(pts, device=dev)
    lens = off[sel + 1] - off[sel]
    ok = bool((lens.cpu().numpy() == np.diff(ro)).all())
    if ok:
        rep = torch.repeat_interleave(torch.arange(len(pts), device=dev), lens)
        pos = torch.arange(int(lens.sum()), device=dev) - torch.as_tensor(ro[:-1], device=dev)[rep]
        ok = bool((ids[off[sel][rep] + pos].cpu().numpy() == ri).all())
    lens_full = np.zeros(N, np.int64)
    lens_full[pts] = np.diff(ro)
    off_sparse = np.concatenate([[0], np.cumsum(lens_full)])
    _, r64, rabs = pn_oracle.tlsph_deformation_grad(c, xcur.cpu().numpy(), off_sparse, ri, m0.cpu().numpy(),
                                                    rho0.cpu().numpy(), Lm.cpu().numpy(), h, fd.kernel_norm,
                                                    r, periodic_box=(bmn, bmx), wide=True)
    ok = ok and bool(np.all(np.abs(F[sel].cpu().numpy() - r64[pts]) <= 1e-5 * rabs[pts] + 1e-30))
    by_build = 12 * N + 4 * N + 4 * (nhs.total_cells() + 1) + 4 * (N + 1) + 4 * P
    by_sweep = 108 * N + 4 * P + 4
    out["config4_nlist_tlsph_200_periodic"] = {
        "workload": "PrecomputedNeighborhoodSearch build (sorted lists) + TLSPH deformation gradient, "
                    "200^3 = 8000000 points, PeriodicBox (BASELINE config 4)", "pairs": P,
        "list_build_ms_min": b_min, "list_build_ms_median": b_med,
        "tlsph_sweep_ms_min": s_min, "tlsph_sweep_ms_median": s_med,
        "value": P / ((b_min + s_min) * 1e-3), "unit": UNIT,
        "l2": "working set (3.5 GB of lists) far larger than the L2",
        "roofline_build": {"bound": "hbm", "achieved": by_build / (b_min * 1e-3) / 1e9, "peak": hbm_peak,
                           "unit": "GB/s", "frac": by_build / (b_min * 1e-3) / 1e9 / hbm_peak,
                           "algorithmic_bytes": by_build},
        "roofline": {"bound": "hbm", "kernel": "k_tlsph_defgrad", "achieved": by_sweep / (s_min * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": by_sweep / (s_min * 1e-3) / 1e9 / hbm_peak,
                     "algorithmic_bytes": by_sweep},
        "oracle_checked": {"what": f"sorted lists and F of {len(pts)} sampled points (lists exact, F within "
                                   "1e-5 * sum|term|)", "equal": ok}}
    return out


def config5_one_gpu(steps=5, timeout_s=240):
    """`bench.py --slabs --slab-quick` in a child process: 504^3 = 128 M particles through the slab
    code path on one GPU (what N = 2, 4, 8 are compared with).  Never raises."""
    import subprocess
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--slabs", "--slab-quick", "--steps",
                              str(steps), "--warmup", "3"], capture_output=True, text=True,
                             timeout=timeout_s, env=env)
        rows = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if not rows:
            return {"error": (out.stderr or out.stdout)[-400:]}
        d = json.loads(rows[-1])
        return {"workload": d["config"]["workload"], "metric": d["metric"], "value": d["value"],
                "unit": d["unit"], "ms_per_step": d["ms_per_step"], "steps": d["steps"],
                "particles": d["config"]["particles_total"], "gpu_launches": d["gpu_launches"],
                "what": "strong-scaling base line of `bench.py --gpus N` for N > 1 (same code path, one slab)"}
    except Exception as exc:
        return {"error": repr(exc)}


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
    if world > 1 or args.slabs:
        import torch.distributed as dist
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        else:
            # one rank of the slab code path (strong-scaling base line: the 504^3 cloud on 1 GPU)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        from pnb200 import slabs
        args.clock_sampler_cls = ClockSampler        # rank 0 samples its GPU during the timed loop
        return slabs.bench_multi_gpu(args, rank, world, dev, METRIC, UNIT)

    n = args.lattice
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
    # the GPU's pair count (the numerator of `value`) against the CPU oracle on a sampled slab
    oracle_check = None
    if not args.no_cpu_baseline:
        from oracle import pn_oracle
        pn_oracle.set_threads(host_threads())
        cB = B.cpu().numpy()
        og = pn_oracle.Grid(3, r, mn, mx)
        og.build(cB)
        m = 250_000
        pts = np.concatenate([np.arange(m), np.arange(N // 2, N // 2 + m), np.arange(N - m, N)]).astype(np.int64)
        ref = og.count_neighbors(cB, cB, points=pts)[pts]
        got = cnt[torch.as_tensor(pts, device=dev)].cpu().numpy()     # cnt holds the counts of B
        oracle_check = {"what": f"neighbour counts of {len(pts)} points (three slabs of the update! target "
                                "cloud) against the CPU oracle", "equal": bool((ref == got).all()),
                        "pairs_in_sample": int(ref.sum())}
        assert oracle_check["equal"], "GPU pair counts differ from the oracle"
        del og, cB

    # The timed loop uses the stream-ordered forms of both calls (update_(..., blocking=False),
    # foreach_point_neighbor(..., blocking=False)): every step is only enqueued, check_(nhs) after
    # the loop synchronises and reads the error word of the whole chain.  The blocking forms (the
    # reference's semantics) are timed next to it (`blocking` in the JSON line).
    def step(s, blocking=False):
        y = coords[(s + 1) % 2]
        pn.update_(nhs, y, y, points_moving=(True, True), blocking=blocking)
        pn.foreach_point_neighbor(closure, y, y, nhs, blocking=blocking)

    for s in range(args.warmup):
        step(s, blocking=True)
    for s in range(2):
        step(s)
    pn.check_(nhs)
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
        pn.update_(nhs, y, y, points_moving=(True, True), blocking=False)
        upd_ev[s][1].record()
        pn.foreach_point_neighbor(closure, y, y, nhs, blocking=False)
        total_pairs += pairs[(s + 1) % 2]
    ev1.record()
    pn.check_(nhs)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = int(_lib.lib().pnb_launch_count()) - launches0
    prof = _lib.profile(enable=False)
    update_ms = float(np.mean([a.elapsed_time(b) for a, b in upd_ev]))
    value = total_pairs / (ms_total * 1e-3)
    # the blocking forms of both calls (the reference's semantics) for comparison
    ub_min, ub_med = timed_events(lambda: pn.update_(nhs, A, A, points_moving=(True, True)), 10)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b_steps = max(2, min(args.steps, 6))
    b0.record()
    for s in range(b_steps):
        step(s, blocking=True)
    b1.record()
    torch.cuda.synchronize()
    blocking_ms = b0.elapsed_time(b1) / b_steps

    # ---- the same step with non-zero velocities (viscosity branch active) ----------------------
    vm, _, _ = wcsph_state_torch(N, r, 3, dev, moving=True)
    closure_m = pn.WCSPHInteract(dv, vm, vm, mass, mass, pressure, pressure, smoothing_length=h,
                                 sound_speed=T(10.0), alpha=T(0.02), beta=T(0.0), delta=T(0.1))
    for s in range(2):
        pn.update_(nhs, coords[s % 2], coords[s % 2], blocking=False)
        pn.foreach_point_neighbor(closure_m, coords[s % 2], coords[s % 2], nhs)
    m0_, m1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m_steps = max(2, min(args.steps, 6))
    m0_.record()
    for s in range(m_steps):
        y = coords[(s + 1) % 2]
        pn.update_(nhs, y, y, blocking=False)
        pn.foreach_point_neighbor(closure_m, y, y, nhs)
    m1_.record()
    torch.cuda.synchronize()
    moving_ms = m0_.elapsed_time(m1_) / m_steps
    del vm, closure_m

    # ---- e2e: the same step through the library's host-buffer entry point -----------------------
    # pnb200.HostStepper (pnb_hoststep_*, include/pnb200.h): every step copies its coordinates,
    # state and pressure from pinned host memory and its dv back; three streams and double
    # buffers inside the library overlap the copies of neighbouring steps with the kernels.
    hA, hB = A.cpu().pin_memory(), B.cpu().pin_memory()
    hv, hp, hm = v.cpu().pin_memory(), pressure.cpu().pin_memory(), mass.cpu().pin_memory()
    hdv = [torch.empty((N, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    hcoords = [hA, hB]
    stepper = pn.HostStepper(nhs, N)
    # like TrixiParticles (compute_pressure! before interact!, smoothed_particle_hydrodynamics.jl:99)
    # the pressure is derived on the DEVICE from the density row of v: the host inputs of a step are
    # the coordinates and v
    stepper.set_state_equation(sound_speed=10.0, reference_density=1000.0)
    e2e_steps = max(4, min(args.steps, 50))

    def run_e2e(k):
        done = 0
        for s in range(k):
            stepper.submit(hcoords[(s + 1) % 2], hv, None, hdv[s % 2], closure, mass_host=hm if s == 0 else None)
            done += pairs[(s + 1) % 2]
        stepper.wait()
        return done

    run_e2e(3)
    torch.cuda.synchronize()
    ht0 = stepper.host_times()
    t0 = time.perf_counter()
    e2e_pairs = run_e2e(e2e_steps)
    e2e_ms = 1e3 * (time.perf_counter() - t0)      # wall clock around submit ... wait
    e2e_host = stepper.host_times(since=ht0)
    # the device result of the last step must be what the resident path computes
    e2e_same = None
    try:
        yl = coords[e2e_steps % 2]
        pn.update_(nhs, yl, yl)
        p_dev = torch.empty_like(pressure)
        pn.compute_pressure_(p_dev, v, sound_speed=10.0, reference_density=1000.0)
        closure_p = pn.WCSPHInteract(dv, v, v, mass, mass, p_dev, p_dev, smoothing_length=h,
                                     sound_speed=T(10.0), alpha=T(0.02), beta=T(0.0), delta=T(0.1))
        pn.foreach_point_neighbor(closure_p, yl, yl, nhs)
        ref_dv = dv.cpu().double()
        got_dv = hdv[(e2e_steps - 1) % 2].double()
        # same kernels, but the order of the records inside a cell (atomic arrival order of the
        # one-pass build) differs from run to run: equal up to summation order
        e2e_same = bool(((ref_dv - got_dv).norm() <= 1e-5 * ref_dv.norm()).item())
    except Exception:
        pass
    # one step alone (latency): submit + wait
    t0 = time.perf_counter()
    for s in range(3):
        stepper.submit(hcoords[s % 2], hv, None, hdv[0], closure)
        stepper.wait()
    e2e_serial_ms = 1e3 * (time.perf_counter() - t0) / 3
    del stepper
    h2d = int(hA.numel() * 4 + hv.numel() * 4)
    d2h = int(hdv[0].numel() * 4)

    # ---- roofline -----------------------------------------------------------------------------
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    sweep_ms, sweep_n = prof["k_sweep_cells"]
    sweep_avg = sweep_ms / max(sweep_n, 1)
    bytes_sweep = 56 * N + 4 * (C_cells + 1)          # WCSPH interact
    ach = bytes_sweep / (sweep_avg * 1e-3) / 1e9
    build_names = [k for k in ("k_bucket_scatter", "k_cell_hist", "k_scatter_points")
                   if prof.get(k, (0.0, 0))[1]]
    build_ms = sum(prof[k][0] for k in build_names) / args.steps
    bytes_update = 28 * N + 4 * (C_cells + 1)          # 16N + 4(C+1) + 12N (cell-ordered coordinates)
    ach_u = bytes_update / (build_ms * 1e-3) / 1e9
    ach_call = bytes_update / (update_ms * 1e-3) / 1e9
    P_avg = float(np.mean(pairs))
    K_avg = float(np.mean(kref))
    flops = 8.0 * K_avg + 70.0 * P_avg                 # non-FMA FP32 operations
    fp32_peak = SM_COUNT * FP32_LANES * sm_max_mhz * 1e6 / 1e12
    ovf_ms, _ = prof.get("k_sweep_overflow", (0.0, 0))
    prep_ms, _ = prof.get("k_flat_tiles", (0.0, 0))
    sweep_all = sweep_avg + (ovf_ms + prep_ms) / max(sweep_n, 1)
    fp32_ach = flops / (sweep_all * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic("k_sweep_flat<3,false,WcsphClT<false>,false>") if n == 254 else (None, None)
    traffic_u, traffic_u_src = ncu_traffic("k_bucket_scatter") if n == 254 else (None, None)

    cfg = workload_config(n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "scaling_note": "N = 1 is the headline configuration (BASELINE config 3, 254^3); --gpus N > 1 runs "
                        "BASELINE config 5 (504^3) in STRONG scaling, whose one-GPU base line is "
                        "configs['5_one_gpu'] of this line",
        "config": cfg,
        "workload_stats": {"cells": C_cells, "pairs_per_step": P_avg, "candidate_tests_per_step": K_avg,
                           "pairs_oracle_check": oracle_check, "git": git_head()},
        "update_ms": update_ms,
        "update_ms_blocking_call": ub_med,
        "calls": "stream-ordered (update_ / foreach_point_neighbor with blocking=False, check_ after the "
                 "loop); blocking = the same steps with the blocking calls of the reference's semantics",
        "blocking": {"ms_per_step": blocking_ms, "value": P_avg / (blocking_ms * 1e-3)},
        "interact_ms": (ms_total / args.steps) - update_ms,
        "moving": {"what": "same step with velocities ~ N(0, 0.1^2): the viscosity branch runs for "
                           "approaching pairs (two more MUFU per pair)",
                   "ms_per_step": moving_ms, "value": P_avg / (moving_ms * 1e-3)},
        "points_per_s": N / (ms_total / args.steps * 1e-3),
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": e2e_pairs / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                "api": "pnb200.HostStepper.submit/wait = pnb_hoststep_wcsph_submit / pnb_hoststep_wait "
                       "(include/pnb200.h): host pointers in, host pointer out",
                "mode": "pipelined inside the library: 3 streams, double-buffered device arrays, every "
                        "step copies its inputs (coordinates, v) from pinned host memory and its dv back; "
                        "the pressure is computed on the device from the density row of v "
                        "(compute_pressure!, as TrixiParticles does before interact!); wall clock around "
                        "the submits and the final wait",
                "serial_ms_per_step": e2e_serial_ms,
                "serial_value": P_avg / (e2e_serial_ms * 1e-3),
                "host_ms_per_submit": e2e_host,
                "matches_resident_path": e2e_same},
        "roofline": {"bound": "hbm", "kernel": "k_sweep_flat<3,false,WcsphClT<false>,false>", "achieved": ach,
                     "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "launch_ms": sweep_avg,
                     "algorithmic_bytes": bytes_sweep,
                     "note": "the fused interaction is FP32-issue bound (~100 flop/B), not HBM bound "
                             "(SURVEY.md 8d); see fp32_pipe for the binding roofline"},
        "fp32_pipe": {"achieved_tflops": fp32_ach, "peak_tflops": fp32_peak,
                      "frac": fp32_ach / fp32_peak,
                      "sweep_ms": sweep_all,
                      "model": "8*K_ref + 70*P non-FMA operations over the tile kernel + its tile-table "
                               "and overflow kernels; peak = 148 SM x 128 lanes x max clock"},
        "roofline_update": {"bound": "hbm", "kernels": build_names, "achieved": ach_u,
                            "peak": hbm_peak, "unit": "GB/s", "frac": ach_u / hbm_peak,
                            "achieved_call": ach_call, "frac_call": ach_call / hbm_peak,
                            "traffic": traffic_u, "traffic_source": traffic_u_src,
                            "device_ms": build_ms, "algorithmic_bytes": bytes_update,
                            "per_kernel_ms": {k: prof[k][0] / max(prof[k][1], 1) for k in build_names},
                            "call_ms": update_ms,
                            "call": "update_(nhs, y, y, blocking=False): stream-ordered, CUDA events around "
                                    "the call inside the step loop; the error word is read by check_ after it",
                            "layout": "buckets (one pass)" if "k_bucket_scatter" in build_names
                                      else "CSR (two passes)"},
        "kernel_ms": {k: (v_[0] / v_[1] if v_[1] else 0.0) for k, v_ in prof.items()},
    }
    del hA, hB, hv, hp, hdv
    if not args.no_sub_configs:
        try:
            line["configs"] = sub_configs(dev, hbm_peak, fp32_peak)
        except Exception as exc:      # the headline must survive a failing side measurement
            line["configs"] = {"error": repr(exc)}
    if not args.no_sub_configs and n == 254:
        # BASELINE config 5 on this one GPU: the strong-scaling base line of `bench.py --gpus N > 1`
        # (a child process: the slab path sets up torch.distributed with one rank)
        torch.cuda.empty_cache()
        line["configs"]["5_one_gpu"] = config5_one_gpu()
    if not args.no_cpu_baseline:
        cb = run_cpu_arm(n, 2, 0, budget_s=25.0)
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
    ap.add_argument("--lattice", type=int, default=254, help="lattice points per dimension (N = 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-configs", action="store_true", help="skip the config 1 / 2 / 4 sub-lines")
    ap.add_argument("--slab-mode", choices=["gather", "split"], default=None,
                    help="how the multi-GPU step hides the exchange (default: gather)")
    ap.add_argument("--slabs", action="store_true",
                    help="run the slab-decomposed config 5 path even on one GPU (strong-scaling base line)")
    ap.add_argument("--slab-lattice", type=int, default=504,
                    help="lattice of the slab-decomposed cloud (504 = BASELINE config 5)")
    ap.add_argument("--no-overlap", action="store_true", help="slab path: exchange not overlapped (A/B)")
    ap.add_argument("--slab-quick", action="store_true", help="slab path: headline only, no A/B runs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
