#!/usr/bin/env python
"""Times BASELINE.json's single-GPU configurations 1-4 on the device (CUDA events, L2 flushed
between repetitions for the small clouds) and prints one JSON object per configuration.

  C1 count_neighbors, 64^3 (262k points)
  C2 n-body gravity, 101^3 (1.03M points)
  C3 WCSPH step 254^3 (bench.py is the contract version of this one)
  C4 PrecomputedNeighborhoodSearch build (sorted lists) + TLSPH deformation gradient, 200^3
     (8M points) with a PeriodicBox

usage: python tools/config_times.py [--configs 1,2,3,4] [--reps 10]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200"))
sys.path.insert(0, REPO)
import numpy as np
import torch
import pnb200 as pn
from pnb200 import _lib
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1,2,3,4")
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
T = np.float32
dev = torch.device("cuda")
HBM = bench.measured_peaks()[0]
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    times = []
    for _ in range(reps):
        if flush:
            flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.min(times)), float(np.median(times))


def lattice(n):
    N = n ** 3
    r = T(3.0) / T(n + 1)
    A = bench.lattice_cloud_torch((n, n, n), n, 0, 1, dev)
    nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, cell_list=pn.FullGridCellList(
        min_corner=np.zeros(3, T), max_corner=np.ones(3, T), search_radius=r))
    pn.initialize_(nhs, A, A)
    return N, r, A, nhs


def pairs_of(nhs, A):
    cnt = torch.zeros(A.shape[0], dtype=torch.int64, device=dev)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), A, A, nhs)
    return int(cnt.sum())


out = []
for c in [int(v) for v in args.configs.split(",")]:
    if c == 1:
        N, r, A, nhs = lattice(64)
        cnt = torch.zeros(N, dtype=torch.int64, device=dev)
        f = pn.CountNeighbors(cnt)
        mn, med = timed(lambda: pn.foreach_point_neighbor(f, A, A, nhs), args.reps, True)
        P = int(cnt.sum())
        umn, umed = timed(lambda: pn.update_(nhs, A, A), args.reps, True)
        out.append({"config": 1, "what": "count_neighbors 64^3", "N": N, "pairs": P,
                    "sweep_ms_min": mn, "sweep_ms_median": med, "gpairs_per_s": P / mn / 1e6,
                    "update_ms_min": umn, "cells": nhs.total_cells()})
    elif c == 2:
        N, r, A, nhs = lattice(101)
        P = pairs_of(nhs, A)
        dv = torch.zeros((N, 3), device=dev)
        mass = (1e10 * (torch.rand(N, device=dev) + 1)).to(torch.float32)
        f = pn.NBodyGravity(dv, mass, T(6.6743e-11))
        mn, med = timed(lambda: pn.foreach_point_neighbor(f, A, A, nhs), args.reps, True)
        umn, umed = timed(lambda: pn.update_(nhs, A, A), args.reps, True)
        out.append({"config": 2, "what": "n-body 101^3", "N": N, "pairs": P, "sweep_ms_min": mn,
                    "sweep_ms_median": med, "gpairs_per_s": P / mn / 1e6, "update_ms_min": umn})
    elif c == 3:
        N, r, A, nhs = lattice(254)
        P = pairs_of(nhs, A)
        v, mass, pressure = bench.wcsph_state_torch(N, r, 3, dev)
        dv = torch.zeros((N, 4), device=dev)
        f = pn.WCSPHInteract(dv, v, v, mass, mass, pressure, pressure, smoothing_length=r / T(2),
                             sound_speed=T(10.0))
        mn, med = timed(lambda: pn.foreach_point_neighbor(f, A, A, nhs), args.reps, False)
        umn, umed = timed(lambda: pn.update_(nhs, A, A), args.reps, False)
        C = nhs.total_cells()
        out.append({"config": 3, "what": "WCSPH 254^3", "N": N, "pairs": P, "interact_ms_min": mn,
                    "interact_ms_median": med, "gpairs_per_s": P / mn / 1e6,
                    "update_ms_min": umn, "update_ms_median": umed,
                    "update_hbm_frac": (28 * N + 4 * (C + 1)) / (umn * 1e-3) / 1e9 / HBM})
        del v, dv, pressure
    elif c == 4:
        n = 200
        N = n ** 3
        s = T(1.0) / T(n + 1)
        r = T(3.0) / T(n + 1)
        A = bench.lattice_cloud_torch((n, n, n), n, 0, 5, dev)
        bmn = np.full(3, s / T(2), T)
        bmx = np.full(3, (T(n) + T(0.5)) * s, T)
        A = torch.minimum(torch.maximum(A, torch.as_tensor(bmn + T(1e-6), device=dev)),
                          torch.as_tensor(bmx - T(1e-6), device=dev)).contiguous()
        box = pn.PeriodicBox(min_corner=bmn, max_corner=bmx)
        nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, periodic_box=box,
                                           cell_list=pn.FullGridCellList(min_corner=bmn, max_corner=bmx,
                                                                         search_radius=r))
        pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=N, periodic_box=box,
                                                  update_neighborhood_search=nhs, max_neighbors=128,
                                                  transpose_backend=True)
        _lib.profile(enable=True, reset=True); _lib.profile(reset=True)
        bmin, bmed = timed(lambda: pn.initialize_(pre, A, A), max(3, args.reps // 2), False)
        prof = _lib.profile(enable=False)
        off, ids = pre.export_csr()
        P = int(ids.numel())
        del off, ids
        xcur = (A + 0.01 * r * torch.sin(2 * np.pi * A)).contiguous()
        mass = torch.full((N,), 0.1, device=dev)
        rho0 = torch.full((N,), 1000.0, device=dev)
        Lm = (torch.eye(3, device=dev).reshape(1, 9) + 0.05 * torch.randn(N, 9, device=dev)).contiguous()
        F = torch.zeros((N, 9), device=dev)
        f = pn.TLSPHDeformationGradient(F, xcur, mass, rho0, Lm, smoothing_length=r / T(2), ndims_=3)
        smin, smed = timed(lambda: pn.foreach_point_neighbor(f, A, A, pre), args.reps, False)
        bytes_sweep = 108 * N + 4 * P + 4
        # the other TLSPH kernel: PK1 stress + penalty forces over the same lists (8f rank 2)
        pk1 = torch.zeros((N, 9), device=dev)
        pmin, pmed = timed(lambda: pn.compute_pk1_corrected_(pk1, F, Lm, young_modulus=T(1.4e6),
                                                            poisson_ratio=T(0.4)), args.reps, False)
        dvs = torch.zeros((N, 3), device=dev)
        fi = pn.TLSPHInteract(dvs, xcur, mass, rho0, pk1, F, smoothing_length=r / T(2),
                              young_modulus=T(1.4e6), penalty_alpha=T(0.1))
        imin, imed = timed(lambda: pn.foreach_point_neighbor(fi, A, A, pre), args.reps, False)
        bytes_force = 4 * P + 8 * N + 104 * N + 12 * N
        out.append({"config": 4, "what": "Precomputed (periodic, sorted) + TLSPH F 200^3", "N": N,
                    "tlsph_force_ms_min": imin, "tlsph_force_ms_median": imed,
                    "tlsph_force_hbm_gbs": bytes_force / (imin * 1e-3) / 1e9,
                    "tlsph_force_hbm_frac": bytes_force / (imin * 1e-3) / 1e9 / HBM,
                    "pk1_corrected_ms_min": pmin,
                    "pk1_corrected_hbm_frac": 108 * N / (pmin * 1e-3) / 1e9 / HBM,
                    "pairs": P, "n_cells": list(nhs.n_cells), "list_build_ms_min": bmin,
                    "list_build_ms_median": bmed, "tlsph_ms_min": smin, "tlsph_ms_median": smed,
                    "tlsph_hbm_gbs": bytes_sweep / (smin * 1e-3) / 1e9,
                    "tlsph_hbm_frac": bytes_sweep / (smin * 1e-3) / 1e9 / HBM,
                    "list_build_kernels_ms": {k: ms / max(prof["k_sort_lists"][1], 1)
                                              for k, (ms, cnt_) in prof.items() if cnt_}})
    elif c == 6:
        # two point sets (fluid-boundary style): x = a dense slab of the cloud
        # (16 cell layers, jittered by a third of a spacing), y = the whole 254^3 cloud
        N, r, A, nhs = lattice(254)
        sel = (A[:, 2] > 0.3) & (A[:, 2] < 0.3 + 16 * float(r))
        X = (A[sel] + (T(1.0) / T(3.0)) * (r / T(3)) * torch.randn(int(sel.sum()), 3, device=dev)).clamp_(0.0, 1.0).contiguous()
        nxp = X.shape[0]
        v, mass, pressure = bench.wcsph_state_torch(N, r, 3, dev)
        vx, mx_, px_ = v[sel].contiguous(), mass[sel].contiguous(), pressure[sel].contiguous()
        res = {"config": 6, "what": "two sets: x = dense slab (16 cell layers), y = 254^3", "Nx": nxp, "N": N}
        for mode, name in ((1, "tiles"), (0, "per_point")):
            _lib.lib().pnb_set_twoset_tiles(mode)
            cnt = torch.zeros(nxp, dtype=torch.int64, device=dev)
            f = pn.CountNeighbors(cnt)
            mn, med = timed(lambda: pn.foreach_point_neighbor(f, X, A, nhs), args.reps, False)
            P = int(cnt.sum())
            dv = torch.zeros((nxp, 4), device=dev)
            fw = pn.WCSPHInteract(dv, vx, v, mx_, mass, px_, pressure, smoothing_length=r / T(2),
                                  sound_speed=T(10.0))
            wmn, wmed = timed(lambda: pn.foreach_point_neighbor(fw, X, A, nhs), args.reps, False)
            res.update({"pairs": P, f"count_ms_{name}": mn, f"count_gpairs_per_s_{name}": P / mn / 1e6,
                        f"wcsph_ms_{name}": wmn, f"wcsph_gpairs_per_s_{name}": P / wmn / 1e6})
        _lib.lib().pnb_set_twoset_tiles(1)
        out.append(res)
    elif c == 7:
        # SpatialHashingCellList (8f rank 3) against the FullGridCellList on the same cloud
        res = {"config": 7, "what": "SpatialHashingCellList(list_size = 2 N) vs FullGridCellList"}
        for n in (64, 101):
            N, r, A, full = lattice(n)
            hashed = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N,
                                                  cell_list=pn.SpatialHashingCellList[3](list_size=2 * N))
            pn.initialize_(hashed, A, A)
            cnt = torch.zeros(N, dtype=torch.int64, device=dev)
            f = pn.CountNeighbors(cnt)
            hmn, _ = timed(lambda: pn.foreach_point_neighbor(f, A, A, hashed), args.reps, True)
            P = int(cnt.sum())
            fmn, _ = timed(lambda: pn.foreach_point_neighbor(f, A, A, full), args.reps, True)
            assert int(cnt.sum()) == P
            umn, _ = timed(lambda: pn.update_(hashed, A, A), args.reps, True)
            _, coll = hashed.export_hash_table()
            res[f"n{n}"] = {"N": N, "pairs": P, "count_ms_hashed": hmn, "count_ms_full_grid": fmn,
                            "gpairs_per_s_hashed": P / hmn / 1e6, "update_ms_hashed": umn,
                            "colliding_keys": int(coll.sum())}
        out.append(res)
    torch.cuda.empty_cache()
for o in out:
    print(json.dumps(o))
