#!/usr/bin/env python
"""Per-kernel device times of the hot path for one lattice size (development aid).
usage: python tools/kernel_times.py [--lattice 254] [--reps 5] [--closures count,nbody,wcsph,nlist]"""
import argparse
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
ap.add_argument("--lattice", type=int, default=254)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--closures", default="count,nbody,wcsph")
ap.add_argument("--moving", action="store_true", help="non-zero velocities (viscosity branch active)")
ap.add_argument("--wpc", type=int, default=0, help="warps per cell override (2 or 4)")
ap.add_argument("--build", type=int, default=25, help="build kernel variant bits")
ap.add_argument("--half", type=int, default=-1, help="0 = exact Float32 test instead of the fp16 pre-filter")
ap.add_argument("--flat", type=int, default=1, help="1 = k_sweep_flat (default), 0 = k_sweep_tiles of round 1")
args = ap.parse_args()
n = args.lattice
T = np.float32
dev = torch.device("cuda")
N = n ** 3
r = T(3.0) / T(n + 1)
A = bench.lattice_cloud_torch((n, n, n), n, 0, 1, dev)
nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, cell_list=pn.FullGridCellList(
    min_corner=np.zeros(3, T), max_corner=np.ones(3, T), search_radius=r))
v, mass, pressure = bench.wcsph_state_torch(N, r, 3, dev)
if args.moving:
    v[:, :3] = 0.1 * torch.randn(N, 3, device=dev)
dv = torch.zeros((N, 4), device=dev)
cnt = torch.zeros(N, dtype=torch.int64, device=dev)
dv3 = torch.zeros((N, 3), device=dev)
m_nb = (1e10 * (torch.rand(N, device=dev) + 1)).to(torch.float32)
cl = {
    "count": pn.CountNeighbors(cnt),
    "nbody": pn.NBodyGravity(dv3, m_nb, T(6.6743e-11)),
    "wcsph": pn.WCSPHInteract(dv, v, v, mass, mass, pressure, pressure, smoothing_length=r / T(2),
                              sound_speed=T(10.0)),
}
_lib.lib().pnb_set_tuning(args.wpc, args.half)
_lib.lib().pnb_set_sweep_kernel(args.flat)
_lib.lib().pnb_set_build_tuning(args.build)
pn.initialize_(nhs, A, A)
pn.foreach_point_neighbor(cl["count"], A, A, nhs)
P = int(cnt.sum())
print(f"N={N} cells={nhs.total_cells()} pairs={P} ({P / N:.1f}/pt)")
for name in args.closures.split(","):
    if name == "nlist":
        pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=N,
                                                  update_neighborhood_search=nhs, max_neighbors=128)
        _lib.profile(enable=True, reset=True); _lib.profile(reset=True)
        for _ in range(args.reps):
            pn.initialize_(pre, A, A)
        prof = _lib.profile(enable=False)
    else:
        for _ in range(2):
            pn.update_(nhs, A, A)
            pn.foreach_point_neighbor(cl[name], A, A, nhs)
        _lib.profile(enable=True, reset=True); _lib.profile(reset=True)
        for _ in range(args.reps):
            pn.update_(nhs, A, A)
            pn.foreach_point_neighbor(cl[name], A, A, nhs)
        prof = _lib.profile(enable=False)
    line = "  ".join(f"{k}={ms / cnt_:.3f}ms" for k, (ms, cnt_) in prof.items() if cnt_)
    sw = prof["k_sweep_cells"]
    print(f"{name:6s}: {line}   sweep {P / (sw[0] / sw[1] * 1e-3) / 1e9:.1f} Gpairs/s")
