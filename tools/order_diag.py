#!/usr/bin/env python
"""update! device time as a function of the ORDER of the points (same cloud): the reference
generator's order (cells sorted with dim 1 most significant), linear cell order (dim 1 fastest =
the order of the bucket memory) and a random shuffle."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
import pnb200 as pn
from pnb200 import _lib
import bench
n = 254
T = np.float32; dev = torch.device("cuda"); N = n ** 3; r = T(3.0) / T(n + 1)
A = bench.lattice_cloud_torch((n, n, n), n, 0, 1, dev)
nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, cell_list=pn.FullGridCellList(
    min_corner=np.zeros(3, T), max_corner=np.ones(3, T), search_radius=r))
pn.initialize_(nhs, A, A)
cells = nhs.point_cells(A).long()
lin_order = A[torch.argsort(cells, stable=True)].contiguous()
v, mass, pressure = bench.wcsph_state_torch(N, r, 3, dev)
dv = torch.zeros((N, 4), device=dev)
f = pn.WCSPHInteract(dv, v, v, mass, mass, pressure, pressure, smoothing_length=r / T(2), sound_speed=T(10.0))
for name, X in (("generator order", A), ("linear cell order", lin_order)):
    pn.initialize_(nhs, X, X)
    for _ in range(3):
        pn.update_(nhs, X, X); pn.foreach_point_neighbor(f, X, X, nhs)
    _lib.profile(enable=True, reset=True); _lib.profile(reset=True)
    for _ in range(10):
        pn.update_(nhs, X, X); pn.foreach_point_neighbor(f, X, X, nhs)
    prof = _lib.profile(enable=False)
    print(f"{name:18s}: " + "  ".join(f"{k}={ms / c:.4f}" for k, (ms, c) in prof.items() if c), flush=True)
