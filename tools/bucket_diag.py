#!/usr/bin/env python
"""Where does k_bucket_scatter spend its time?  DIAG variants (results invalid, timing only):
0 full kernel, 1 non-returning atomics, 2 no stores, 3 = 1 + 2, 4 no atomics, 6 loads + cell arithmetic only."""
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
pn.update_(nhs, A, A)
for diag in (0, 1, 2, 3, 4, 6, 0):
    _lib.lib().pnb_set_build_tuning(25 | (diag << 8))
    for _ in range(3):
        pn.update_(nhs, A, A)
    _lib.profile(enable=True, reset=True); _lib.profile(reset=True)
    for _ in range(20):
        pn.update_(nhs, A, A)
    prof = _lib.profile(enable=False)
    print(f"diag={diag}: " + "  ".join(f"{k}={ms / c:.4f}" for k, (ms, c) in prof.items() if c), flush=True)
_lib.lib().pnb_set_build_tuning(25)
