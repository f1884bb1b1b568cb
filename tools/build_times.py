#!/usr/bin/env python
"""Device times of the update! kernels for the build-kernel variants (development aid)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
import pnb200 as pn
from pnb200 import _lib
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 254
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 6]
T = np.float32; dev = torch.device("cuda"); N = n ** 3; r = T(3.0) / T(n + 1)
A = bench.lattice_cloud_torch((n, n, n), n, 0, 1, dev)
R = A[torch.randperm(N, device=dev)].contiguous()          # same cloud, random point order
nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, cell_list=pn.FullGridCellList(
    min_corner=np.zeros(3, T), max_corner=np.ones(3, T), search_radius=r))
for name, X in (("cell-sorted", A), ("shuffled", R)):
    for v in variants:
        _lib.lib().pnb_set_build_tuning(v)
        for _ in range(3):
            pn.update_(nhs, X, X)
        _lib.profile(enable=True, reset=True); _lib.profile(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            pn.update_(nhs, X, X)
        ev1.record(); torch.cuda.synchronize()
        prof = _lib.profile(enable=False)
        line = "  ".join(f"{k}={ms / c:.4f}" for k, (ms, c) in prof.items() if c)
        cs, _ = nhs.export_csr()          # next update! of this loop starts from a CSR build again
        print(f"{name:12s} variant={v}: {line}  call={ev0.elapsed_time(ev1) / 10:.4f} ms")
