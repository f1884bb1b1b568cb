#!/usr/bin/env python
"""k_bucket_scatter: match.any lane groups (default) vs. runs of adjacent lanes (tuning bit 2048, the version before).
Checks the CSR export of both against the oracle on a small cloud, then times config 3."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
import pnb200 as pn
from pnb200 import _lib
from oracle import pn_oracle as oracle
import bench
T = np.float32; dev = torch.device("cuda")
L = _lib.lib()
# 25 default (match.any lane groups); +2048 runs of adjacent lanes
VARIANTS = (25, 25 | 2048)
# parity on a small cloud (sorted and shuffled), update! path = buckets
c, r, mn, mx = pn.benchmark_cloud((33, 31, 29), seed=12)
rng = np.random.default_rng(1)
for cloud in (c, c[rng.permutation(len(c))]):
    og = oracle.Grid(3, r, mn, mx); og.build(cloud)
    x = torch.from_numpy(np.ascontiguousarray(cloud)).to(dev)
    for variant in VARIANTS:
        L.pnb_set_build_tuning(variant)
        nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=len(cloud), cell_list=pn.FullGridCellList(
            min_corner=mn, max_corner=mx, search_radius=r))
        pn.initialize_(nhs, x, x)
        pn.update_(nhs, x, x)
        pn.update_(nhs, x, x)
        cs, cp = nhs.export_csr()
        ok = (cs.cpu().numpy() == og.cell_start).all() and (cp.cpu().numpy() == og.cell_points).all()
        print(f"variant {variant}: parity {'OK' if ok else 'FAILED'}", flush=True)
n = 254
N = n ** 3; r = T(3.0) / T(n + 1)
A = bench.lattice_cloud_torch((n, n, n), n, 0, 1, dev)
nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, cell_list=pn.FullGridCellList(
    min_corner=np.zeros(3, T), max_corner=np.ones(3, T), search_radius=r))
pn.initialize_(nhs, A, A)
pn.update_(nhs, A, A)
for variant in VARIANTS + VARIANTS:
    L.pnb_set_build_tuning(variant)
    for _ in range(3):
        pn.update_(nhs, A, A)
    _lib.profile(enable=True, reset=True); _lib.profile(reset=True)
    for _ in range(30):
        pn.update_(nhs, A, A)
    prof = _lib.profile(enable=False)
    print(f"variant={variant}: " + "  ".join(f"{k}={ms / c:.4f}" for k, (ms, c) in prof.items() if c), flush=True)
L.pnb_set_build_tuning(25)
