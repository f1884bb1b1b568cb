#!/usr/bin/env python
"""Runs update! + count on a SpatialHashingCellList (list_size = 2 N) a few times: the workload for
ncu captures of the hashed path.  usage: hashed_run.py [n=101] [reps=3]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
import pnb200 as pn
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 101
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = np.float32; dev = torch.device("cuda"); N = n ** 3; r = T(3.0) / T(n + 1)
A = bench.lattice_cloud_torch((n, n, n), n, 0, 1, dev)
nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N,
                                   cell_list=pn.SpatialHashingCellList[3](list_size=2 * N))
pn.initialize_(nhs, A, A)
cnt = torch.zeros(N, dtype=torch.int64, device=dev)
for _ in range(reps):
    pn.update_(nhs, A, A)
    pn.foreach_point_neighbor(pn.CountNeighbors(cnt), A, A, nhs)
torch.cuda.synchronize()
print("N", N, "pairs", int(cnt.sum()))
