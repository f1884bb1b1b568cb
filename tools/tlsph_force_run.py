#!/usr/bin/env python
"""Runs the TLSPH force sweep (pnb_tlsph_interact_f32) a few times on a periodic n^3 cloud: the
workload for ncu captures of k_tlsph_interact.  usage: tlsph_force_run.py [n=160] [reps=3]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200")); sys.path.insert(0, REPO)
import numpy as np, torch
import pnb200 as pn
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = np.float32; dev = torch.device("cuda"); N = n ** 3
s = T(1.0) / T(n + 1); r = T(3.0) / T(n + 1)
A = bench.lattice_cloud_torch((n, n, n), n, 0, 5, dev)
bmn = np.full(3, s / T(2), T); bmx = np.full(3, (T(n) + T(0.5)) * s, T)
A = torch.minimum(torch.maximum(A, torch.as_tensor(bmn + T(1e-6), device=dev)),
                  torch.as_tensor(bmx - T(1e-6), device=dev)).contiguous()
box = pn.PeriodicBox(min_corner=bmn, max_corner=bmx)
nhs = pn.GridNeighborhoodSearch[3](search_radius=r, n_points=N, periodic_box=box,
                                   cell_list=pn.FullGridCellList(min_corner=bmn, max_corner=bmx, search_radius=r))
pre = pn.PrecomputedNeighborhoodSearch[3](search_radius=r, n_points=N, periodic_box=box,
                                          update_neighborhood_search=nhs, max_neighbors=128)
pn.initialize_(pre, A, A)
xcur = (A + 0.01 * r * torch.sin(2 * np.pi * A)).contiguous()
mass = torch.full((N,), 0.1, device=dev); rho0 = torch.full((N,), 1000.0, device=dev)
Lm = (torch.eye(3, device=dev).reshape(1, 9) + 0.05 * torch.randn(N, 9, device=dev)).contiguous()
F = torch.zeros((N, 9), device=dev)
pn.foreach_point_neighbor(pn.TLSPHDeformationGradient(F, xcur, mass, rho0, Lm, smoothing_length=r / T(2), ndims_=3), A, A, pre)
pk1 = torch.zeros((N, 9), device=dev)
pn.compute_pk1_corrected_(pk1, F, Lm, young_modulus=T(1.4e6), poisson_ratio=T(0.4))
dv = torch.zeros((N, 3), device=dev)
fi = pn.TLSPHInteract(dv, xcur, mass, rho0, pk1, F, smoothing_length=r / T(2), young_modulus=T(1.4e6), penalty_alpha=T(0.1))
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pn.foreach_point_neighbor(fi, A, A, pre); e1.record(); torch.cuda.synchronize()
    print(f"N={N} pairs={pre._lists.n_pairs} force sweep {e0.elapsed_time(e1):.3f} ms", flush=True)
