"""Synthetic inputs: the perturbed-lattice generator of the reference's tests and benchmarks.

reference: test/point_cloud.jl:4-66 (point_cloud, perturb!), benchmarks/run_benchmarks.jl:72-89
(normalisation), :286-294 (FullGridCellList corners).  Julia's Xoshiro/randn stream cannot be
reproduced, so the PRNG is numpy's PCG64 with our own seeds; the distribution is the same.
Host-side input generation only -- nothing here is on the measured path.
"""
from __future__ import annotations

import numpy as np


def perturb_(data: np.ndarray, std_deviation: float, rng: np.random.Generator) -> np.ndarray:
    """perturb!(data, std_deviation)  (test/point_cloud.jl:60-66)."""
    data += std_deviation * rng.standard_normal(data.shape)
    return data


def point_cloud(n_points_per_dimension, search_radius, *, seed: int = 1,
                perturbation_factor_position: float = 1.0, shuffle: bool = False, sort=None):
    """point_cloud(n_points_per_dimension, search_radius; seed, ...)  (test/point_cloud.jl:4-58).

    Returns Float64 coordinates of shape (N, NDIMS) (the memory of Julia's NDIMS x N matrix):
    integer lattice 1..n_d with spacing 1 (first dimension fastest), perturbed twice with
    sigma = 0.05, then stably sorted by the cell tuple floor(x / search_radius) of the
    once-perturbed positions with dimension 1 most significant (sortperm of SVectors, :49).
    """
    if sort is None:
        sort = not shuffle
    dims = tuple(int(v) for v in n_points_per_dimension)
    nd = len(dims)
    rng = np.random.default_rng(seed)
    # CartesianIndices order: first index fastest
    grids = np.meshgrid(*[np.arange(1, n + 1, dtype=np.float64) for n in dims], indexing="ij")
    coords = np.stack([g.ravel(order="F") for g in grids], axis=1)
    coords += perturbation_factor_position * 0.05 * rng.standard_normal(coords.shape)
    cell = np.floor(coords / float(search_radius)).astype(np.int64) + 1
    perturb_(coords, perturbation_factor_position * 0.05, rng)
    if sort:
        if shuffle:
            raise ValueError("cannot sort and shuffle at the same time")
        # lexicographic, dimension 1 most significant; np.lexsort sorts by the LAST key first
        perm = np.lexsort(tuple(cell[:, d] for d in reversed(range(nd))))
        coords = coords[perm]
    elif shuffle:
        coords = coords[rng.permutation(coords.shape[0])]
    return np.ascontiguousarray(coords)


def benchmark_cloud(n_points_per_dimension, search_radius_factor=np.float32(3.0), *, seed: int = 1,
                    shuffle: bool = False):
    """The input of run_benchmark (benchmarks/run_benchmarks.jl:81-89, 286-294).

    Returns (coordinates float32 (N, NDIMS) normalised to the unit box, search_radius float32,
    min_corner, max_corner of the FullGridCellList)."""
    dims = tuple(int(v) for v in n_points_per_dimension)
    factor = np.float32(search_radius_factor)
    c64 = point_cloud(dims, float(factor), seed=seed, shuffle=shuffle)
    coords = c64.astype(np.float32)                 # convert.(typeof(search_radius_factor), ...)
    domain_size = max(dims) + 1
    coords /= np.float32(domain_size)               # coordinates ./= domain_size
    r = np.float32(factor / np.float32(domain_size))
    min_corner = np.zeros(len(dims), dtype=np.float32)
    max_corner = (np.asarray(dims, dtype=np.float64) / max(dims)).astype(np.float32)
    return np.ascontiguousarray(coords), r, min_corner, max_corner


def benchmark_cloud_torch(dims, domain_n=None, *, z0=0, seed=1, device="cuda", sort=True):
    """`benchmark_cloud` for clouds too large to generate with numpy in reasonable time: the same
    distribution (test/point_cloud.jl:4-58 + benchmarks/run_benchmarks.jl:81-89) generated with
    torch on `device`.  Lattice indices 1..dims[0] x 1..dims[1] x (z0+1)..(z0+dims[2]), sigma =
    0.05 applied twice, stably sorted by the cell tuple of the once-perturbed positions with
    dimension 1 most significant, Float32, divided by (domain_n + 1).

    Returns (coordinates float32 (N, 3) torch tensor on `device`, search_radius float32,
    min_corner, max_corner).  Input generation only -- nothing here is on the measured path."""
    import torch
    nx, ny, nz = (int(v) for v in dims)
    if domain_n is None:
        domain_n = max(nx, ny, nz)
    N = nx * ny * nz
    gen = torch.Generator(device=device).manual_seed(seed)
    k = torch.arange(N, device=device, dtype=torch.int64)
    c = torch.stack([(k % nx).to(torch.float64) + 1.0,
                     ((k // nx) % ny).to(torch.float64) + 1.0,
                     (k // (nx * ny)).to(torch.float64) + 1.0 + z0], dim=1)
    del k
    c += 0.05 * torch.randn(N, 3, device=device, dtype=torch.float64, generator=gen)
    if sort:
        cell = torch.floor(c / 3.0).to(torch.int64)
        key = (cell[:, 0] * (ny + 8) + cell[:, 1]) * (nz + z0 + 8) + cell[:, 2]
        del cell
    c += 0.05 * torch.randn(N, 3, device=device, dtype=torch.float64, generator=gen)
    if sort:
        perm = torch.sort(key, stable=True).indices
        del key
        c = c[perm]
        del perm
    out = (c.to(torch.float32) / np.float32(domain_n + 1)).contiguous()
    r = np.float32(np.float32(3.0) / np.float32(domain_n + 1))
    mn = np.zeros(3, np.float32)
    mx = (np.asarray([nx, ny, nz + z0], np.float64) / domain_n).astype(np.float32)
    return out, r, mn, mx
