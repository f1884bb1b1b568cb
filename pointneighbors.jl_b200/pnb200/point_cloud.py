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
