"""pnb200 -- B200-native fixed-radius neighbourhood search behind the PointNeighbors.jl API.

The compute path is libpnb200.so (hand-written CUDA for sm_100a, C ABI in include/pnb200.h);
this package is the thin host-side mirror of the reference's Julia interface for that path.
"""
from . import _build, _lib
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all
from .point_cloud import point_cloud, perturb_, benchmark_cloud, benchmark_cloud_torch

__all__ = list(_api_all) + ["point_cloud", "perturb_", "benchmark_cloud", "benchmark_cloud_torch", "build", "library_path"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libpnb200.so in-tree for sm_100a."""
    return _build.build(force=force, verbose=verbose)


def library_path() -> str:
    return _lib.library_path()
