"""Host-side logic of the multi-GPU slab decomposition, world_size 2 and 3 over gloo on CPU:
ownership, migration and ghost selection of pnb200.slabs.SlabExchange against the global truth."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nd, result_dir):
    import sys
    sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pnb200
        from pnb200.slabs import SlabExchange, split_layers
        T = np.float32
        dims = (12, 10, 40)[-nd:] if nd < 3 else (12, 10, 40)
        c, r, mn, mx = pnb200.benchmark_cloud(dims, seed=5)
        N = len(c)
        ex = SlabExchange(nd, r, mn, mx, rank, world)
        assert ex.layers == split_layers(ex.grid_size[-1], world)
        assert ex.layers[0][0] == 2 and ex.layers[-1][1] == ex.grid_size[-1] - 1
        ids = np.arange(N, dtype=np.float32)
        rows0 = torch.from_numpy(np.concatenate([c, ids[:, None], (ids * 2)[:, None]], axis=1))
        own0 = rows0[ex.owned_mask(rows0[:, :nd])]
        # move every point by up to ~0.4 cells: some cross slab boundaries
        rng = np.random.default_rng(9)
        moved = (c + (0.4 * r) * rng.uniform(-1, 1, c.shape).astype(T)).astype(T)
        moved = np.clip(moved, mn, mx).astype(T)
        rows1 = own0.clone()
        rows1[:, :nd] = torch.from_numpy(moved)[own0[:, nd].to(torch.int64)]
        local, n_own = ex.exchange(rows1)
        # global truth
        cz = ex.cell_layer(torch.from_numpy(moved)).numpy()
        exp_own = set(np.nonzero((cz >= ex.z_lo) & (cz <= ex.z_hi))[0].tolist())
        exp_ghost = set()
        if rank > 0:
            exp_ghost |= set(np.nonzero(cz == ex.z_lo - 1)[0].tolist())
        if rank + 1 < world:
            exp_ghost |= set(np.nonzero(cz == ex.z_hi + 1)[0].tolist())
        got_ids = local[:, nd].to(torch.int64).numpy()
        assert set(got_ids[:n_own].tolist()) == exp_own
        assert set(got_ids[n_own:].tolist()) == exp_ghost
        assert len(got_ids) == len(set(got_ids.tolist()))
        # payload columns travel with the points, coordinates are the moved ones
        assert np.array_equal(local[:, nd + 1].numpy(), 2 * local[:, nd].numpy())
        assert np.array_equal(local[:, :nd].numpy(), moved[got_ids])
        # every point is owned by exactly one rank
        counts = torch.zeros(N, dtype=torch.int64)
        counts[torch.from_numpy(got_ids[:n_own])] = 1
        dist.all_reduce(counts)
        assert bool((counts == 1).all())
        # the window covers the owned layers plus ghost and padding layers
        lo, hi = ex.window
        assert lo[-1] == max(1, ex.z_lo - 2) and hi[-1] == min(ex.grid_size[-1], ex.z_hi + 2)
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nd", [(2, 3), (3, 3), (2, 2)])
def test_slab_exchange_gloo(tmp_path, world, nd):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nd, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert (tmp_path / f"ok{rank}").exists()
