"""Multi-GPU slab decomposition of the grid search: one process per GPU, one ghost cell layer
exchanged per step -- by kernels that store into the neighbouring GPU's memory (csrc/link.cu, the
overlapped step below) or over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

There is no counterpart in the reference (single device).  Design:

  * the GLOBAL FullGridCellList is cut along its slowest cell dimension (the last one): linear
    cell index = c1 + (c2-1) n1 + (c3-1) n1 n2 (src/cell_lists/full_grid.jl:157-161), so a slab
    of cell layers is a contiguous range of cells;
  * rank g owns the cell layers [z_lo, z_hi]; because cell_size >= search_radius a point only
    interacts with the 3^d stencil (src/nhs_grid.jl:577-583), so the rank needs exactly the
    points of the layers z_lo - 1 and z_hi + 1 of its neighbours (ghosts);
  * every rank builds a WINDOW of the global grid (pnb_grid_create_window_f32) with the global
    cell arithmetic, so cell assignments and neighbour sets are bit-identical to the
    undecomposed search; results are taken for the owned points only;
  * one exchange round per step and direction carries both the migrants (points whose cell left
    the slab) and the boundary layer (ghosts): a rank sends upwards every owned point with
    cz >= z_hi, the receiver keeps cz >= its z_lo as its own and cz == z_lo - 1 as ghosts, the
    sender keeps what it sent with cz == z_hi + 1 as its own ghosts (symmetrically downwards).

The exchange logic is plain torch + torch.distributed and runs on CPU tensors with gloo
(tests/test_slabs_gloo.py); the compute (build + sweep) is the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import ArgumentError, check


def global_grid(ndims: int, search_radius, min_corner, max_corner):
    """(padded_min, grid_size) of the global FullGridCellList, from the library's host arithmetic."""
    mn = np.ascontiguousarray(min_corner, dtype=np.float32)
    mx = np.ascontiguousarray(max_corner, dtype=np.float32)
    pmin = (C.c_float * 3)()
    gs = (C.c_int64 * 3)()
    check(_lib.lib().pnb_grid_params_f32(ndims, np.float32(search_radius),
                                         mn.ctypes.data_as(_lib._pf), mx.ctypes.data_as(_lib._pf),
                                         None, None, pmin, None, gs, None, None))
    return np.array(pmin[:ndims], dtype=np.float32), tuple(int(v) for v in gs[:ndims])


def split_layers(n_layers_total: int, world: int) -> List[Tuple[int, int]]:
    """Owned layer ranges [z_lo, z_hi] (global 1-based cell coordinates) of every rank: the valid
    layers 2 .. n-1 in `world` contiguous, nearly equal parts."""
    valid = n_layers_total - 2
    if valid < 3 * world:
        raise ArgumentError(f"{valid} cell layers cannot be split into {world} slabs of >= 3 layers")
    base, extra = divmod(valid, world)
    out, z = [], 2
    for g in range(world):
        n = base + (1 if g < extra else 0)
        out.append((z, z + n - 1))
        z += n
    return out


class SlabExchange:
    """Ownership, migration and ghost exchange for one rank.  Device-agnostic (CPU/gloo or CUDA/NCCL)."""

    def __init__(self, ndims: int, search_radius, min_corner, max_corner, rank: int, world: int,
                 group=None):
        self.ndims = int(ndims)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.search_radius = np.float32(search_radius)
        self.padded_min, self.grid_size = global_grid(ndims, search_radius, min_corner, max_corner)
        self.layers = split_layers(self.grid_size[-1], world)
        self.z_lo, self.z_hi = self.layers[rank]
        # window of the global grid held by this rank: owned layers, one ghost layer and one
        # (empty) padding layer on each side, clipped to the global grid
        lo = [1] * ndims
        hi = list(self.grid_size)
        lo[-1] = max(1, self.z_lo - 2)
        hi[-1] = min(self.grid_size[-1], self.z_hi + 2)
        self.window = (tuple(lo), tuple(hi))
        self.last_stats = {}

    # cell layer of every point: floor((z - min_corner) / cell_size) + 1 in Float32
    # (src/cell_lists/full_grid.jl:93), the same IEEE operations as the library's cell kernel
    def cell_layer(self, coords):
        import torch
        z = coords[:, self.ndims - 1]
        # tensor / tensor: torch turns `tensor / python_scalar` into a multiplication by the
        # reciprocal on CUDA, which would not be the reference's true division
        pmin = torch.tensor(float(self.padded_min[-1]), dtype=torch.float32, device=coords.device)
        cs = torch.tensor(float(self.search_radius), dtype=torch.float32, device=coords.device)
        q = (z - pmin) / cs
        return torch.floor(q).to(torch.int64) + 1

    def owned_mask(self, coords):
        cz = self.cell_layer(coords)
        return (cz >= self.z_lo) & (cz <= self.z_hi)

    def _sendrecv(self, send_up, send_down):
        """Exchange variable-length row blocks with rank+1 (up) and rank-1 (down)."""
        import torch
        import torch.distributed as dist
        dev = send_up.device
        ncol = send_up.shape[1]
        up, down = self.rank + 1, self.rank - 1
        has_up, has_down = up < self.world, down >= 0
        n_send = torch.tensor([send_up.shape[0], send_down.shape[0]], dtype=torch.int64, device=dev)
        n_recv = torch.zeros(2, dtype=torch.int64, device=dev)   # [from up, from down]
        ops = []
        if has_up:
            ops += [dist.P2POp(dist.isend, n_send[0:1], up, self.group),
                    dist.P2POp(dist.irecv, n_recv[0:1], up, self.group)]
        if has_down:
            ops += [dist.P2POp(dist.isend, n_send[1:2], down, self.group),
                    dist.P2POp(dist.irecv, n_recv[1:2], down, self.group)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        n_up, n_down = (int(v) for v in n_recv.tolist())
        recv_up = torch.empty((n_up, ncol), dtype=send_up.dtype, device=dev)
        recv_down = torch.empty((n_down, ncol), dtype=send_up.dtype, device=dev)
        ops = []
        if has_up:
            if send_up.shape[0]:
                ops.append(dist.P2POp(dist.isend, send_up, up, self.group))
            if n_up:
                ops.append(dist.P2POp(dist.irecv, recv_up, up, self.group))
        if has_down:
            if send_down.shape[0]:
                ops.append(dist.P2POp(dist.isend, send_down, down, self.group))
            if n_down:
                ops.append(dist.P2POp(dist.irecv, recv_down, down, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recv_up, recv_down

    def exchange(self, rows):
        """rows: (n, ndims + k) float32, the first ndims columns are the coordinates of the points
        this rank currently holds as its own.  Returns (local_rows, n_own): the new owned points
        first, then the ghosts.  Convenience form (copies everything); step loops use
        exchange_inplace."""
        import torch
        n = rows.shape[0]
        nd = self.ndims
        cap = n + n // 8 + 1024
        coords = torch.empty((cap, nd), dtype=rows.dtype, device=rows.device)
        state = torch.empty((cap, rows.shape[1] - nd), dtype=rows.dtype, device=rows.device)
        coords[:n] = rows[:, :nd]
        state[:n] = rows[:, nd:]
        (coords, state), n_own, n_local = self.exchange_inplace([coords, state], n)
        return torch.cat([coords[:n_local], state[:n_local]], dim=1).contiguous(), n_own

    def _classify_cuda(self, coords, n, has_up, has_down):
        """up / down / leave index lists (ascending int64) from pnb_slab_classify_f32."""
        import torch
        dev = coords.device
        cap = getattr(self, "_idx_cap", 0)
        while True:
            if cap == 0:
                cap = max(n // 8, 1 << 16)
            if getattr(self, "_idx_bufs", None) is None or self._idx_bufs[0].numel() < cap:
                self._idx_bufs = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3)]
                self._idx_counts = torch.zeros(4, dtype=torch.int32, device=dev)
                self._idx_cap = cap
            counts = (C.c_int64 * 3)()
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            check(_lib.lib().pnb_slab_classify_f32(
                coords.data_ptr(), n, self.ndims, np.float32(self.padded_min[-1]),
                np.float32(self.search_radius), self.z_lo, self.z_hi, int(has_up), int(has_down),
                self._idx_bufs[0].data_ptr(), self._idx_bufs[1].data_ptr(),
                self._idx_bufs[2].data_ptr(), cap, self._idx_counts.data_ptr(), counts, stream))
            if max(counts) <= cap:
                break
            cap = int(max(counts)) + int(max(counts)) // 4 + 1024      # lists were too short: retry
            self._idx_bufs = None
        # ascending order like torch.nonzero (the lists come out of the kernel unordered)
        return tuple(torch.sort(b[:int(c)]).values.to(torch.int64)
                     for b, c in zip(self._idx_bufs, counts))

    def _exchange_inplace_cuda(self, arrays, n):
        """exchange_inplace on the device: two library calls (pnb_slab_pack_f32 before and
        pnb_slab_unpack_f32 after the NCCL exchange) instead of ~100 torch operations per step.
        The order of the points inside the owned / ghost ranges is not deterministic."""
        import torch
        L = _lib.lib()
        coords = arrays[0]
        dev = coords.device
        has_up, has_down = self.rank + 1 < self.world, self.rank > 0
        widths = [1 if a.ndim == 1 else a.shape[1] for a in arrays]
        W = sum(widths)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        pmin, cs = np.float32(self.padded_min[-1]), np.float32(self.search_radius)

        def table(arrs):
            t = _lib.SlabArrays()
            for k, (a, w) in enumerate(zip(arrs, widths)):
                t.ptr[k] = a.data_ptr()
                t.width[k] = w
            t.n_arrays = len(arrs)
            return t

        cap = getattr(self, "_x_cap", 0) or max(n // 8, 1 << 16)
        while True:
            if getattr(self, "_x_bufs", None) is None or self._x_cap < cap:
                self._x_idx = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3)]
                self._x_cnt = torch.zeros(4, dtype=torch.int32, device=dev)
                self._x_bufs = [torch.empty((cap, W), dtype=torch.float32, device=dev) for _ in range(2)]
                self._x_cap = cap
            counts = (C.c_int64 * 3)()
            tab = table(arrays)
            check(L.pnb_slab_pack_f32(C.byref(tab), n, self.ndims, pmin, cs, self.z_lo, self.z_hi,
                                      int(has_up), int(has_down), self._x_idx[0].data_ptr(),
                                      self._x_idx[1].data_ptr(), self._x_idx[2].data_ptr(),
                                      self._x_cap, self._x_bufs[0].data_ptr(),
                                      self._x_bufs[1].data_ptr(), self._x_cnt.data_ptr(), counts,
                                      stream))
            if max(counts) <= self._x_cap:
                break
            cap = int(max(counts)) + int(max(counts)) // 4 + 1024
        n_up, n_down, n_leave = (int(c) for c in counts)
        send_up, send_down = self._x_bufs[0][:n_up], self._x_bufs[1][:n_down]
        recv_up, recv_down = self._sendrecv(send_up, send_down)
        n_ru, n_rd = recv_up.shape[0], recv_down.shape[0]
        n_stay = n - n_leave
        need = n_stay + n_ru + n_rd + n_up + n_down
        if need > coords.shape[0]:
            new_cap = need + need // 8 + 1024
            grown = []
            for a in arrays:
                g = torch.empty((new_cap,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev)
                g[:n] = a[:n]
                grown.append(g)
            arrays = grown
        n_scratch = 16 + 3 * n_leave + n_ru + n_rd + n_up + n_down + 16
        if getattr(self, "_x_scratch", None) is None or self._x_scratch.numel() < n_scratch:
            self._x_scratch = torch.empty(n_scratch + n_scratch // 4, dtype=torch.int32, device=dev)
        out = (C.c_int64 * 2)()
        tab = table(arrays)
        check(L.pnb_slab_unpack_f32(C.byref(tab), n, self.ndims, pmin, cs, self.z_lo, self.z_hi,
                                    int(has_up), int(has_down), self._x_idx[2].data_ptr(), n_leave,
                                    recv_up.data_ptr(), n_ru, recv_down.data_ptr(), n_rd,
                                    send_up.data_ptr(), n_up, send_down.data_ptr(), n_down,
                                    self._x_scratch.data_ptr(), out, stream))
        n_own, n_local = int(out[0]), int(out[1])
        self.last_stats = {"sent_up": n_up, "sent_down": n_down, "ghosts": n_local - n_own,
                           "migrated_in": n_own - n_stay, "migrated_out": n_leave,
                           "bytes_sent": (n_up + n_down) * W * 4}
        return arrays, n_own, n_local

    def exchange_inplace(self, arrays, n):
        """Structure-of-arrays form used every step.  arrays: list of float32 tensors with a
        common leading capacity dimension, arrays[0] = coordinates (cap, ndims), the others (cap,)
        or (cap, k) per-point state; the first n rows are the points this rank owns.  Migrants and
        ghosts are exchanged with the two neighbours in one round; on return rows [0, n_own) are
        the owned points (holes left by emigrants are filled from the tail, so the order of owned
        points may change), rows [n_own, n_local) the ghosts.  Only boundary rows are ever copied.
        Returns (arrays, n_own, n_local); buffers are re-allocated if too small."""
        import torch
        nd = self.ndims
        coords = arrays[0]
        dev = coords.device
        if coords.is_cuda and all(a.dtype == torch.float32 and a.is_contiguous() for a in arrays) \
                and len(arrays) <= 8 and not getattr(self, "force_torch_path", False):
            return self._exchange_inplace_cuda(arrays, n)
        has_up, has_down = self.rank + 1 < self.world, self.rank > 0
        widths = [1 if a.ndim == 1 else a.shape[1] for a in arrays]
        empty = torch.zeros(0, dtype=torch.int64, device=dev)
        if coords.is_cuda:
            # one pass of the library over the owned points instead of ~12 elementwise / nonzero
            # passes (and three host synchronisations) of torch
            up_idx, down_idx, leave_idx = self._classify_cuda(coords, n, has_up, has_down)
        else:
            cz = self.cell_layer(coords[:n])
            up_idx = torch.nonzero(cz >= self.z_hi).flatten() if has_up else empty
            down_idx = torch.nonzero(cz <= self.z_lo).flatten() if has_down else empty
            # leaving through the global bottom / top: kept, the following update! raises the
            # reference's domain error instead of losing the particle
            leave = torch.zeros_like(cz, dtype=torch.bool)
            if has_down:
                leave |= cz < self.z_lo
            if has_up:
                leave |= cz > self.z_hi
            leave_idx = torch.nonzero(leave).flatten()
        cz_up = self.cell_layer(coords[up_idx])
        cz_down = self.cell_layer(coords[down_idx])

        def pack(idx):
            return torch.cat([a[idx].reshape(idx.numel(), w) for a, w in zip(arrays, widths)],
                             dim=1).contiguous()

        send_up, send_down = pack(up_idx), pack(down_idx)
        recv_up, recv_down = self._sendrecv(send_up, send_down)
        cz_u = self.cell_layer(recv_up[:, :nd])
        cz_d = self.cell_layer(recv_down[:, :nd])
        # what I sent and is now one layer outside my slab stays with me as a ghost
        my_ghost_up = send_up[cz_up == self.z_hi + 1] if has_up else send_up[:0]
        my_ghost_down = send_down[cz_down == self.z_lo - 1] if has_down else send_down[:0]
        mig_in = torch.cat([recv_up[cz_u <= self.z_hi], recv_down[cz_d >= self.z_lo]])
        ghosts = torch.cat([recv_down[cz_d == self.z_lo - 1], my_ghost_down,
                            recv_up[cz_u == self.z_hi + 1], my_ghost_up])
        # emigrants: fill their holes from the tail (swap-with-last, like deleteatat!,
        # src/vector_of_vectors.jl:123-139)
        n_leave = int(leave_idx.numel())
        n_stay = n - n_leave
        if n_leave:
            tail = torch.arange(n_stay, n, device=dev)
            leaving_tail = torch.zeros(n - n_stay, dtype=torch.bool, device=dev)
            leaving_tail[leave_idx[leave_idx >= n_stay] - n_stay] = True
            fillers = tail[~leaving_tail]
            holes = leave_idx[leave_idx < n_stay]
            for a in arrays:
                a[holes] = a[fillers]
        n_own = n_stay + mig_in.shape[0]
        n_local = n_own + ghosts.shape[0]
        if n_local > coords.shape[0]:
            cap = n_local + n_local // 8 + 1024
            grown = []
            for a in arrays:
                g = torch.empty((cap,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev)
                g[:n_stay] = a[:n_stay]
                grown.append(g)
            arrays = grown
        col = 0
        for a, w in zip(arrays, widths):
            a[n_stay:n_own] = mig_in[:, col:col + w].reshape((mig_in.shape[0],) + tuple(a.shape[1:]))
            a[n_own:n_local] = ghosts[:, col:col + w].reshape((ghosts.shape[0],) + tuple(a.shape[1:]))
            col += w
        self.last_stats = {"sent_up": int(send_up.shape[0]), "sent_down": int(send_down.shape[0]),
                           "ghosts": int(ghosts.shape[0]), "migrated_in": int(mig_in.shape[0]),
                           "migrated_out": n_leave,
                           "bytes_sent": int((send_up.numel() + send_down.numel()) * 4)}
        return arrays, n_own, n_local


class SlabNeighborhoodSearch:
    """A GridNeighborhoodSearch over one slab (window of the global grid) + its exchange."""

    def __init__(self, ndims: int, search_radius, min_corner, max_corner, rank: int, world: int,
                 group=None):
        from .api import FullGridCellList, GridNeighborhoodSearch
        self.exchange = SlabExchange(ndims, search_radius, min_corner, max_corner, rank, world, group)
        r = np.float32(search_radius)
        cl = FullGridCellList(min_corner=min_corner, max_corner=max_corner, search_radius=r)
        self.nhs = GridNeighborhoodSearch(ndims, search_radius=r, cell_list=cl)
        self.nhs._window = self.exchange.window
        self.ndims = int(ndims)

    def step_inputs(self, rows):
        """Exchange, then split the local rows into contiguous coordinates and the other columns
        (convenience form for tests; step loops call exchange.exchange_inplace)."""
        local, n_own = self.exchange.exchange(rows)
        coords = local[:, :self.ndims].contiguous()
        return local, coords, n_own

    def update_(self, coords):
        from .api import update_
        return update_(self.nhs, coords, coords, points_moving=(True, True))


# ---------------------------------------------------------------------------------------------
# row exchange over NVLink peer memory (csrc/link.cu)
# ---------------------------------------------------------------------------------------------
class LinkUnavailable(RuntimeError):
    """The peer-memory link could not be set up on some rank (raised on every rank)."""


class SlabLink:
    """Receive area of this rank + the mapped areas of rank - 1 / rank + 1 (cudaIpc).  The handles
    travel once through torch.distributed (all_gather of 64 bytes); afterwards a step's exchange is
    two kernels on the sender and one on the receiver, no NCCL call."""

    def __init__(self, exchange: "SlabExchange", cap_rows: int, width: int, device):
        import torch
        import torch.distributed as dist
        L = _lib.lib()
        self.handle = C.c_void_p()
        self.cap, self.width = int(cap_rows), int(width)
        # every rank must use the same capacity
        c = torch.tensor([self.cap], dtype=torch.int64, device=device)
        dist.all_reduce(c, op=dist.ReduceOp.MAX, group=exchange.group)
        self.cap = int(c.item())
        # every collective below is executed by every rank whatever happens locally; a rank that
        # cannot create, export or map a receive area (cudaIpc unavailable: no peer access, a
        # container without shared IPC namespace) makes ALL ranks raise LinkUnavailable together
        err = None
        blob = (C.c_ubyte * 64)()
        try:
            if os.environ.get("PNB_SLAB_LINK_FAIL_RANK", "") == str(exchange.rank):     # test hook
                raise RuntimeError("PNB_SLAB_LINK_FAIL_RANK")
            check(L.pnb_slab_link_create(self.cap, self.width, C.byref(self.handle)))
            check(L.pnb_slab_link_export(self.handle, blob))
        except Exception as exc:        # noqa: BLE001 -- reported below, on every rank
            err = exc
        mine = torch.tensor(list(blob), dtype=torch.uint8, device=device)
        every = [torch.zeros_like(mine) for _ in range(exchange.world)]
        dist.all_gather(every, mine, group=exchange.group)
        r = exchange.rank

        def raw(k):
            if k < 0 or k >= exchange.world:
                return None
            return (C.c_ubyte * 64)(*every[k].cpu().tolist())

        if err is None:
            try:
                check(L.pnb_slab_link_connect(self.handle, raw(r - 1), raw(r + 1)))
            except Exception as exc:    # noqa: BLE001
                err = exc
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int64, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=exchange.group)
        if int(ok.item()) == 0:
            self.close()
            raise LinkUnavailable(str(err) if err is not None else "a neighbouring rank could not map the link")
        self.seq = 0
        self.leave_idx = torch.empty(2 * self.cap, dtype=torch.int32, device=device)
        dist.barrier(group=exchange.group)

    def close(self):
        if self.handle:
            _lib.lib().pnb_slab_link_destroy(self.handle)
            self.handle = C.c_void_p()


# ---------------------------------------------------------------------------------------------
# the overlapped WCSPH step of one rank (CUDA only)
# ---------------------------------------------------------------------------------------------
class OverlappedWCSPHStep:
    """update! + WCSPH interact! of one slab with the migrant + ghost exchange off the critical path
    (DESIGN.md 6).  Default (MODE "gather", EXCHANGE "p2p"):

        main : pnb_slab_link_send -- one kernel classifies the owned points, packs the rows of the
               boundary layer / the leavers and stores them into the neighbours' memory (NVLink)
        main : stream-ordered update! of the OWNED points, payload gather of the interior layers
                                                          <- the neighbours' rows arrive meanwhile
        side : pnb_slab_link_recv (waits for both neighbours' flags; the host waits for this only)
        main : append the rows to the arrays (pnb_slab_append_strided_f32) and to the cell list
               (pnb_grid_append_f32), payload gather of the boundary layers, ONE sweep of all
               owned layers
        host : waits for an event behind the append (not for the sweep), checks the error word,
               enqueues the compaction (pnb_slab_compact_f32): leavers out, migrants in

    MODE "split" hides the exchange behind a separate sweep of the interior layers instead (two
    sweep launches); EXCHANGE "nccl" sends counts, then rows, through torch.distributed.  Only the
    owned layers are swept (the ghosts are candidates, never query points).  The first step of a
    search (no bucket capacity known yet) and steps whose update! overflowed a bucket or received a
    migrant deeper than DEPTH layers inside the slab run the same sequence without the overlap.
    Results: dv[:n_own_new] belongs to arrays[k][:n_own_new]; the sweep may still be running when
    step() returns (stream-ordered: use the arrays on the current stream, or synchronise)."""

    DEPTH = 2          # boundary layers per side that wait for the exchange
    RESERVE_CTAS = 2   # SMs the interior sweep leaves to the NCCL kernels (MODE "split")
    SIDE_PRIORITY = os.environ.get("PNB_SLAB_PRIORITY", "1") != "0"
    PIPELINE = os.environ.get("PNB_SLAB_PIPELINE", "1") != "0"     # no host wait for the sweep
    EXCHANGE = os.environ.get("PNB_SLAB_EXCHANGE", "p2p")           # "p2p" (csrc/link.cu) | "nccl"
    LINK_SEND_ON_MAIN = os.environ.get("PNB_SLAB_SEND_ON_MAIN", "1") != "0"
    LINK_LAYERS = 4    # capacity of a link message in cell layers of owned points
    MODE = "gather"    # "gather": the exchange hides behind update! + the payload gather of the
    #                    interior layers, then ONE sweep of all owned layers;  "split": it hides
    #                    behind a separate sweep of the interior layers (two sweep launches)

    def __init__(self, slab: "SlabNeighborhoodSearch", closure_kwargs: dict):
        import torch
        self.slab = slab
        self.ex = slab.exchange
        self.kw = dict(closure_kwargs)
        # HIGH priority: the block scheduler hands free slots to the oldest kernel first, so the
        # classification / pack / NCCL kernels of an equal-priority side stream only start at the
        # tail of the main stream's kernel (measured: 0.92 ms of waiting behind 1.8 ms of cover)
        self.side = torch.cuda.Stream(priority=-1 if self.SIDE_PRIORITY else 0)
        self.flags = None
        self.scratch = None
        self.counters = None
        self.cnt_host = None
        self.link = None
        self.link_error = None
        self.last = {}

    def _table(self, arrs):
        t = _lib.SlabArrays()
        for k, a in enumerate(arrs):
            t.ptr[k] = a.data_ptr()
            t.width[k] = 1 if a.ndim == 1 else a.shape[1]
        t.n_arrays = len(arrs)
        return t

    def step(self, arrays, n, dv, overlap=True, profile=False):
        """arrays = [coords (cap, nd), v (cap, nd + 1), mass (cap,), pressure (cap,), ...], the first n
        rows owned; dv (cap, nd + 1).  Returns (arrays, dv, n_own_new).  profile=True records
        CUDA events between the phases on the main stream (self.last["phase_ms"])."""
        import torch
        from . import api as pn
        L = _lib.lib()
        ex, nhs, nd = self.ex, self.slab.nhs, self.slab.ndims
        dev = arrays[0].device
        main = torch.cuda.current_stream()
        stream = C.c_void_p(main.cuda_stream)
        sstream = C.c_void_p(self.side.cuda_stream)
        has_up, has_down = ex.rank + 1 < ex.world, ex.rank > 0
        W = sum(1 if a.ndim == 1 else a.shape[1] for a in arrays)
        pmin, cs = np.float32(ex.padded_min[-1]), np.float32(ex.search_radius)
        marks = []

        def mark(name):
            if profile:
                e = torch.cuda.Event(enable_timing=True)
                e.record(main)
                marks.append((name, e))

        smarks = []

        def smark(name):
            if profile:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.side)
                smarks.append((name, e))

        mark("start")
        coords = arrays[0]
        closure = pn.WCSPHInteract(dv, arrays[1], arrays[1], arrays[2], arrays[2], arrays[3], arrays[3],
                                   **self.kw)
        off = ex.window[0][-1] - 1                       # local cell layer = global - off
        D = self.DEPTH
        lo_i, hi_i = ex.z_lo + D, ex.z_hi - D            # interior layers (global)
        g = nhs._grid()
        can_overlap = overlap and nhs._handle is not None and nhs.layout() == "buckets" and lo_i <= hi_i
        # ---- 1. classification of the owned points on the SIDE stream (nothing on the main
        #         stream waits for it), 2. owned points: stream-ordered update! + interior layers ----
        self.side.wait_stream(main)
        cap = getattr(ex, "_x_cap", 0) or max(n // 8, 1 << 16)
        if getattr(ex, "_x_bufs", None) is None or ex._x_cap < cap:
            ex._x_idx = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3)]
            ex._x_cnt = torch.zeros(4, dtype=torch.int32, device=dev)
            ex._x_bufs = [torch.empty((cap, W), dtype=torch.float32, device=dev) for _ in range(2)]
            ex._x_cap = cap
        use_link = self.EXCHANGE == "p2p" and ex.world > 1
        if use_link:
            # ---- exchange over peer memory, part 1: classify + pack + NVLink stores in ONE kernel
            #      on the side stream (csrc/link.cu), enqueued before anything else of the step ------
            if self.link is None:
                layers = max(ex.z_hi - ex.z_lo + 1, 1)
                try:
                    self.link = SlabLink(ex, max(self.LINK_LAYERS * n // layers, 1 << 16), W, dev)
                except LinkUnavailable as exc:
                    # no peer mapping on this machine: the same rows travel through NCCL instead
                    # (every rank takes this branch together); said loudly, and in bench.py's line
                    import warnings
                    warnings.warn(f"pnb200: peer-memory link unavailable ({exc}); exchanging through NCCL")
                    self.EXCHANGE = "nccl"
                    self.link_error = str(exc)
                    use_link = False
        if use_link:
            link = self.link
            link.seq += 1
            tab = self._table(arrays)
            if self.LINK_SEND_ON_MAIN:
                # (next to the one-pass update! the send kernel and the build slow each other down:
                #  measured 0.9 + 1.3 ms side by side at 64 M points, 0.45 ms for the build alone)
                check(L.pnb_slab_link_send(link.handle, C.byref(tab), n, nd, pmin, cs, ex.z_lo, ex.z_hi,
                                           link.leave_idx.data_ptr(), link.seq, stream))
                self.side.wait_stream(main)
                mark("send")
            else:
                check(L.pnb_slab_link_send(link.handle, C.byref(tab), n, nd, pmin, cs, ex.z_lo, ex.z_hi,
                                           link.leave_idx.data_ptr(), link.seq, sstream))
                smark("sent")
        if can_overlap:
            check(L.pnb_grid_build_async_f32(g, coords.data_ptr(), n, stream))
            nhs._y_ref = coords
            can_overlap = nhs.layout() == "buckets"
        mark("update!")
        split = self.MODE == "split"
        if can_overlap and split:
            # the interior sweep is a persistent kernel: leave a few CTA slots free, otherwise the
            # NCCL kernels of the side stream are not scheduled before it ends (measured: the
            # whole exchange, 0.43 ms at 8 GPUs, was exposed behind a full grid)
            L.pnb_set_sweep_reserve(self.RESERVE_CTAS if (has_up or has_down) else 0)
            check(L.pnb_wcsph_interact_layers_async_f32(
                g, coords.data_ptr(), n, arrays[1].data_ptr(), arrays[2].data_ptr(), arrays[3].data_ptr(),
                C.byref(closure.params), dv.data_ptr(), lo_i - off, hi_i - off, 1, 0, 0, stream))
            L.pnb_set_sweep_reserve(0)
            mark("interior")
        elif can_overlap:
            # payload of the interior layers (the rest follows after the append)
            check(L.pnb_wcsph_interact_layers_async_f32(
                g, coords.data_ptr(), n, arrays[1].data_ptr(), arrays[2].data_ptr(), arrays[3].data_ptr(),
                C.byref(closure.params), dv.data_ptr(), lo_i - off, hi_i - off, 1, 0, 1, stream))
            mark("gather (interior)")
        if use_link:
            # ---- 3. part 2: wait for the neighbours' rows (one host synchronisation) ---------------
            p_down, p_up = C.c_void_p(), C.c_void_p()
            cnts = (C.c_int64 * 5)()
            check(L.pnb_slab_link_recv(link.handle, link.seq, C.byref(p_down), C.byref(p_up), cnts, sstream))
            smark("received")
            n_rd, n_ru, n_down, n_up, n_leave = (int(c) for c in cnts)
            recv_down_ptr, recv_up_ptr = p_down.value or 0, p_up.value or 0
            leave_ptr = link.leave_idx.data_ptr()
        elif ex.world == 1:
            # one slab: nothing to exchange (a point outside the grid is reported by update!)
            n_up = n_down = n_leave = n_ru = n_rd = 0
            recv_down_ptr = recv_up_ptr = 0
            leave_ptr = ex._x_idx[2].data_ptr()
        else:
            counts = (C.c_int64 * 3)()
            while True:
                check(L.pnb_slab_classify_f32(coords.data_ptr(), n, nd, pmin, cs, ex.z_lo, ex.z_hi, int(has_up),
                                              int(has_down), ex._x_idx[0].data_ptr(), ex._x_idx[1].data_ptr(),
                                              ex._x_idx[2].data_ptr(), ex._x_cap, ex._x_cnt.data_ptr(), counts,
                                              sstream))           # synchronises the side stream only
                if max(counts) <= ex._x_cap:
                    break
                cap = int(max(counts)) + int(max(counts)) // 4 + 1024
                ex._x_idx = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3)]
                ex._x_bufs = [torch.empty((cap, W), dtype=torch.float32, device=dev) for _ in range(2)]
                ex._x_cap = cap
            smark("classified")
            n_up, n_down, n_leave = (int(c) for c in counts)
            send_up, send_down = ex._x_bufs[0][:n_up], ex._x_bufs[1][:n_down]
            tab = self._table(arrays)
            check(L.pnb_slab_pack_rows_f32(C.byref(tab), ex._x_idx[0].data_ptr(), n_up, send_up.data_ptr(), sstream))
            check(L.pnb_slab_pack_rows_f32(C.byref(tab), ex._x_idx[1].data_ptr(), n_down, send_down.data_ptr(), sstream))
            # ---- 3. exchange on the side stream (NCCL: counts, then rows) ---------------------------
            smark("packed")
            with torch.cuda.stream(self.side):
                recv_up, recv_down = ex._sendrecv(send_up, send_down)
            smark("received")
            recv_up.record_stream(main)
            recv_down.record_stream(main)
            n_ru, n_rd = recv_up.shape[0], recv_down.shape[0]
            recv_down_ptr, recv_up_ptr = recv_down.data_ptr(), recv_up.data_ptr()
            leave_ptr = ex._x_idx[2].data_ptr()
        main.wait_stream(self.side)
        mark("wait for the exchange")
        n_app = n_ru + n_rd
        n_rows = n + n_app
        # ---- 4. append what arrived ---------------------------------------------------------------
        if n_rows > coords.shape[0]:
            # (rare: capacity exceeded) grow the arrays; the cell list was built from the old
            # coordinate array, so this step finishes without the overlap
            new_cap = n_rows + n_rows // 8 + 1024
            torch.cuda.synchronize()
            grown = []
            for a in arrays + [dv]:
                t = torch.empty((new_cap,) + tuple(a.shape[1:]), dtype=a.dtype, device=dev)
                t[:n] = a[:n]
                grown.append(t)
            arrays, dv = grown[:-1], grown[-1]
            coords = arrays[0]
            closure = pn.WCSPHInteract(dv, arrays[1], arrays[1], arrays[2], arrays[2], arrays[3],
                                       arrays[3], **self.kw)
            can_overlap = False
        if self.flags is None or self.flags.numel() < coords.shape[0]:
            self.flags = torch.empty(coords.shape[0], dtype=torch.uint8, device=dev)
            self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
        tab = self._table(arrays)
        row_stride = int(L.pnb_slab_link_row_stride(link.handle)) if use_link else 0
        check(L.pnb_slab_append_strided_f32(C.byref(tab), n, nd, pmin, cs, ex.z_lo, ex.z_hi,
                                            D - 1 if split else D, recv_up_ptr, n_ru, recv_down_ptr, n_rd,
                                            row_stride, self.flags.data_ptr(), self.counters.data_ptr(), stream))
        redo = not can_overlap
        if can_overlap:
            check(L.pnb_grid_append_f32(g, coords.data_ptr(), n, n_app, stream))
            mark("append")
            # what the host has to know about this step (migrants, bucket overflow) is final here:
            # an event behind the append, waited for AFTER the sweep has been launched, so that
            # the host never waits for the sweep itself and the next step's launches queue behind it
            if self.cnt_host is None:
                self.cnt_host = torch.zeros(4, dtype=torch.int32).pin_memory()
            self.cnt_host.copy_(self.counters, non_blocking=True)
            settled = torch.cuda.Event()
            settled.record(main)
            # ---- 5. boundary layers of both sides, one launch -----------------------------------
            a0, b0 = ex.z_lo, min(ex.z_lo + D - 1, ex.z_hi)
            a1, b1 = max(ex.z_hi - D + 1, ex.z_lo + D), ex.z_hi
            ptrs = (g, coords.data_ptr(), n_rows, arrays[1].data_ptr(), arrays[2].data_ptr(),
                    arrays[3].data_ptr(), C.byref(closure.params), dv.data_ptr())
            if split:
                check(L.pnb_wcsph_interact_layers_async_f32(*ptrs, a0 - off, b0 - off, a1 - off, b1 - off,
                                                            0, stream))
                mark("boundary")
            else:
                # payload of the layers the appended rows can lie in, then ONE sweep of all owned layers
                check(L.pnb_wcsph_interact_layers_async_f32(*ptrs, a0 - 1 - off, b0 - off, a1 - off,
                                                            b1 + 1 - off, 1, stream))
                mark("gather (boundary)")
                check(L.pnb_wcsph_interact_layers_async_f32(*ptrs, ex.z_lo - off, ex.z_hi - off, 1, 0, 2,
                                                            stream))
                mark("sweep")
            if self.PIPELINE:
                settled.synchronize()
                cnt_h = self.cnt_host.tolist()
                pn.check_(nhs, settled=True)
            else:
                cnt_h = self.counters.tolist()
                pn.check_(nhs)
            # a bucket overflowed (the library rebuilt the list) or a migrant landed deep inside
            # the slab: the sweeps above are void
            redo = nhs.layout() != "buckets" or cnt_h[1] != 0
        else:
            cnt_h = self.counters.tolist()
        n_mig = int(cnt_h[0])
        if redo:
            loc = coords[:n_rows]
            pn.update_(nhs, loc, loc, points_moving=(True, True))
            if nhs.layout() == "buckets":
                check(L.pnb_wcsph_interact_layers_async_f32(
                    g, loc.data_ptr(), n_rows, arrays[1].data_ptr(), arrays[2].data_ptr(),
                    arrays[3].data_ptr(), C.byref(closure.params), dv.data_ptr(), ex.z_lo - off,
                    ex.z_hi - off, 1, 0, 0, stream))
                pn.check_(nhs)
            else:
                f = pn.WCSPHInteract(dv[:n_rows], arrays[1][:n_rows], arrays[1][:n_rows], arrays[2][:n_rows],
                                     arrays[2][:n_rows], arrays[3][:n_rows], arrays[3][:n_rows], **self.kw)
                pn.foreach_point_neighbor(f, loc, loc, nhs)
            mark("sweep (not overlapped)")
        # ---- 6. compaction: leavers out, migrants in (dv travels with its rows) ------------------
        need = 2 * (n_leave + n_app + 8) + 16
        if self.scratch is None or self.scratch.numel() < need:
            self.scratch = torch.empty(need + need // 4, dtype=torch.int32, device=dev)
        tab = self._table(arrays + [dv])
        check(L.pnb_slab_compact_f32(C.byref(tab), n, n_app, leave_ptr, n_leave, n_mig,
                                     self.flags.data_ptr(), self.scratch.data_ptr(),
                                     self.counters.data_ptr(), stream))
        mark("compaction")
        n_new = n - n_leave + n_mig
        self.last = {"sent_up": n_up, "sent_down": n_down, "received": n_app, "migrated_in": n_mig,
                     "migrated_out": n_leave, "ghosts": n_app - n_mig + n_leave, "overlapped": bool(can_overlap and not redo),
                     "bytes_sent": (n_up + n_down) * W * 4, "rows": n_rows}
        if profile:
            torch.cuda.synchronize()
            self.last["phase_ms"] = {marks[k + 1][0]: marks[k][1].elapsed_time(marks[k + 1][1])
                                     for k in range(len(marks) - 1)}
            # when the side stream reached its marks, counted from the start of the step
            self.last["phase_ms"].update({"side: " + nm: marks[0][1].elapsed_time(e) for nm, e in smarks})
        return arrays, dv, n_new


# ---------------------------------------------------------------------------------------------
# benchmark at N > 1 (called by bench.py): weak scaling, 254 lattice layers per GPU
# ---------------------------------------------------------------------------------------------
def lattice_planes(n, k_lo, k_hi, domain_n, seed, device):
    """Perturbed lattice planes k_lo..k_hi (1-based, last dimension) of an n x n x (.) lattice.
    The perturbation of a plane depends only on (seed, plane), so every rank generates identical
    coordinates for the planes it shares with a neighbour.  Float32, normalised by domain_n + 1."""
    import torch
    out = []
    gen = torch.Generator(device=device)
    kk = torch.arange(n * n, device=device, dtype=torch.int64)
    ix = (kk % n).to(torch.float64) + 1.0
    iy = (kk // n).to(torch.float64) + 1.0
    for k in range(k_lo, k_hi + 1):
        gen.manual_seed(seed * 1000003 + k)
        c = torch.stack([ix, iy, torch.full_like(ix, float(k))], dim=1)
        c += 0.05 * torch.randn(n * n, 3, device=device, dtype=torch.float64, generator=gen)
        c += 0.05 * torch.randn(n * n, 3, device=device, dtype=torch.float64, generator=gen)
        out.append((c.to(torch.float32) / np.float32(domain_n + 1)))
    return torch.cat(out).contiguous()


def bench_multi_gpu(args, rank, world, dev, metric, unit):
    """BASELINE config 5: the n^3 cloud (n = 504: 128 024 064 particles, 171^3 cells) cut into
    `world` slabs along the last cell dimension -- STRONG scaling, the cloud is the same for every
    N.  A step = migrant + ghost exchange, update!, WCSPH interact! of the owned layers, with the
    exchange hidden behind the sweep of the interior layers (OverlappedWCSPHStep)."""
    import json
    import time
    import torch
    import torch.distributed as dist
    from . import api as pn

    T = np.float32
    n = int(getattr(args, "slab_lattice", 504))
    overlap = not bool(getattr(args, "no_overlap", False))
    r = T(3.0) / T(n + 1)
    mn, mx = np.zeros(3, T), np.ones(3, T)
    slab = SlabNeighborhoodSearch(3, r, mn, mx, rank, world)
    ex = slab.exchange
    # lattice planes that can fall into my layers (3 planes per layer, one plane of margin)
    k_lo = max(1, 3 * (ex.z_lo - 2) - 1)
    k_hi = min(n, 3 * (ex.z_hi - 2) + 3)
    parts = []
    for k0 in range(k_lo, k_hi + 1, 32):          # plane blocks: bounded temporary memory
        cand = lattice_planes(n, k0, min(k0 + 31, k_hi), n, 1, dev)
        parts.append(cand[ex.owned_mask(cand)])
        del cand
    A = torch.cat(parts).contiguous()
    del parts
    # cell-sorted order like the single-GPU cloud (dimension 1 most significant)
    cellk = torch.floor(A.to(torch.float64) * (n + 1) / 3.0).to(torch.int64)
    key = (cellk[:, 0] * (n + 8) + cellk[:, 1]) * (n + 8) + cellk[:, 2]
    del cellk
    A = A[torch.sort(key, stable=True).indices].contiguous()
    del key
    N = A.shape[0]
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    B = (A + (T(4e-4) * r) * torch.randn(N, 3, device=dev, generator=gen)).contiguous()
    gen2 = torch.Generator(device=dev).manual_seed(2000 + rank)
    rho = 1000.0 + torch.rand(N, device=dev, generator=gen2, dtype=torch.float32)
    mass = torch.full((N,), float(T(0.1) * (r / T(3))), device=dev, dtype=torch.float32)
    pressure = T(100.0) * (rho - T(1000.0))
    vfull = torch.zeros((N, 4), device=dev, dtype=torch.float32)     # v = vcat(velocity, density')
    vfull[:, 3] = rho
    layer_pts = int(N / max(ex.z_hi - ex.z_lo + 1, 1))
    cap = N + 4 * layer_pts + N // 64 + 4096          # owned + ghost layers + head room

    def with_cap(t):
        out = torch.empty((cap,) + tuple(t.shape[1:]), device=dev, dtype=torch.float32)
        out[:N] = t
        return out

    # two independent particle buffers (positions A / B of the same cloud), each [coords, v, m, p]
    bufs = [[with_cap(X), with_cap(vfull), with_cap(mass), with_cap(pressure)] for X in (A, B)]
    dvs = [torch.empty((cap, 4), device=dev, dtype=torch.float32) for _ in range(2)]
    n_cur = [N, N]
    # pinned host copies of what a host-side caller owns (e2e leg)
    host = [[t[:N].cpu().pin_memory() for t in (b[0], b[1], b[3])] for b in bufs]
    host_dv = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
    del A, B, vfull, rho, mass, pressure
    h = T(r / T(2))
    stepper = OverlappedWCSPHStep(slab, dict(smoothing_length=h, sound_speed=T(10.0), alpha=T(0.02),
                                             beta=T(0.0), delta=T(0.1)))
    stepper.MODE = str(getattr(args, "slab_mode", None) or OverlappedWCSPHStep.MODE)
    phase_ev = []

    def step(s, timed=False, ovl=True, e2e=False):
        k = (s + 1) % 2
        if e2e:
            # the owned rows' coordinates, state and pressure come from pinned host memory
            m = min(n_cur[k], host[k][0].shape[0])
            bufs[k][0][:m].copy_(host[k][0][:m], non_blocking=True)
            bufs[k][1][:m].copy_(host[k][1][:m], non_blocking=True)
            bufs[k][3][:m].copy_(host[k][2][:m], non_blocking=True)
        if timed:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            evs[0].record()
        bufs[k], dvs[k], n_cur[k] = stepper.step(bufs[k], n_cur[k], dvs[k], overlap=ovl)
        if timed:
            evs[1].record()
            phase_ev.append(evs)
        if e2e:
            host_dv[:n_cur[k]].copy_(dvs[k][:n_cur[k]], non_blocking=True)
        return n_cur[k]

    # pairs of the owned points of both clouds (untimed): full count sweep on the local window
    def owned_pairs(k):
        # rebuild the local cloud (owned + ghosts) the way the step saw it: one more exchange
        arrs, n_own, n_local = ex.exchange_inplace([a.clone() for a in bufs[k]], n_cur[k])
        loc = arrs[0][:n_local].contiguous()
        pn.update_(slab.nhs, loc, loc, points_moving=(True, True))
        cnt = torch.zeros(n_local, dtype=torch.int64, device=dev)
        pn.foreach_point_neighbor(pn.CountNeighbors(cnt), loc, loc, slab.nhs)
        return int(cnt[:n_own].sum())

    step(0, ovl=False)
    step(1, ovl=False)
    pairs = [0, 0]
    pairs[1] = owned_pairs(1)                 # step s uses buffer (s + 1) % 2
    pairs[0] = owned_pairs(0)
    for s in range(max(args.warmup, 3)):
        step(s, ovl=overlap)
    torch.cuda.synchronize()
    launches0 = int(_lib.lib().pnb_launch_count())
    _lib.profile(enable=True, reset=True)
    _lib.profile(reset=True)
    sampler = None
    if rank == 0 and getattr(args, "clock_sampler_cls", None) is not None:
        sampler = args.clock_sampler_cls(dev.index or 0)
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()                 # all ranks start the timed steps together (the contract's barrier)
    torch.cuda.synchronize()
    ev0.record()
    my_pairs = 0
    n_overlapped = 0
    for s in range(args.steps):
        step(s, timed=True, ovl=overlap)
        n_overlapped += int(stepper.last.get("overlapped", False))
        my_pairs += pairs[(s + 1) % 2]
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    dist.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = int(_lib.lib().pnb_launch_count()) - launches0
    prof = _lib.profile(enable=False)
    stats = dict(stepper.last)
    quick = bool(getattr(args, "slab_quick", False))      # headline only (sub-line of the N = 1 bench)
    # phases of the overlapped step (CUDA events on the main stream, a few extra steps)
    phase_acc = {}
    for s in range(0 if quick else 4):
        k = (s + 1) % 2
        bufs[k], dvs[k], n_cur[k] = stepper.step(bufs[k], n_cur[k], dvs[k], overlap=overlap, profile=True)
        for name, v_ in stepper.last.get("phase_ms", {}).items():
            phase_acc[name] = phase_acc.get(name, 0.0) + v_ / 4
    # the same steps without the overlap (exposed exchange = difference), then from host buffers
    k_ab = max(2, min(args.steps, 4))
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    a0.record()
    for s in range(0 if quick else k_ab):
        step(s, ovl=False)
    a1.record()
    torch.cuda.synchronize()
    ms_noovl = a0.elapsed_time(a1) / k_ab
    dist.barrier()
    # ... and with the other way of hiding the exchange
    mode_main = stepper.MODE
    stepper.MODE = "split" if mode_main == "gather" else "gather"
    if not quick:
        step(0, ovl=True)
        step(1, ovl=True)
    torch.cuda.synchronize()
    dist.barrier()
    a0.record()
    for s in range(0 if quick else k_ab):
        step(s, ovl=True)
    a1.record()
    torch.cuda.synchronize()
    ms_other = a0.elapsed_time(a1) / k_ab
    mode_other, stepper.MODE = stepper.MODE, mode_main
    dist.barrier()
    # ---- e2e: every step's inputs come from pinned host memory, its dv goes back ------------------
    # Pipelined like pnb_hoststep_* on one GPU: the H2D copies of step s + 1 (copy-in stream) and the
    # D2H copy of step s - 1 (copy-out stream) overlap the kernels of step s; the two particle
    # buffers alternate, so step s + 1 writes the buffer step s - 1 has finished with.
    k_e2e = 1 if quick else max(2, min(args.steps, 20))
    copy_in, copy_out = torch.cuda.Stream(), torch.cuda.Stream()
    main_s = torch.cuda.current_stream()
    ev_in, ev_done, ev_out = {}, {}, {}

    def issue_h2d(s):
        k = (s + 1) % 2
        if s - 2 in ev_done:
            copy_in.wait_event(ev_done[s - 2])        # the buffer's previous step is through
        else:
            copy_in.wait_stream(main_s)
        m = min(n_cur[k], host[k][0].shape[0])
        with torch.cuda.stream(copy_in):
            bufs[k][0][:m].copy_(host[k][0][:m], non_blocking=True)
            bufs[k][1][:m].copy_(host[k][1][:m], non_blocking=True)
            bufs[k][3][:m].copy_(host[k][2][:m], non_blocking=True)
            ev_in[s] = torch.cuda.Event()
            ev_in[s].record(copy_in)

    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    e2e_pairs = 0
    issue_h2d(0)
    for s in range(k_e2e):
        if s + 1 < k_e2e:
            issue_h2d(s + 1)
        k = (s + 1) % 2
        main_s.wait_event(ev_in[s])
        if s - 2 in ev_out:
            main_s.wait_event(ev_out[s - 2])          # dv of this buffer's previous step has left
        step(s, ovl=overlap)
        ev_done[s] = torch.cuda.Event()
        ev_done[s].record(main_s)
        copy_out.wait_event(ev_done[s])
        with torch.cuda.stream(copy_out):
            host_dv[:n_cur[k]].copy_(dvs[k][:n_cur[k]], non_blocking=True)
            ev_out[s] = torch.cuda.Event()
            ev_out[s].record(copy_out)
        e2e_pairs += pairs[k]
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / k_e2e
    h2d = int(sum(t.numel() * 4 for t in host[0]))
    d2h = int(n_cur[0] * 16)
    t = torch.tensor([ms, float(my_pairs), float(N), float(stats.get("bytes_sent", 0)),
                      float(stats.get("ghosts", 0)), ms_noovl, e2e_ms, float(e2e_pairs), float(h2d),
                      float(d2h), float(n_overlapped), ms_other], device=dev, dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    tmin = t.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    names = sorted(phase_acc)
    ph = torch.tensor([phase_acc[k_] for k_ in names], device=dev, dtype=torch.float64)
    if ph.numel():
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    phase_max = {k_: float(v_) for k_, v_ in zip(names, ph.tolist())}
    ph_mine = torch.tensor([phase_acc[k_] for k_ in names], device=dev, dtype=torch.float64)
    ph_all = [torch.zeros_like(ph_mine) for _ in range(world)] if names else []
    if ph_all:
        dist.all_gather(ph_all, ph_mine)
    phase_by_rank = {k_: [round(float(t_[i_]), 3) for t_ in ph_all] for i_, k_ in enumerate(names)}
    st_all = [torch.zeros(args.steps, device=dev, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(st_all, torch.tensor([a_.elapsed_time(b_) for a_, b_ in phase_ev[:args.steps]],
                                         device=dev, dtype=torch.float64))
    step_by_rank = [round(float(t_.median()), 3) for t_ in st_all]
    # where the timed loop spent its time besides the steps themselves (start-up skew of the ranks)
    gaps = [ev0.elapsed_time(phase_ev[0][0])] + \
           [phase_ev[k_][1].elapsed_time(phase_ev[k_ + 1][0]) for k_ in range(args.steps - 1)] + \
           [phase_ev[args.steps - 1][1].elapsed_time(ev1)]
    gp_all = [torch.zeros(len(gaps), device=dev, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gp_all, torch.tensor(gaps, device=dev, dtype=torch.float64))
    loop_by_rank = {"loop_ms": [round(float(a_.sum() + b_.sum()), 3) for a_, b_ in zip(st_all, gp_all)],
                    "first_step_ms": [round(float(t_[0]), 3) for t_ in st_all],
                    "gaps_ms": [[round(float(x_), 3) for x_ in t_] for t_ in gp_all]}
    if rank == 0:
        ms_max = float(tmax[0])
        total_pairs = float(tsum[1])
        gs = ex.grid_size
        # the dominant kernel on rank 0: k_sweep_flat (launches_per_step says how many per step)
        import os as _os
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        pk = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))),
                           "MEASURED_PEAKS.json")
        if _os.path.exists(pk):
            with open(pk) as fh:
                hbm_peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        sw_ms, sw_n = prof.get("k_sweep_cells", (0.0, 0))
        sweep_step_ms = sw_ms / max(args.steps, 1)
        cells_window = 1
        for a_, b_ in zip(ex.window[0], ex.window[1]):
            cells_window *= (b_ - a_ + 1)
        alg_bytes = 56 * N + 4 * (cells_window + 1)
        ach = alg_bytes / max(sweep_step_ms * 1e-3, 1e-9) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_sweep_flat<3,false,WcsphClT<false>,false> (rank 0, all launches "
                    "of a step)", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "launch_ms": sw_ms / max(sw_n, 1), "launches_per_step": sw_n / max(args.steps, 1),
                    "algorithmic_bytes": alg_bytes,
                    "note": "FP32-issue bound, not HBM bound (see the N = 1 line and DESIGN.md 5.2)"}
        line = {
            "metric": metric, "value": total_pairs / (ms_max * 1e-3), "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"WCSPH step 3D, slab-decomposed (BASELINE config 5): {n}^3 = {int(tsum[2])} "
                                   f"particles, {gs[0]}x{gs[1]}x{gs[2]} cells, over {world} GPU(s) "
                                   f"({(gs[2] - 2) // world} cell layers per GPU), per step: migrant + ghost "
                                   "exchange with the neighbouring slabs, update!, interact! of the owned layers",
                       "exchange": ("none (one slab)" if world == 1 else
                                    "NVLink peer memory: one kernel classifies, packs and stores the rows into the "
                                    "neighbour's buffer (csrc/link.cu)" if stepper.EXCHANGE == "p2p"
                                    else "NCCL send/recv (counts, then rows)"
                                    + (f" -- peer-memory link unavailable: {stepper.link_error}"
                                       if stepper.link_error else "")),
                       "particles_total": int(tsum[2]), "search_radius": float(r),
                       "velocities": "zero (reference benchmark)",
                       "ghost_points_per_rank_max": int(tmax[4]),
                       "exchange_bytes_per_rank_max": int(tmax[3]),
                       "l2": "per-GPU inputs larger than the 126 MB L2"},
            "overlap": {"enabled": overlap, "steps_overlapped_min_over_ranks": int(tmin[10]),
                        "ms_per_step_without_overlap": float(tmax[5]),
                        "mode": mode_main,
                        "ms_per_step_mode_" + mode_other: float(tmax[11]),
                        "what": {"gather": "exchange on a side stream behind update! and the payload gather "
                                           "of the interior layers; then the rows are appended and ONE "
                                           "sweep covers all owned layers",
                                 "split": "exchange on a side stream behind the sweep of the interior "
                                          "layers; the boundary layers (2 per side) wait for it"}[mode_main]},
            "step_ms_rank0": [round(a_.elapsed_time(b_), 3) for a_, b_ in phase_ev[:args.steps]],
            "phase_ms_rank0": phase_acc,
            "phase_ms_max_over_ranks": phase_max,
            "phase_ms_by_rank": phase_by_rank,
            "step_ms_median_by_rank": step_by_rank,
            "timed_loop_by_rank": loop_by_rank,
            "roofline": roofline,
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": {"value": float(tsum[7]) / k_e2e / (float(tmax[6]) * 1e-3), "unit": unit,
                    "h2d_bytes_per_step": int(tsum[8]), "d2h_bytes_per_step": int(tsum[9]),
                    "ms_per_step": float(tmax[6]), "steps": k_e2e,
                    "mode": "every rank copies the coordinates, state and pressure of its owned rows from "
                            "pinned host memory before every step and its dv back after it; copy-in and "
                            "copy-out streams overlap the copies of neighbouring steps with the kernels "
                            "(wall clock over the steps, max over ranks)"},
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
