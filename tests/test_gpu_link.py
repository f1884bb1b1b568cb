"""The peer-memory row exchange (csrc/link.cu) on ONE device: three slabs of a cloud held by
three links of the same process (pnb_slab_link_connect_local), so that the middle slab has two
neighbours.  What arrives must be exactly the rows pnb_slab_classify_f32 + pnb_slab_pack_rows_f32
(the verified two-step path of the NCCL exchange) select -- as sets, the order inside a message
is up to the atomics.  The 2- and 4-GPU tests run the same kernels through cudaIpc mappings."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def pn():
    import pnb200
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    return pnb200


def _sorted_rows(t):
    a = t if isinstance(t, np.ndarray) else t.cpu().numpy()
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("steps", [3])
def test_link_matches_classify_and_pack(pn, steps):
    from pnb200 import _lib
    from pnb200.slabs import SlabExchange
    L = _lib.lib()
    T = np.float32
    rng = np.random.default_rng(11)
    r = T(0.05)
    mn, mx = np.zeros(3, T), np.ones(3, T)
    world = 3
    exs = [SlabExchange(3, r, mn, mx, k, world) for k in range(world)]
    N = 60000
    base = rng.random((N, 3)).astype(T)
    widths = [3, 4, 1, 1, 1]                      # coordinates, v, mass, pressure, id: 10 floats
    W = sum(widths)
    cap = 1 << 15
    links = []
    for _ in range(world):
        h = C.c_void_p()
        _lib.check(L.pnb_slab_link_create(cap, W, C.byref(h)))
        links.append(h)
    for k in range(world):
        _lib.check(L.pnb_slab_link_connect_local(links[k], links[k - 1] if k > 0 else None,
                                                 links[k + 1] if k + 1 < world else None))
    stride = L.pnb_slab_link_row_stride(links[0])
    assert stride == 12
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        for seq in range(1, steps + 1):
            # every rank holds the points it owned at the PREVIOUS positions; they moved
            prev = base + (seq - 1) * T(0.01) * rng.standard_normal((N, 3)).astype(T)
            cur = np.clip(prev + T(0.3) * r * rng.standard_normal((N, 3)).astype(T), mn, mx).astype(T)
            prev_t = torch.as_tensor(np.clip(prev, mn, mx).astype(T), device="cuda")
            held = [exs[k].owned_mask(prev_t).cpu().numpy() for k in range(world)]
            expect, tabs, keep = [], [], []
            for k in range(world):
                idx = np.nonzero(held[k])[0]
                n = len(idx)
                arrs = [torch.as_tensor(cur[idx], device="cuda").contiguous(),
                        torch.as_tensor(rng.random((n, 4)).astype(T), device="cuda"),
                        torch.as_tensor(rng.random(n).astype(T), device="cuda"),
                        torch.as_tensor(rng.random(n).astype(T), device="cuda"),
                        torch.as_tensor(idx.astype(T), device="cuda")]
                tab = _lib.SlabArrays()
                for a_i, a in enumerate(arrs):
                    tab.ptr[a_i] = a.data_ptr()
                    tab.width[a_i] = widths[a_i]
                tab.n_arrays = len(arrs)
                keep.append(arrs)
                tabs.append((tab, n))
                # reference: classify + pack of the two-step exchange
                ex = exs[k]
                has_up, has_down = k + 1 < world, k > 0
                lists = [torch.empty(n + 1, dtype=torch.int32, device="cuda") for _ in range(3)]
                cnt_dev = torch.zeros(4, dtype=torch.int32, device="cuda")
                counts = (C.c_int64 * 3)()
                _lib.check(L.pnb_slab_classify_f32(arrs[0].data_ptr(), n, 3, T(ex.padded_min[-1]), r, ex.z_lo,
                                                   ex.z_hi, int(has_up), int(has_down), lists[0].data_ptr(),
                                                   lists[1].data_ptr(), lists[2].data_ptr(), n + 1,
                                                   cnt_dev.data_ptr(), counts, None))
                n_up, n_down, n_leave = (int(c) for c in counts)
                rows = torch.cat([a.reshape(n, -1) for a in arrs], dim=1)
                expect.append({"up": rows[lists[0][:n_up].long()], "down": rows[lists[1][:n_down].long()],
                               "leave": np.sort(lists[2][:n_leave].cpu().numpy())})
            leave_bufs = [torch.full((2 * cap,), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
            torch.cuda.synchronize()        # the sends run on other streams than the fills above
            for k in range(world):
                tab, n = tabs[k]
                ex = exs[k]
                _lib.check(L.pnb_slab_link_send(links[k], C.byref(tab), n, 3, T(ex.padded_min[-1]), r, ex.z_lo,
                                                ex.z_hi, leave_bufs[k].data_ptr(), seq,
                                                C.c_void_p(streams[k].cuda_stream)))
            for k in range(world):
                p_down, p_up = C.c_void_p(), C.c_void_p()
                cnts = (C.c_int64 * 5)()
                _lib.check(L.pnb_slab_link_recv(links[k], seq, C.byref(p_down), C.byref(p_up), cnts,
                                                C.c_void_p(streams[k].cuda_stream)))
                n_rd, n_ru, n_down, n_up, n_leave = (int(c) for c in cnts)
                assert n_down == expect[k]["down"].shape[0] and n_up == expect[k]["up"].shape[0]
                assert np.array_equal(np.sort(leave_bufs[k][:n_leave].cpu().numpy()), expect[k]["leave"])
                for n_r, ptr, src, key in ((n_rd, p_down, k - 1, "up"), (n_ru, p_up, k + 1, "down")):
                    if src < 0 or src >= world:
                        assert n_r == 0
                        continue
                    want = expect[src][key]
                    assert n_r == want.shape[0] and n_r > 0
                    got = np.empty((n_r, stride), np.float32)
                    _lib.check(L.pnb_memcpy_d2h(got.ctypes.data, ptr, got.nbytes, None))
                    assert (got[:, W:] == 0).all()                              # padding columns
                    assert np.array_equal(_sorted_rows(got[:, :W]), _sorted_rows(want)), (seq, k, key)
            del keep
    finally:
        torch.cuda.synchronize()
        for h in links:
            L.pnb_slab_link_destroy(h)


def test_link_errors(pn):
    from pnb200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.pnb_slab_link_create(0, 10, C.byref(h)) != 0
    assert L.pnb_slab_link_create(16, 64, C.byref(h)) != 0
    _lib.check(L.pnb_slab_link_create(1 << 10, 10, C.byref(h)))
    try:
        # receive before send, and a step number out of order
        cnts = (C.c_int64 * 5)()
        p0, p1 = C.c_void_p(), C.c_void_p()
        assert L.pnb_slab_link_recv(h, 1, C.byref(p0), C.byref(p1), cnts, None) != 0
        tab = _lib.SlabArrays()
        y = torch.rand(100, 3, device="cuda")
        s9 = torch.rand(100, 7, device="cuda")
        tab.ptr[0], tab.width[0] = y.data_ptr(), 3
        tab.ptr[1], tab.width[1] = s9.data_ptr(), 7
        tab.n_arrays = 2
        leave = torch.empty(2 << 10, dtype=torch.int32, device="cuda")
        assert L.pnb_slab_link_send(h, C.byref(tab), 100, 3, np.float32(0), np.float32(0.1), 2, 5,
                                    leave.data_ptr(), 2, None) != 0
        assert b"step" in L.pnb_last_error()
        assert L.pnb_slab_link_set_timeout(h, C.c_double(-1.0)) != 0
        _lib.check(L.pnb_slab_link_set_timeout(h, C.c_double(5.0)))
        # no neighbours: nothing is sent, nothing leaves
        _lib.check(L.pnb_slab_link_send(h, C.byref(tab), 100, 3, np.float32(0), np.float32(0.1), 2, 5,
                                        leave.data_ptr(), 1, None))
        _lib.check(L.pnb_slab_link_recv(h, 1, C.byref(p0), C.byref(p1), cnts, None))
        assert list(cnts) == [0, 0, 0, 0, 0]
    finally:
        L.pnb_slab_link_destroy(h)


def test_link_capacity_exceeded(pn):
    """A message larger than the link capacity: nothing is written behind the buffer (the rows that
    do not fit are dropped) and BOTH sides get PNB_ERR_LIST_FULL from the receive call."""
    from pnb200 import _lib
    from pnb200.slabs import SlabExchange
    L = _lib.lib()
    T = np.float32
    rng = np.random.default_rng(2)
    r = T(0.1)
    mn, mx = np.zeros(3, T), np.ones(3, T)
    exs = [SlabExchange(3, r, mn, mx, k, 2) for k in range(2)]
    pts = rng.random((20000, 3)).astype(T)
    cap = 64                                      # far fewer rows than a boundary layer holds
    links = []
    for _ in range(2):
        h = C.c_void_p()
        _lib.check(L.pnb_slab_link_create(cap, 3, C.byref(h)))
        links.append(h)
    _lib.check(L.pnb_slab_link_connect_local(links[0], None, links[1]))
    _lib.check(L.pnb_slab_link_connect_local(links[1], links[0], None))
    try:
        keep = []
        for k in range(2):
            t = torch.as_tensor(pts, device="cuda")
            own = t[exs[k].owned_mask(t)].contiguous()
            keep.append(own)
            tab = _lib.SlabArrays()
            tab.ptr[0], tab.width[0], tab.n_arrays = own.data_ptr(), 3, 1
            leave = torch.empty(2 * cap, dtype=torch.int32, device="cuda")
            keep.append(leave)
            torch.cuda.synchronize()
            _lib.check(L.pnb_slab_link_send(links[k], C.byref(tab), own.shape[0], 3, T(exs[k].padded_min[-1]), r,
                                            exs[k].z_lo, exs[k].z_hi, leave.data_ptr(), 1, None))
        for k in range(2):
            p0, p1 = C.c_void_p(), C.c_void_p()
            cnts = (C.c_int64 * 5)()
            st = L.pnb_slab_link_recv(links[k], 1, C.byref(p0), C.byref(p1), cnts, None)
            assert st == _lib.PNB_ERR_LIST_FULL, st
            assert b"capacity" in L.pnb_last_error()
            assert max(cnts[0], cnts[1]) > cap and max(cnts[2], cnts[3]) > cap
    finally:
        torch.cuda.synchronize()
        for h in links:
            L.pnb_slab_link_destroy(h)


def test_link_neighbour_timeout(pn):
    """A neighbour that never publishes its step: the waiting kernel gives up after the time limit
    and the receive call returns PNB_ERR_STATE -- the GPU is not left spinning."""
    import time
    from pnb200 import _lib
    L = _lib.lib()
    links = []
    for _ in range(2):
        h = C.c_void_p()
        _lib.check(L.pnb_slab_link_create(1 << 10, 3, C.byref(h)))
        links.append(h)
    _lib.check(L.pnb_slab_link_connect_local(links[0], None, links[1]))
    _lib.check(L.pnb_slab_link_connect_local(links[1], links[0], None))
    try:
        _lib.check(L.pnb_slab_link_set_timeout(links[0], C.c_double(0.3)))
        y = torch.rand(500, 3, device="cuda")
        tab = _lib.SlabArrays()
        tab.ptr[0], tab.width[0], tab.n_arrays = y.data_ptr(), 3, 1
        leave = torch.empty(2 << 10, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        _lib.check(L.pnb_slab_link_send(links[0], C.byref(tab), 500, 3, np.float32(0), np.float32(0.1), 2, 5,
                                        leave.data_ptr(), 1, None))
        p0, p1 = C.c_void_p(), C.c_void_p()
        cnts = (C.c_int64 * 5)()
        t0 = time.perf_counter()
        st = L.pnb_slab_link_recv(links[0], 1, C.byref(p0), C.byref(p1), cnts, None)
        dt = time.perf_counter() - t0
        assert st == _lib.PNB_ERR_STATE and b"did not send step 1" in L.pnb_last_error()
        assert 0.25 < dt < 5.0, dt
    finally:
        torch.cuda.synchronize()
        for h in links:
            L.pnb_slab_link_destroy(h)
