"""Times k_link_classify_send on ONE device at the per-rank size of config 5 on 8 GPUs (an inner
slab: 21 cell layers of the 504^2 lattice, both neighbours): two links of this process connected
with pnb_slab_link_connect_local, so the row stores go to local memory instead of NVLink -- what
is left is the classification pass + packing.  Also the target of the ncu capture in profiles/."""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointneighbors.jl_b200"))
from pnb200 import _lib  # noqa: E402
from pnb200.slabs import SlabExchange, lattice_planes  # noqa: E402

L = _lib.lib()
T = np.float32
n = 504
r = T(3.0) / T(n + 1)
ex = SlabExchange(3, r, np.zeros(3, T), np.ones(3, T), 3, 8)
dev = torch.device("cuda", 0)
k_lo, k_hi = 3 * (ex.z_lo - 2) - 1, 3 * (ex.z_hi - 2) + 3
pts = lattice_planes(n, k_lo, k_hi, n, 1, dev)
pts = pts[ex.owned_mask(pts)].contiguous()
N = pts.shape[0]
arrs = [pts, torch.rand(N, 4, device=dev), torch.rand(N, device=dev), torch.rand(N, device=dev)]
W = 9
tab = _lib.SlabArrays()
for i, a in enumerate(arrs):
    tab.ptr[i] = a.data_ptr()
    tab.width[i] = 1 if a.ndim == 1 else a.shape[1]
tab.n_arrays = len(arrs)
cap = 4 * N // 21
links = []
for _ in range(3):
    h = C.c_void_p()
    _lib.check(L.pnb_slab_link_create(cap, W, C.byref(h)))
    links.append(h)
_lib.check(L.pnb_slab_link_connect_local(links[1], links[0], links[2]))
leave = torch.empty(2 * cap, dtype=torch.int32, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ms = []
for s in range(1, reps + 1):
    ev[0].record()
    _lib.check(L.pnb_slab_link_send(links[1], C.byref(tab), N, 3, T(ex.padded_min[-1]), r, ex.z_lo, ex.z_hi,
                                    leave.data_ptr(), s, None))
    ev[1].record()
    torch.cuda.synchronize()
    ms.append(ev[0].elapsed_time(ev[1]))
cnt = torch.zeros(4, dtype=torch.int32)
print(f"{N} points, row width {W} (stride {L.pnb_slab_link_row_stride(links[1])}): send {np.median(ms[3:]):.4f} ms "
      f"(min {min(ms):.4f}); algorithmic bytes: z of every point {4 * N / 1e6:.0f} MB (12-byte stride: "
      f"{12 * N / 1e6:.0f} MB of sectors) + 2 layers of rows read and written")
