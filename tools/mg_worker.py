"""One rank of the multi-GPU parity test (tests/test_multigpu.py::_worker) as a stand-alone process,
so that compute-sanitizer can wrap it:  RANK=r WORLD_SIZE=n MASTER_PORT=p python tools/mg_worker.py"""
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import test_multigpu  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
out = os.environ.get("PNB_MG_OUT") or tempfile.mkdtemp()
test_multigpu._worker(rank, world, int(os.environ.get("MASTER_PORT", "29533")), out)
print(f"rank {rank}: worker finished", "ok" if rank != 0 or os.path.exists(os.path.join(out, "ok")) else "FAILED")
