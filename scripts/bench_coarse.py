"""Coarse probe alone (csrc/coarse_tc.cu vs the streaming path): 10,000 queries x 65,536 random centroids, d = 128.
   python scripts/bench_coarse.py [nlist] [nq] [nprobe]     (PYROPE_COARSE_STREAMING=1 selects the old path)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyrope_b200 as pg  # noqa: E402
from pyrope_b200 import _lib  # noqa: E402

nlist = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
P = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dim = 128
_lib.check(pg.load().pyrope_gpu_init(0))
rng = np.random.default_rng(1)
cent = rng.random((nlist, dim), dtype=np.float32)
cb = (rng.random((16, 256, 8), dtype=np.float32) - 0.5) * 0.5
ix = pg.GpuIndex(pg.IVF_PQ, dim, pg.L2, nlist=nlist, m=16, k=256)
ix.set_codebooks(cent, cb)
ix.add(rng.random((200_000, dim), dtype=np.float32))
ix.build()
Q = torch.rand((nq, dim), device="cuda")
pr = torch.empty((nq, P), dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ix.coarse_probe_device(Q.data_ptr(), nq, P, pr.data_ptr(), stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    ix.coarse_probe_device(Q.data_ptr(), nq, P, pr.data_ptr(), stream=st)
e1.record()
torch.cuda.synchronize()
print(f"coarse probe nlist={nlist} nq={nq} P={P} streaming={os.environ.get('PYROPE_COARSE_STREAMING', '0')}: "
      f"{e0.elapsed_time(e1) / reps:.3f} ms per batch, launches {ix.last_search_launches()}")
if os.environ.get("PROFILE_ONE"):
    torch.cuda.profiler.start()
    ix.coarse_probe_device(Q.data_ptr(), nq, P, pr.data_ptr(), stream=st)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
