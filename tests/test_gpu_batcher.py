"""The host micro-batcher (csrc/batcher.cu): concurrent one-query Search calls — the way
VectorCommandSet.cs:458 drives an IVectorIndex — share batched launches and each gets its own result."""
import threading

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.parity import assert_topk_equivalent

pytestmark = pytest.mark.gpu


def test_concurrent_single_query_calls_are_batched():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    base = orc.random_vectors(8000, 128, 42)
    qs = orc.random_vectors(256, 128, 1337)
    ref = orc.IvfPqIndex(128, orc.L2, m=16, k=256, nlist=16)
    ref.add_batch(base)
    ref.build()
    ix = pg.GpuIndex(pg.IVF_PQ, 128, pg.L2, nlist=16, m=16, k=256)
    ix.add(base)
    ix.build()
    b = pg.Batcher(ix, max_batch=64, max_wait_us=2000)
    out = [None] * len(qs)

    def worker(lo, hi):
        for i in range(lo, hi):
            nprobe = 4 if i % 2 == 0 else 2          # two option groups in flight at once
            out[i] = b.search(qs[i], 10, nprobe=nprobe)

    th = [threading.Thread(target=worker, args=(t * 8, t * 8 + 8)) for t in range(32)]
    [t.start() for t in th]
    [t.join() for t in th]
    st = b.stats()
    assert st["queries"] == 256 and st["batches"] < 256          # callers really shared launches
    for i in range(len(qs)):
        ids, sc = ref.search(qs[i], 10, nprobe=4 if i % 2 == 0 else 2)
        gs, gr = out[i]
        assert_topk_equivalent(ids, sc, gr, gs, ctx=f"batched query {i}")
    # errors come back per caller, not as a crash
    with pytest.raises(pg.PyropeGpuError):
        flat = pg.GpuIndex(pg.FLAT, 128, pg.L2)
        flat.add(base[:100])
        pg.Batcher(flat).search(qs[0], 0)                          # topK <= 0 on FLAT -> OUT_OF_RANGE
    b.close()
