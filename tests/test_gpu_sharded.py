"""One index over several GPUs from ONE process (pyrope_sharded_*, csrc/sharded.cu): the entry point a C# GpuVectorIndex
P/Invokes to get N devices behind IVectorIndex.  Results must equal the single-GPU index's (and hence the oracle's) for
FLAT (row blocks) and IVF_FLAT / IVF_PQ (lists sharded by list id, coarse stage split by query, probe lists exchanged by
peer stores, in-kernel threshold exchange), including rows still in the pre-build buffer (they live on every shard: the
merge keeps one entry per row).  Runs with however many devices the box has (1, 2, 4 or 8)."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.parity import assert_batch_equivalent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    return pg


def _ndev(gpu):
    import ctypes as C
    n = C.c_int32(0)
    gpu._lib.check(gpu.load().pyrope_gpu_device_count(C.byref(n)))
    return n.value


def _counts(gpu):
    n = _ndev(gpu)
    return [c for c in (1, 2, 4, 8) if c <= n]


def _s(ix, Q, k, **kw):
    sc, rows, cnt = ix.search(Q, k, **kw)
    return rows, sc, cnt


def test_sharded_flat_matches_oracle(gpu):
    base = orc.random_vectors(30_000, 96, 42)
    Q = orc.random_vectors(200, 96, 1337)
    ref = orc.FlatIndex(96, orc.IP)
    ref.add_batch(base)
    want = ref.search_batch(Q, 20)
    for n in _counts(gpu):
        sx = gpu.ShardedIndex(n, gpu.FLAT, 96, gpu.INNER_PRODUCT)
        assert sx.add(base[:17_000]) == 0 and sx.add(base[17_000:]) == 17_000   # two add calls: 2 n row blocks
        assert sx.stats() == 30_000
        assert_batch_equivalent(want, _s(sx, Q, 20), ctx=f"sharded FLAT n={n}")
        assert sx.delete_row(int(want[0][0][0]))                               # a global row ordinal
        got = _s(sx, Q[:1], 20)
        assert int(want[0][0][0]) not in got[0][0]
        sx.close()


@pytest.mark.parametrize("kind", ["IVF_FLAT", "IVF_PQ"])
def test_sharded_ivf_equals_single_gpu_and_oracle(gpu, kind):
    dim, n_rows, nlist = 128, 60_000, 64
    base = orc.random_vectors(n_rows, dim, 7)
    extra = orc.random_vectors(300, dim, 9)                                     # stays in the write buffer after the build
    Q = orc.random_vectors(500, dim, 8)
    gk = gpu.IVF_FLAT if kind == "IVF_FLAT" else gpu.IVF_PQ
    one = gpu.GpuIndex(gk, dim, gpu.L2, nlist=nlist, m=16, k=256)
    one.add(base)
    one.build()
    one.add(extra)
    nprobe = 8
    want = _s(one, Q, 10, nprobe=nprobe)
    if kind == "IVF_PQ":
        ref = orc.IvfPqIndex(dim, orc.L2, m=16, k=256, nlist=nlist)
    else:
        ref = orc.IvfFlatIndex(dim, orc.L2, nlist=nlist)
    ref.add_batch(base)
    ref.build()
    ref.add_batch(extra, ids=np.arange(n_rows, n_rows + len(extra)))
    assert_batch_equivalent(ref.search_batch(Q, 10, nprobe=nprobe), want, ctx=f"{kind} single GPU vs oracle")
    for n in _counts(gpu):
        sx = gpu.ShardedIndex(n, gk, dim, gpu.L2, nlist=nlist, m=16, k=256)
        sx.add(base)
        sx.build()
        sx.add(extra)
        assert sx.stats() == n_rows + len(extra)
        for rep in range(3):                                                    # repeated batches: epochs of the exchange move on
            got = _s(sx, Q, 10, nprobe=nprobe)
            assert_batch_equivalent(want, got, ctx=f"sharded {kind} n={n} rep={rep}")
        assert_batch_equivalent(_s(one, Q[:7], 3, nprobe=2), _s(sx, Q[:7], 3, nprobe=2), ctx=f"sharded {kind} n={n} small batch")
        if n > 1:
            with pytest.raises(gpu.PyropeGpuError):
                sx.search(Q[:4], 5, max_scans=100, nprobe=2)                    # MaxScans cannot be split over devices
        sx.close()


def test_sharded_device_resident_search(gpu):
    import torch
    dim = 128
    base = orc.random_vectors(40_000, dim, 3)
    Q = orc.random_vectors(256, dim, 4)
    n = max(_counts(gpu))
    sx = gpu.ShardedIndex(n, gpu.IVF_PQ, dim, gpu.L2, nlist=32, m=16, k=256)
    sx.add(base)
    sx.build()
    want = sx.search(Q, 10, nprobe=4)
    torch.cuda.set_device(0)
    dq = torch.from_numpy(Q).cuda()
    sc = torch.empty((256, 10), dtype=torch.float32, device="cuda")
    rw = torch.empty((256, 10), dtype=torch.int64, device="cuda")
    cn = torch.empty((256,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    sx.search_device(dq.data_ptr(), 256, 10, sc.data_ptr(), rw.data_ptr(), cn.data_ptr(), nprobe=4)
    np.testing.assert_array_equal(want[0], sc.cpu().numpy())
    np.testing.assert_array_equal(want[1], rw.cpu().numpy())
    assert sx.last_search_ms() > 0
    sx.close()


def test_shard_views_are_borrowed(gpu):
    """ShardedIndex.shard(i) hands out a view for device-resident feeds; dropping the view must not destroy the shard."""
    import gc
    base = orc.random_vectors(5_000, 32, 1)
    sx = gpu.ShardedIndex(1, gpu.FLAT, 32, gpu.L2)
    sx.add(base)
    view, dev = sx.shard(0)
    assert dev == 0 and view.stats()["live"] == 5_000
    del view
    gc.collect()
    sc, rows, cnt = sx.search(base[:3], 1)
    assert list(rows[:, 0]) == [0, 1, 2]
    sx.close()
