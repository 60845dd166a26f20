"""IVectorIndex.Snapshot / Load through the C ABI: a freshly created index loaded from a snapshot answers
exactly like the index that wrote it, for all three kinds, and stays writable afterwards."""
import numpy as np
import pytest

from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    return pg


def _same(a, b):
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("kind", ["flat", "ivf_flat", "ivf_pq"])
def test_snapshot_load_round_trip(gpu, tmp_path, kind):
    base = orc.random_vectors(6000, 128, 42)
    q = orc.random_vectors(80, 128, 1337)
    mk = {"flat": lambda: gpu.GpuIndex(gpu.FLAT, 128, gpu.L2),
          "ivf_flat": lambda: gpu.GpuIndex(gpu.IVF_FLAT, 128, gpu.COSINE, nlist=16),
          "ivf_pq": lambda: gpu.GpuIndex(gpu.IVF_PQ, 128, gpu.L2, nlist=16, m=16, k=256)}[kind]
    a = mk()
    a.add(base[:5000])
    if kind != "flat":
        a.build()
    a.add(base[5000:])                      # FLAT: more rows; IVF: rows in the post-build buffer
    for r in (3, 77, 5500):
        if kind == "ivf_pq" and r < 5000:   # IVF_PQ deletes only buffered rows (IvfPqVectorIndex.cs:48-53)
            continue
        assert a.delete_row(r)
    kw = {} if kind == "flat" else {"nprobe": 4}
    want = a.search(q, 10, **kw)
    path = tmp_path / f"{kind}.bin"
    a.snapshot(path)
    b = mk()
    b.load(path)
    assert b.stats() == a.stats() and b.is_built() == a.is_built()
    _same(want, b.search(q, 10, **kw))
    if kind != "flat":
        np.testing.assert_array_equal(a.centroids(), b.centroids())
        _same(a.lists(), b.lists())
    # still writable: the same writes on both give the same answers
    extra = orc.random_vectors(20, 128, 5)
    for ix in (a, b):
        assert ix.add(extra) == 6000
        assert ix.delete_row(6001)
    _same(a.search(q, 10, **kw), b.search(q, 10, **kw))
    # wrong target / missing file
    other = gpu.GpuIndex(gpu.FLAT, 64, gpu.L2)
    with pytest.raises(gpu.PyropeGpuError):
        other.load(path)
    with pytest.raises(gpu.PyropeGpuError):
        b.load(tmp_path / "missing.bin")
