"""GPU tests of the string-id index classes (pyrope_b200/vector_index.py -> csrc/vindex.cu) and of Head+Tail on
the device (pyrope_delta_*: fused head ∪ tail search with the head's copy of an id winning, device-to-device
compaction).  The first half re-expresses the reference's own xunit cases
(tests/Pyrope.GarnetServer.Tests/Vector/{BruteForce,IvfFlat,IvfPq,Delta}VectorIndexTests.cs) against the GPU
classes, by the reference's names; the second half drives random write / compact / search sequences against the
CPU oracle (oracle/oracle.c through pyoracle) composed the way DeltaVectorIndex.cs composes its two sides.
Bar: distances within 1e-4 relative, ids identical modulo ties, centroids / lists bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.parity import assert_topk_equivalent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vi():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    from pyrope_b200 import vector_index
    return vector_index


# ------------------------------------------------------------------ BruteForceVectorIndexTests.cs
def test_bruteforce_cosine_returns_closest(vi):  # :10-20
    index = vi.BruteForceVectorIndex(2, vi.VectorMetric.Cosine)
    index.Add("a", [1.0, 0.0])
    index.Add("b", [0.0, 1.0])
    results = index.Search([1.0, 0.1], 1)
    assert len(results) == 1 and results[0].Id == "a"


def test_bruteforce_upsert_overwrites(vi):  # :23-33
    index = vi.BruteForceVectorIndex(2, vi.VectorMetric.InnerProduct)
    index.Add("a", [1.0, 0.0])
    index.Upsert("a", [0.0, 2.0])
    results = index.Search([0.0, 1.0], 1)
    assert results[0].Id == "a" and results[0].Score > 1.0


def test_bruteforce_delete_removes(vi):  # :36-46
    index = vi.BruteForceVectorIndex(2, vi.VectorMetric.L2)
    index.Add("a", [1.0, 1.0])
    assert index.Delete("a") is True
    assert index.Search([1.0, 1.0], 1) == []
    assert index.Delete("a") is False
    assert index.GetStats().Count == 0


def test_bruteforce_errors(vi):  # :49-53 and BruteForceVectorIndex.cs:141-144, 278, 381-390
    index = vi.BruteForceVectorIndex(2, vi.VectorMetric.L2)
    with pytest.raises(vi.ArgumentException, match="dimension"):
        index.Add("a", [1.0])
    index.Add("a", [1.0, 0.0])
    with pytest.raises(vi.InvalidOperationException, match="already exists"):
        index.Add("a", [0.0, 1.0])
    with pytest.raises(vi.ArgumentException, match="empty"):
        index.Add("   ", [0.0, 1.0])
    with pytest.raises(vi.ArgumentOutOfRangeException):
        index.Search([1.0, 0.0], 0)
    with pytest.raises(vi.ArgumentException, match="dimension"):
        index.Search([1.0, 0.0, 0.0], 1)
    with pytest.raises(vi.ArgumentNullException):
        index.Add("b", None)


def test_bruteforce_maxscans_zero_is_empty(vi):  # :56-65
    index = vi.BruteForceVectorIndex(2, vi.VectorMetric.InnerProduct)
    index.Add("a", [1.0, 0.0])
    index.Add("b", [0.0, 1.0])
    assert index.Search([1.0, 0.0], 1, vi.SearchOptions(MaxScans=0)) == []
    # scanLimit = min(MaxScans, count) <= 0 -> empty (BruteForceVectorIndex.cs:288-289): a negative budget is NOT "no budget"
    assert index.Search([1.0, 0.0], 1, vi.SearchOptions(MaxScans=-3)) == []
    assert len(index.Search([1.0, 0.0], 2, vi.SearchOptions(MaxScans=None))) == 2


def test_bruteforce_deleted_id_can_be_added_again(vi):  # Delete drops the id from _idMap (:240), Add appends
    index = vi.BruteForceVectorIndex(2, vi.VectorMetric.L2)
    index.Add("a", [1.0, 1.0])
    index.Add("b", [2.0, 2.0])
    index.Delete("a")
    index.Add("a", [3.0, 3.0])  # a new row at the END of the scan order
    res = index.Search([3.0, 3.0], 2, vi.SearchOptions(MaxScans=1))
    assert [r.Id for r in res] == ["b"]  # MaxScans counts live rows in insertion order: b comes first now
    assert index.GetStats().Count == 2


# ------------------------------------------------------------------ IvfFlatVectorIndexTests.cs
def test_ivfflat_centroids_null_before_build(vi):  # :12-19
    index = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=2)
    index.Add("a", [1.0, 0.0])
    assert index.GetCentroids() is None


def test_ivfflat_centroids_after_build(vi):  # :22-35
    index = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=2)
    for id_, v in (("a1", [0.1, 0.1]), ("a2", [0.2, 0.2]), ("b1", [10.1, 10.1]), ("b2", [10.2, 10.2])):
        index.Add(id_, v)
    index.Build()
    c = index.GetCentroids()
    assert c is not None and len(c) == 2 and all(len(x) == 2 for x in c)


def test_ivfflat_search_before_build_uses_buffer(vi):  # :52-66
    index = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=2)
    index.Add("a", [1.0, 0.0])
    index.Add("b", [5.0, 5.0])
    res = index.Search([1.0, 0.0], 1)
    assert len(res) == 1 and res[0].Id == "a"


def test_ivfflat_build_clusters_data(vi):  # :69-90
    index = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=2)
    for id_, v in (("a1", [0.1, 0.1]), ("a2", [0.2, 0.2]), ("b1", [10.1, 10.1]), ("b2", [10.2, 10.2])):
        index.Add(id_, v)
    index.Build()
    res = index.Search([0.0, 0.0], 2, vi.SearchOptions(NProbe=1))
    assert len(res) == 2 and all(r.Id.startswith("a") for r in res)


def test_ivfflat_nprobe_increases_recall(vi):  # :93-116
    index = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=3)
    index.Add("c1", [0.0, 0.0])
    index.Add("c2", [5.0, 5.0])
    index.Add("c3", [10.0, 10.0])
    index.Build()
    assert len(index.Search([0.0, 0.0], 3)) == 3  # default nprobe 3 reaches all three lists
    assert len(index.Search([0.0, 0.0], 3, vi.SearchOptions(NProbe=1))) == 1


def test_ivfflat_snapshot_load_preserves_state(vi, tmp_path):  # :119-141
    path = str(tmp_path / "ivf.snap")
    index = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=2)
    index.Add("a", [1.0, 0.0])
    index.Build()
    index.Add("late", [0.0, 3.0])  # still buffered when the snapshot is taken
    index.Snapshot(path)
    loaded = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=2)
    loaded.Load(path)
    res = loaded.Search([1.0, 0.0], 1)
    assert len(res) == 1 and res[0].Id == "a"
    assert {r.Id for r in loaded.Search([0.0, 3.0], 2)} == {"a", "late"}
    assert loaded.Delete("late") and loaded.GetStats().Count == 1
    with pytest.raises(FileNotFoundError):
        loaded.Load(str(tmp_path / "missing.snap"))
    with pytest.raises(vi.ArgumentException):
        loaded.Snapshot("  ")


def test_ivfflat_readd_of_indexed_id_shadows_then_rebuild_keeps_position(vi):
    # IvfFlatVectorIndex.cs:47 + :210 (seenIds) + :91-108 (uniqueData: the list entry keeps its place, takes the
    # buffer's vector).  Checked against the oracle, centroids and lists bit for bit.
    rng = np.random.default_rng(5)
    base = rng.random((400, 8), dtype=np.float32)
    ref = orc.IvfFlatIndex(8, orc.L2, nlist=6)
    index = vi.IvfFlatVectorIndex(8, vi.VectorMetric.L2, nList=6)
    for i in range(400):
        ref.add(i, base[i])
        index.Add(f"v{i}", base[i])
    ref.build()
    index.Build()
    newv = rng.random((40, 8), dtype=np.float32)
    for j, i in enumerate(range(10, 400, 10)):  # re-add ids that now sit in inverted lists
        ref.add(i, newv[j])
        index.Add(f"v{i}", newv[j])
    ref.add(1000, newv[39])
    index.Add("v1000", newv[39])
    q = rng.random((16, 8), dtype=np.float32)
    for qi in q:
        rid, rsc = ref.search(qi, 5, nprobe=6)
        got = index.Search(qi, 5, vi.SearchOptions(NProbe=6))
        assert_topk_equivalent(rid, rsc, [int(r.Id[1:]) for r in got], [r.Score for r in got], ctx="shadowed")
    assert index.GetStats().Count == ref.count()
    ref.build()
    index.Build()
    np.testing.assert_array_equal(ref.centroids(), np.stack(index.GetCentroids()))
    off, rows, _ = index.native().lists()
    ref_lists = ref.lists()
    assert [len(l) for l in ref_lists] == list(np.diff(off))
    for qi in q:
        rid, rsc = ref.search(qi, 5, nprobe=2)
        got = index.Search(qi, 5, vi.SearchOptions(NProbe=2))
        assert_topk_equivalent(rid, rsc, [int(r.Id[1:]) for r in got], [r.Score for r in got], ctx="rebuilt")
    assert index.GetStats().Count == ref.count() == 401


# ------------------------------------------------------------------ IvfPqVectorIndexTests.cs
def test_ivfpq_search_returns_results(vi):  # :41-67
    dim = 128
    index = vi.IvfPqVectorIndex(dim, vi.VectorMetric.L2, m=16, k=256, nList=4)
    rnd = orc.DotNetRandom(42)
    vecs = np.array([[rnd.next_double() for _ in range(dim)] for _ in range(300)], np.float32)
    for i in range(300):
        index.Add(str(i), vecs[i])
    index.Build()
    res = index.Search(vecs[0], 5)
    assert len(res) == 5
    assert index.GetStats().Count == 0  # IvfPqVectorIndex.cs:230 hard-codes 0


def test_ivfpq_delete_touches_only_the_buffer_and_build_replaces_lists(vi):  # IvfPqVectorIndex.cs:48-53, :64, :92
    rng = np.random.default_rng(9)
    index = vi.IvfPqVectorIndex(8, vi.VectorMetric.L2, m=2, k=16, nList=2)
    a = rng.random((40, 8), dtype=np.float32)
    for i in range(40):
        index.Add(f"a{i}", a[i])
    index.Build()
    assert index.Delete("a3") is False           # encoded rows cannot be deleted
    index.Add("a3", a[3] + 5)                    # buffered copy shadows the encoded one
    res = index.Search(a[3] + 5, 1)
    assert res[0].Id == "a3" and abs(res[0].Score) < 1e-6
    assert index.Delete("a3") is True            # buffer entry gone: the encoded copy answers again
    ids = [r.Id for r in index.Search(a[3], 40, vi.SearchOptions(NProbe=2))]
    assert ids.count("a3") == 1
    index.Add("b0", a[0] * 0.5)
    index.Build()                                # re-trains from the buffer only: the 40 encoded rows are gone
    res = index.Search(a[0], 10, vi.SearchOptions(NProbe=2))
    assert [r.Id for r in res] == ["b0"]


# ------------------------------------------------------------------ DeltaVectorIndexTests.cs
@pytest.fixture()
def delta2(vi):
    head = vi.BruteForceVectorIndex(2, vi.VectorMetric.L2)
    tail = vi.BruteForceVectorIndex(2, vi.VectorMetric.L2)
    return vi.DeltaVectorIndex(head, tail), head, tail


def test_delta_add_writes_to_head(delta2):  # :23-36
    d, head, tail = delta2
    d.Add("1", [1.0, 0.0])
    res = head.Search([1.0, 0.0], 1)
    assert len(res) == 1 and res[0].Id == "1"
    assert tail.Search([1.0, 0.0], 1) == []


def test_delta_search_merges(delta2):  # :39-50
    d, head, tail = delta2
    head.Add("head1", [1.0, 0.0])
    tail.Add("tail1", [0.0, 1.0])
    res = d.Search([1.0, 0.0], 10)
    assert [r.Id for r in res] == ["head1", "tail1"]


def test_delta_head_overrides_tail(delta2):  # :53-66
    d, head, tail = delta2
    tail.Add("doc1", [100.0, 100.0])
    head.Add("doc1", [1.0, 0.0])
    res = d.Search([1.0, 0.0], 10)
    assert len(res) == 1 and res[0].Id == "doc1" and abs(res[0].Score) < 1e-3


def test_delta_delete_propagates(delta2):  # :69-79
    d, head, tail = delta2
    head.Add("doc1", [1.0, 0.0])
    tail.Add("doc1", [1.0, 0.0])
    assert d.Delete("doc1") is True
    assert d.Search([1.0, 0.0], 10) == []


def test_delta_centroids(vi, delta2):  # :82-105
    head = vi.BruteForceVectorIndex(2, vi.VectorMetric.L2)
    tail = vi.IvfFlatVectorIndex(2, vi.VectorMetric.L2, nList=2)
    d = vi.DeltaVectorIndex(head, tail)
    tail.Add("a1", [0.1, 0.1])
    tail.Add("b1", [10.0, 10.0])
    tail.Build()
    c = d.GetCentroids()
    assert c is not None and len(c) > 0
    assert delta2[0].GetCentroids() is None  # BruteForce tail: not an ICentroidsProvider


def test_delta_ctor_mismatch_throws(vi):  # DeltaVectorIndex.cs:20-23
    with pytest.raises(vi.ArgumentException, match="dimensions"):
        vi.DeltaVectorIndex(vi.BruteForceVectorIndex(2), vi.BruteForceVectorIndex(3))
    with pytest.raises(vi.ArgumentException, match="metrics"):
        vi.DeltaVectorIndex(vi.BruteForceVectorIndex(2, vi.VectorMetric.L2),
                            vi.BruteForceVectorIndex(2, vi.VectorMetric.Cosine))


# ------------------------------------------------------------------ Head+Tail against the oracle
class RefDelta:
    """DeltaVectorIndex.cs over the CPU oracle's indexes (ids are ints; 'v{i}' on the GPU side)."""

    def __init__(self, dim, metric, tail):
        self.head = orc.FlatIndex(dim, metric)
        self.tail = tail
        self.head_rows = []  # (id, vec) in scan order, None once deleted: BruteForceVectorIndex.Scan()
        self.head_pos = {}

    def add(self, i, v):  # Upsert -> head (:46-57)
        self.head.upsert(i, v)
        if i in self.head_pos:
            self.head_rows[self.head_pos[i]] = (i, v)
        else:
            self.head_pos[i] = len(self.head_rows)
            self.head_rows.append((i, v))

    def delete(self, i):  # :59-74
        h = self.head.delete(i)
        if i in self.head_pos:
            self.head_rows[self.head_pos.pop(i)] = None
        t = self.tail.delete(i)
        return bool(h) or bool(t)

    def search(self, q, k, **kw):  # :76-122
        kw_t = dict(kw)
        hk = {"max_scans": kw["max_scans"]} if "max_scans" in kw else {}
        if isinstance(self.tail, orc.IvfPqIndex):
            kw_t.pop("max_scans", None)
        return orc.delta_merge(self.head.search(q, k, **hk), self.tail.search(q, k, **kw_t), k)

    def build(self):  # :124-158
        for item in self.head_rows:
            if item is None:
                continue
            self.tail.add(item[0], item[1])
            self.head.delete(item[0])
        self.head_rows, self.head_pos = [], {}
        self.tail.build()


def _check(ref, d, vi, Q, k, ctx, **kw):
    opts = vi.SearchOptions(MaxScans=kw.get("max_scans"), NProbe=kw.get("nprobe"))
    got = d.SearchBatch(Q, k, opts)
    for qi, g in zip(Q, got):
        rid, rsc = ref.search(qi, k, **kw)
        assert_topk_equivalent(rid, rsc, [int(r.Id[1:]) for r in g], [r.Score for r in g], ctx=ctx)


@pytest.mark.parametrize("tail_kind", ["ivf_flat", "ivf_pq"])
def test_delta_random_sequence_matches_oracle(vi, tail_kind):
    dim, n0 = 16, 1500
    rng = np.random.default_rng(11)
    base = rng.random((n0 + 600, dim), dtype=np.float32)
    Q = rng.random((24, dim), dtype=np.float32)
    if tail_kind == "ivf_flat":
        rt = orc.IvfFlatIndex(dim, orc.L2, nlist=12)
        gt = vi.IvfFlatVectorIndex(dim, vi.VectorMetric.L2, nList=12)
    else:
        rt = orc.IvfPqIndex(dim, orc.L2, m=4, k=64, nlist=12)
        gt = vi.IvfPqVectorIndex(dim, vi.VectorMetric.L2, m=4, k=64, nList=12)
    ref = RefDelta(dim, orc.L2, rt)
    d = vi.DeltaVectorIndex(vi.BruteForceVectorIndex(dim, vi.VectorMetric.L2), gt)
    # phase 1: everything arrives through the head, then one compaction builds the tail
    for i in range(n0):
        ref.add(i, base[i])
        d.Upsert(f"v{i}", base[i])
    _check(ref, d, vi, Q, 10, "head only")
    ref.build()
    d.Build()
    np.testing.assert_array_equal(rt.centroids(), np.stack(d.GetCentroids()))
    _check(ref, d, vi, Q, 10, "after first compaction", nprobe=4)
    # phase 2: new ids, overwrites of ids that now live in the tail (head must win), deletes on both sides
    for i in range(n0, n0 + 300):
        ref.add(i, base[i])
        d.Upsert(f"v{i}", base[i])
    for j, i in enumerate(range(0, n0, 25)):
        v = base[n0 + 300 + j]
        ref.add(i, v)
        d.Upsert(f"v{i}", v)
    if tail_kind == "ivf_flat":  # IVF_PQ cannot delete encoded rows (IvfPqVectorIndex.cs:48-53)
        for i in range(3, n0, 40):
            assert ref.delete(i) == d.Delete(f"v{i}")
    for i in range(n0 + 5, n0 + 300, 30):
        assert ref.delete(i) == d.Delete(f"v{i}")
    _check(ref, d, vi, Q, 10, "head over tail", nprobe=4)
    _check(ref, d, vi, Q, 1, "head over tail k=1", nprobe=12)
    _check(ref, d, vi, Q, 50, "head over tail k=50", nprobe=12)
    if tail_kind == "ivf_flat":
        _check(ref, d, vi, Q[:6], 10, "max_scans", nprobe=4, max_scans=120)
    # phase 3: second compaction folds the head into the tail (Dictionary order of uniqueData preserved)
    ref.build()
    d.Build()
    np.testing.assert_array_equal(rt.centroids(), np.stack(d.GetCentroids()))
    if tail_kind == "ivf_pq":
        off, rows, codes = gt.native().lists()
        for c, (ids, rc) in enumerate(rt.lists()):
            np.testing.assert_array_equal(rc, codes[off[c]:off[c + 1]])
    _check(ref, d, vi, Q, 10, "after second compaction", nprobe=4)
    assert d.Search(Q[0], 3) and d._head.GetStats().Count == 0


def test_delta_snapshot_load_roundtrip(vi, tmp_path):  # DeltaVectorIndex.cs:160-222
    dim = 8
    rng = np.random.default_rng(3)
    X = rng.random((300, dim), dtype=np.float32)
    d = vi.create_index("IVF_FLAT", dim, vi.VectorMetric.L2, {"nlist": 4})
    for i in range(250):
        d.Add(f"v{i}", X[i])
    d.Build()
    for i in range(250, 300):
        d.Add(f"v{i}", X[i])
    d.Upsert("v7", X[7] * 0.25)
    path = str(tmp_path / "delta.snap")
    d.Snapshot(path)
    assert open(path).read().startswith('{"Type": "Delta"')
    e = vi.create_index("IVF_FLAT", dim, vi.VectorMetric.L2, {"nlist": 4})
    e.Load(path)
    opts = vi.SearchOptions(NProbe=4)
    for q in X[::37]:
        a, b = d.Search(q, 5, opts), e.Search(q, 5, opts)
        assert [r.Id for r in a] == [r.Id for r in b]
        assert [r.Score for r in a] == [r.Score for r in b]
    assert e.GetStats() == d.GetStats()
    assert e.Delete("v7") and not e.Delete("v7")


def test_fvecs_streams_into_an_index(vi, tmp_path):  # FvecsReader.cs -> pyrope_index_add_fvecs
    import struct

    import pyrope_b200 as pg
    from pyrope_b200 import formats as fm
    rng = np.random.default_rng(21)
    X = rng.random((3000, 32), dtype=np.float32)
    p = tmp_path / "base.fvecs"
    with open(p, "wb") as f:
        for r in X:
            f.write(struct.pack("<i", 32))
            f.write(r.tobytes())
    ix = pg.GpuIndex(pg.FLAT, 32, pg.L2)
    assert fm.AddFvecs(ix, str(p), limit=2500) == 2500
    ref = orc.FlatIndex(32, orc.L2)
    ref.add_batch(fm.ReadFvecs(str(p), limit=2500))
    Q = rng.random((8, 32), dtype=np.float32)
    sc, rows, cnt = ix.search(Q, 10)
    rid, rsc, rcn = ref.search_batch(Q, 10)
    for i in range(8):
        assert_topk_equivalent(rid[i], rsc[i], rows[i], sc[i], ctx="fvecs")
    bad = pg.GpuIndex(pg.FLAT, 16, pg.L2)
    with pytest.raises(ValueError, match="dimension"):
        fm.AddFvecs(bad, str(p))


def test_bruteforce_quantization_flag(vi):  # BruteForceVectorIndexTests.cs:68-108
    index = vi.BruteForceVectorIndex(12, vi.VectorMetric.L2)
    index.EnableQuantization = True
    vec1 = np.zeros(12, np.float32)
    vec1[0] = 1.0
    index.Add("a", vec1)
    res = index.Search(vec1, 1)
    assert len(res) == 1 and res[0].Id == "a"
    # an upsert while the flag is off resets the quantised form: invisible to quantised search afterwards (:84-108)
    index.EnableQuantization = False
    index.Upsert("a", vec1)
    index.EnableQuantization = True
    assert index.Search(vec1, 1) == []
    ivf = vi.IvfFlatVectorIndex(4)  # the flag exists on BruteForceVectorIndex only
    import pyrope_b200 as pg
    assert pg.load().pyrope_vindex_set_quantization(ivf._v, 1) == pg._lib.ERR_INVALID_STATE
