"""The reference's own unit tests for the hot path, re-expressed against the CPU oracle.

Each test names the reference test it restates (paths under
/root/reference/tests/Pyrope.GarnetServer.Tests/Vector/).  These, plus the System.Random
known answers, are what pins the oracle: the reference holds no golden
top-k lists or fixtures for this path.
"""
import numpy as np
import pytest

from oracle import pyoracle as orc


# ---------------------------------------------------------------- System.Random known answers
def test_dotnet_random_known_answers():
    # widely published values of the legacy seeded System.Random
    assert orc.DotNetRandom(0).next() == 1559595546
    assert orc.DotNetRandom(1).next() == 534011718
    assert orc.DotNetRandom(42).next() == 1434747710
    assert orc.DotNetRandom(42).next_double() == 0.6681064659115423


def test_dotnet_random_sequence_properties():
    r = orc.DotNetRandom(42)
    xs = [r.next() for _ in range(1000)]
    assert all(0 <= x < 2**31 - 1 for x in xs)
    assert len(set(xs)) == 1000
    v = orc.random_vectors(10, 128, 42)
    assert v.dtype == np.float32 and v.shape == (10, 128)
    assert np.float32(0.6681064659115423) == v[0, 0]
    assert 0.45 < v.mean() < 0.55 and v.min() >= 0 and v.max() <= 1.0


# ---------------------------------------------------------------- VectorMathTests.cs
def test_dot_matches_reference():  # VectorMathTests.cs:10-21
    a = np.array([1, 2, 3, 4, 5], np.float32)
    b = np.array([2, 3, 4, 5, 6], np.float32)
    exp = np.float32(0)
    for x, y in zip(a, b):
        exp = np.float32(exp + np.float32(x * y))
    assert abs(orc.dot(a, b) - exp) <= 1e-6


def test_l2sq_matches_reference():  # :23-38
    a = np.array([1, 2, 3, 4, 5], np.float32)
    b = np.array([2, 3, 4, 5, 6], np.float32)
    assert abs(orc.l2sq(a, b) - 5.0) <= 1e-6


def test_norm_matches_reference():  # :40-51
    a = np.array([1, 2, 3, 4, 5], np.float32)
    assert abs(orc.norm(a) - np.float32(np.sqrt(np.float32(55.0)))) <= 1e-6


def test_cosine_matches_reference():  # :53-64
    assert abs(orc.cosine([1, 0, 0], [0, 1, 0])) <= 1e-6
    c = np.array([1, 2, 3], np.float32)
    assert abs(orc.cosine(c, c) - 1.0) <= 1e-6


def _ramps():
    dim = 1024 + 13
    a = (np.arange(dim, dtype=np.float32) * np.float32(0.001)).astype(np.float32)
    b = (np.arange(dim, dtype=np.float32) * np.float32(0.0005)).astype(np.float32)
    return a, b


def test_large_vector_matches_reference():  # :66-83 (tolerance 1.0 as in the reference)
    a, b = _ramps()
    ed = np.float32(0)
    el = np.float32(0)
    for x, y in zip(a, b):
        ed = np.float32(ed + np.float32(x * y))
        d = np.float32(x - y)
        el = np.float32(el + np.float32(d * d))
    assert abs(orc.dot(a, b) - ed) <= 1.0
    assert abs(orc.l2sq(a, b) - el) <= 1.0


def test_dimension_mismatch_throws():  # :85-107
    a, b = np.zeros(10, np.float32), np.zeros(11, np.float32)
    for fn in (orc.dot, orc.l2sq, orc.cosine):
        with pytest.raises(ValueError, match="dimension"):
            fn(a, b)


def test_unsafe_matches_safe():  # :108-130 (1e-4 abs as in the reference)
    a, b = _ramps()
    # the reference asserts |unsafe - safe| <= 1e-4 on values ~180; hold the oracle to the
    # fp32-resolution equivalent (1 ulp at 180 is 1.5e-5) and to float64 truth
    assert abs(orc.dot_unsafe(a, b) - orc.dot(a, b)) <= 1e-3
    assert abs(orc.l2sq_unsafe(a, b) - orc.l2sq(a, b)) <= 1e-3
    assert abs(orc.dot_unsafe(a, b) - float(np.dot(a.astype(np.float64), b.astype(np.float64)))) <= 2e-4 * 180
    d = a.astype(np.float64) - b.astype(np.float64)
    assert abs(orc.l2sq_unsafe(a, b) - float(d @ d)) <= 2e-4 * 90


# ---------------------------------------------------------------- BruteForceVectorIndexTests.cs
def test_flat_cosine_returns_closest():  # :10-20
    ix = orc.FlatIndex(2, orc.COSINE)
    ix.add(0, [1, 0])
    ix.add(1, [0, 1])
    ids, _ = ix.search([1, 0.1], 1)
    assert list(ids) == [0]


def test_flat_upsert_overwrites():  # :23-33
    ix = orc.FlatIndex(2, orc.IP)
    ix.add(0, [1, 0])
    ix.upsert(0, [0, 2])
    ids, sc = ix.search([0, 1], 1)
    assert ids[0] == 0 and sc[0] > 1.0


def test_flat_delete_removes():  # :36-46
    ix = orc.FlatIndex(2, orc.L2)
    ix.add(0, [1, 1])
    assert ix.delete(0)
    ids, _ = ix.search([1, 1], 1)
    assert len(ids) == 0


def test_flat_wrong_dimension_throws():  # :49-53
    ix = orc.FlatIndex(2, orc.L2)
    with pytest.raises(ValueError, match="dimension"):
        ix.add(0, [1])


def test_flat_maxscans_zero_returns_empty():  # :56-65
    ix = orc.FlatIndex(2, orc.IP)
    ix.add(0, [1, 0])
    ix.add(1, [0, 1])
    ids, _ = ix.search([1, 0], 1, max_scans=0)
    assert len(ids) == 0


def test_flat_duplicate_add_and_topk_validation():  # BruteForceVectorIndex.cs:143, :278
    ix = orc.FlatIndex(2, orc.L2)
    ix.add(7, [1, 0])
    with pytest.raises(KeyError):
        ix.add(7, [1, 0])
    with pytest.raises(IndexError):
        ix.search([1, 0], 0)


# ---------------------------------------------------------------- IvfFlatVectorIndexTests.cs
def test_ivfflat_centroids_before_after_build():  # :12-48
    ix = orc.IvfFlatIndex(2, orc.L2, nlist=2)
    ix.add(0, [1, 0])
    assert ix.centroids() is None
    ix = orc.IvfFlatIndex(2, orc.L2, nlist=2)
    for i, v in enumerate([[0.1, 0.1], [0.2, 0.2], [10.1, 10.1], [10.2, 10.2]]):
        ix.add(i, v)
    ix.build()
    c = ix.centroids()
    assert c is not None and c.shape == (2, 2)


def test_ivfflat_search_before_build_hits_buffer():  # :52-66
    ix = orc.IvfFlatIndex(2, orc.L2, nlist=2)
    ix.add(0, [1, 0])
    ix.add(1, [5, 5])
    ids, _ = ix.search([1, 0], 1)
    assert list(ids) == [0]


def test_ivfflat_build_clusters_data():  # :69-90
    ix = orc.IvfFlatIndex(2, orc.L2, nlist=2)
    names = {0: "a1", 1: "a2", 2: "b1", 3: "b2"}
    for i, v in enumerate([[0.1, 0.1], [0.2, 0.2], [10.1, 10.1], [10.2, 10.2]]):
        ix.add(i, v)
    ix.build()
    ids, _ = ix.search([0, 0], 2)
    assert len(ids) == 2
    assert all(names[int(i)].startswith("a") for i in ids)


def test_ivfflat_nprobe_all_returns_everything():  # :93-116
    ix = orc.IvfFlatIndex(2, orc.L2, nlist=3)
    ix.set_nprobe(1)
    for i, v in enumerate([[0, 0], [5, 5], [10, 10]]):
        ix.add(i, v)
    ix.build()
    ix.set_nprobe(3)
    ids, _ = ix.search([0, 0], 3)
    assert len(ids) == 3


def test_ivfflat_single_vector_build_and_search():  # :119-141 minus the JSON snapshot (out of scope)
    ix = orc.IvfFlatIndex(2, orc.L2, nlist=2)
    ix.add(0, [1, 0])
    ix.build()
    ids, _ = ix.search([1, 0], 1)
    assert list(ids) == [0]


def test_ivfflat_empty_search():  # :144-177 tail: empty index returns empty
    ix = orc.IvfFlatIndex(2, orc.L2)
    ids, _ = ix.search([0, 0], 1)
    assert len(ids) == 0


# ---------------------------------------------------------------- IvfPqVectorIndexTests.cs
def test_pq_train_and_encode_dimensions():  # :11-38
    dim, m, k = 16, 4, 256
    pq = orc.ProductQuantizer(dim, m, k)
    rng = orc.DotNetRandom(42)
    data = np.array([[np.float32(rng.next_double()) for _ in range(dim)] for _ in range(100)], np.float32)
    pq.train(data)
    code = pq.encode(np.full(dim, 0.5, np.float32))
    assert len(code) == m
    assert pq.ksub() == [100] * 4  # k clipped to the number of training points


def test_ivfpq_search_returns_results():  # :41-67
    dim = 128
    ix = orc.IvfPqIndex(dim, orc.L2, m=16, k=256, nlist=4)
    rng = orc.DotNetRandom(123)
    for i in range(100):
        ix.add(i, np.array([np.float32(rng.next_double()) for _ in range(dim)], np.float32))
    ix.build()
    ids, sc = ix.search(np.full(dim, 0.5, np.float32), 5)
    assert len(ids) == 5
    assert all(sc[i] >= sc[i + 1] for i in range(4))


def test_pq_ctor_validation():  # ProductQuantizer.cs:18-19
    with pytest.raises(ValueError):
        orc.ProductQuantizer(10, 3, 256)
    with pytest.raises(ValueError):
        orc.ProductQuantizer(16, 4, 257)


# ---------------------------------------------------------------- DeltaVectorIndexTests.cs
def test_delta_merge_returns_both():  # :39-50
    ids, sc = orc.delta_merge(([1], [-0.5]), ([2], [-0.25]), 2)
    assert list(ids) == [2, 1]


def test_delta_head_overrides_tail():  # :53-66
    ids, sc = orc.delta_merge(([1], [0.0]), ([1], [-4.0]), 1)
    assert list(ids) == [1] and abs(sc[0]) <= 0.001
