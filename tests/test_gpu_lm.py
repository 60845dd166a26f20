"""GPU parity tests of the list-major IVF_PQ scan (pq_lm.cu: pairs grouped by list, four queries per work
item, interleaved float4 lookup tables, TMA-staged codes, exact re-score of the survivors) and of the
tensor-core shortlist used for coarse assignment at large nlist.  Same bar as tests/test_gpu_parity.py:
distances within 1e-4 relative of the oracle, ids identical modulo ties, assignments bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.parity import assert_batch_equivalent, recall_at_k

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    return pg


def _s(ix, Q, k, **kw):
    sc, rows, cnt = ix.search(Q, k, **kw)
    return rows, sc, cnt


def _pair(gpu, base, dim, nlist, m=16):
    ref = orc.IvfPqIndex(dim, orc.L2, m=m, k=256, nlist=nlist)
    ref.add_batch(base)
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_PQ, dim, gpu.L2, nlist=nlist, m=m, k=256)
    ix.add(base)
    ix.build()
    off, rows, codes = ix.lists()
    for c, (ids, rc) in enumerate(ref.lists()):
        np.testing.assert_array_equal(rc, codes[off[c]:off[c + 1]])
    return ref, ix


@pytest.fixture(scope="module")
def mid(gpu):
    base = orc.random_vectors(20_000, 128, 42)
    q = orc.random_vectors(300, 128, 1337)
    ref, ix = _pair(gpu, base, 128, 32)
    return base, q, ref, ix


def test_lm_many_queries_per_list(mid):
    """300 queries x 8 probes over 32 lists: ~75 queries per list, every slot of most items in use."""
    base, q, ref, ix = mid
    assert_batch_equivalent(ref.search_batch(q, 10, nprobe=8), _s(ix, q, 10, nprobe=8), ctx="lm nprobe=8")
    assert ix.last_search_launches() >= 8  # grouping kernels + scan + final + coarse + merge


def test_lm_distances_are_the_reference_values(mid):
    """Survivors are re-scored in the reference's order: scores must be bit-identical to the oracle's
    wherever the id matches."""
    base, q, ref, ix = mid
    rid, rsc, rcn = ref.search_batch(q[:64], 10, nprobe=4)
    gid, gsc, gcn = _s(ix, q[:64], 10, nprobe=4)
    same = rid == gid
    assert same.mean() > 0.99
    np.testing.assert_array_equal(rsc[same], gsc[same])


@pytest.mark.parametrize("k,nprobe", [(1, 1), (100, 16), (10, 32), (37, 5)])
def test_lm_topk_and_nprobe_shapes(mid, k, nprobe):
    base, q, ref, ix = mid
    assert_batch_equivalent(ref.search_batch(q[:120], k, nprobe=nprobe), _s(ix, q[:120], k, nprobe=nprobe),
                            ctx=f"lm k={k} nprobe={nprobe}")


def test_lm_single_query_and_ragged_groups(mid):
    base, q, ref, ix = mid
    for nq in (1, 2, 3, 5, 33):
        assert_batch_equivalent(ref.search_batch(q[:nq], 10, nprobe=3), _s(ix, q[:nq], 10, nprobe=3), ctx=f"lm nq={nq}")


def test_lm_long_lists_multi_segment_and_cold_queues(gpu):
    """Lists of ~4000 codes: several 1024-code TMA segments per item and, for the first items of every
    query (no threshold yet), the queue-prune path."""
    base = orc.random_vectors(16_000, 128, 7)
    q = orc.random_vectors(40, 128, 8)
    ref, ix = _pair(gpu, base, 128, 4)
    for k, nprobe in ((10, 4), (100, 2), (1000, 4)):
        assert_batch_equivalent(ref.search_batch(q, k, nprobe=nprobe), _s(ix, q, k, nprobe=nprobe),
                                ctx=f"lm long lists k={k} nprobe={nprobe}")


def test_lm_dim64_sub4(gpu):
    base = orc.random_vectors(6_000, 64, 11)
    q = orc.random_vectors(50, 64, 12)
    ref, ix = _pair(gpu, base, 64, 16)
    assert_batch_equivalent(ref.search_batch(q, 10, nprobe=4), _s(ix, q, 10, nprobe=4), ctx="lm dim=64")


def test_lm_agrees_with_query_major_kernel(gpu, mid, monkeypatch):
    base, q, ref, ix = mid
    cent = ix.centroids()
    cb, _ = ix.codebooks()
    monkeypatch.setenv("PYROPE_PQ_LM", "0")
    qm = gpu.GpuIndex(gpu.IVF_PQ, 128, gpu.L2, nlist=32, m=16, k=256)
    qm.set_codebooks(cent, cb)
    qm.add(base)
    qm.build()
    a = _s(qm, q, 10, nprobe=8)
    assert qm.last_search_launches() < 8
    b = _s(ix, q, 10, nprobe=8)
    assert_batch_equivalent(a, b, ctx="query-major vs list-major")
    assert recall_at_k(a[0], b[0], 10) > 0.999


def test_lm_shadowed_rows_and_buffer(gpu):
    base = orc.random_vectors(5_000, 128, 21)
    q = orc.random_vectors(30, 128, 22)
    ref, ix = _pair(gpu, base, 128, 8)
    # rows re-added after the build live in the exact buffer and shadow their list entries
    rng = np.random.default_rng(0)
    upd = rng.choice(5_000, 40, replace=False)
    for r in upd:
        v = rng.random(128, dtype=np.float32)
        ref.add_batch(v[None, :], ids=np.array([int(r)]))
        new_row = ix.add(v[None, :])
        ix.shadow_row(int(r), True)
        assert new_row >= 5_000
    rid, rsc, rcn = ref.search_batch(q, 10, nprobe=8)
    gsc, grows, gcn = ix.search(q, 10, nprobe=8)
    # map the GPU's fresh row ordinals of the re-added vectors back to their ids
    remap = {5_000 + i: int(r) for i, r in enumerate(upd)}
    gid = np.vectorize(lambda x: remap.get(int(x), int(x)))(grows)
    assert_batch_equivalent((rid, rsc, rcn), (gid, gsc, gcn), ctx="lm shadowed + buffer")


# ------------------------------------------------------------------------------------------------
# tensor-core shortlist for KMeansUtils.FindNearestCentroid at large nlist
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", [orc.L2, orc.IP, orc.COSINE])
def test_tc_shortlist_assign_bit_exact(gpu, monkeypatch, metric):
    rng = np.random.default_rng(3)
    dim, nc, n = 128, 3000, 6000
    cent = rng.random((nc, dim), dtype=np.float32)
    X = rng.random((n, dim), dtype=np.float32)
    cent[1500] = cent[7]            # duplicate centroids: the lower index must win
    cent[2999] = cent[7]
    X[:50] = cent[rng.choice(nc, 50)]  # rows that coincide with a centroid
    X[50:60] = cent[7]
    a = gpu.coarse_assign(X, cent, metric)           # tensor-core shortlist + exact re-evaluation
    monkeypatch.setenv("PYROPE_ASSIGN_EXACT", "1")
    b = gpu.coarse_assign(X, cent, metric)           # exhaustive exact kernel
    np.testing.assert_array_equal(a, b)
    ref = np.array([orc.find_nearest_centroid(v, cent, metric) for v in X[:400]], np.int32)
    np.testing.assert_array_equal(a[:400], ref)
    if metric == orc.L2:
        assert (a[50:60] == 7).all()


def test_tc_shortlist_kmeans_matches_exact_path(gpu, monkeypatch):
    rng = np.random.default_rng(5)
    X = rng.random((30_000, 64), dtype=np.float32)
    c1, it1 = gpu.kmeans_train(X, 2500, gpu.L2, 3, 42)
    monkeypatch.setenv("PYROPE_ASSIGN_EXACT", "1")
    c2, it2 = gpu.kmeans_train(X, 2500, gpu.L2, 3, 42)
    assert it1 == it2
    np.testing.assert_array_equal(c1, c2)


def test_lm_edge_shapes_empty_lists_and_short_results(gpu):
    """nprobe > nlist, lists left empty by k-means, k larger than everything the probes hold, and a batch that
    mixes all of it: result counts and contents must follow the reference (count = min(k, scored))."""
    rng = np.random.default_rng(9)
    # 3 tight blobs, nlist = 8: several centroids collapse onto the same points -> empty lists after assignment
    centers = rng.random((3, 128), dtype=np.float32)
    base = np.concatenate([c + np.float32(1e-3) * rng.standard_normal((150, 128)).astype(np.float32) for c in centers])
    q = np.concatenate([base[:40] + np.float32(1e-3), rng.random((40, 128), dtype=np.float32)]).astype(np.float32)
    ref, ix = _pair(gpu, base, 128, 8)
    sizes = np.diff(ix.lists()[0])
    for k, nprobe in ((10, 1), (10, 64), (500, 2), (1000, 8)):
        rid, rsc, rcn = ref.search_batch(q, k, nprobe=nprobe)
        got = _s(ix, q, k, nprobe=nprobe)
        assert_batch_equivalent((rid, rsc, rcn), got, ctx=f"lm edge k={k} nprobe={nprobe} sizes={sizes.tolist()}")
        assert (got[2] <= min(k, len(base))).all()


def test_lm_repeated_searches_reuse_scratch(gpu, mid):
    """Different batch sizes / k / nprobe back to back on one handle: per-search scratch (pools, thresholds,
    item blocks) must be re-initialised every time."""
    base, q, ref, ix = mid
    for nq, k, nprobe in ((300, 10, 8), (64, 100, 2), (200, 1, 16), (300, 10, 8)):
        assert_batch_equivalent(ref.search_batch(q[:nq], k, nprobe=nprobe), _s(ix, q[:nq], k, nprobe=nprobe),
                                ctx=f"lm repeat nq={nq} k={k} nprobe={nprobe}")


def test_lm_fixed_point_tables_adversarial(gpu):
    """The scan works on 16-bit fixed-point tables (8 queries per lookup) and re-scores everything within the
    rounding band of the k-th best.  Stress the band: thousands of duplicated vectors (identical codes, i.e. exact
    ties that overflow the per-pair pool regions and go through the redo path), near-duplicates separated by less
    than one quantisation step, rows of very different magnitude (a coarse scale for some queries) and queries
    far outside the data."""
    rng = np.random.default_rng(2024)
    dim = 128
    base = rng.random((12_000, dim), dtype=np.float32)
    base[1000:3000] = base[999]                                            # 2,000 exact duplicates
    base[3000:3600] = base[2999] + rng.random((600, dim), dtype=np.float32) * 2e-4   # inside one quantisation step
    base[6000:6500] *= 40.0                                                # large rows: large residuals, coarse scale
    ref, ix = _pair(gpu, base, dim, 12)
    q = rng.random((96, dim), dtype=np.float32)
    q[:10] = base[999] + 1e-3      # next to the duplicates
    q[10:20] = base[2999] + 1e-4   # next to the near-duplicates
    q[20:26] *= 40.0               # next to the large rows
    q[26:30] = 500.0               # far from everything
    for k, nprobe in ((10, 12), (3, 4), (64, 12), (300, 6)):
        assert_batch_equivalent(ref.search_batch(q, k, nprobe=nprobe), _s(ix, q, k, nprobe=nprobe),
                                ctx=f"lm adversarial k={k} nprobe={nprobe}")
    # scores are the reference's values wherever the ids agree (ties may permute ids)
    rid, rsc, _ = ref.search_batch(q, 10, nprobe=12)
    gid, gsc, _ = _s(ix, q, 10, nprobe=12)
    same = rid == gid
    np.testing.assert_array_equal(rsc[same], gsc[same])


def test_lm_random_shape_sweep(gpu):
    """Random (n, nlist, nq, k, nprobe, deletes-before-build) draws against the oracle."""
    rng = np.random.default_rng(99)
    for trial in range(6):
        dim = int(rng.choice([64, 128]))
        n = int(rng.integers(3_000, 9_000))
        nlist = int(rng.integers(3, 40))
        nq = int(rng.integers(1, 200))
        k = int(rng.choice([1, 5, 10, 50, 128]))
        nprobe = int(rng.integers(1, nlist + 3))
        base = rng.standard_normal((n, dim)).astype(np.float32) * rng.choice([0.1, 1.0, 7.0])
        q = rng.standard_normal((nq, dim)).astype(np.float32)
        ref, ix = _pair(gpu, base, dim, nlist)
        assert_batch_equivalent(ref.search_batch(q, k, nprobe=nprobe), _s(ix, q, k, nprobe=nprobe),
                                ctx=f"lm sweep trial={trial} dim={dim} n={n} nlist={nlist} nq={nq} k={k} nprobe={nprobe}")


# ------------------------------------------------------------------------------------------------
# m < 16: the scan runs on the equivalent 16-table quantiser, the re-score on the index's own
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,m", [(128, 8), (128, 4), (64, 4), (64, 8), (64, 2), (128, 2), (64, 1)])
def test_lm_fewer_sub_quantisers(gpu, dim, m):
    """ProductQuantizer accepts any m that divides dim (ProductQuantizer.cs:18-19; m = 4 is the registry default,
    VectorIndexRegistry.cs:96-106).  Codes and codebooks stay [m]; the list-major kernels see every codeword cut
    into 16/m pieces.  Scores must still be the reference's (sub-vectors of 16..64 floats: L2SquaredUnsafe's
    remainder accumulator, and its four-accumulator block from 32 on)."""
    base = orc.random_vectors(8_000, dim, 31 + m)
    q = orc.random_vectors(150, dim, 32 + m)
    ref, ix = _pair(gpu, base, dim, 16, m=m)
    for k, nprobe in ((10, 4), (100, 2), (1, 16)):
        rid, rsc, rcn = ref.search_batch(q, k, nprobe=nprobe)
        got = _s(ix, q, k, nprobe=nprobe)
        assert ix.last_search_kernel()[0] == "ivfpq_lm_scan_kernel"
        assert_batch_equivalent((rid, rsc, rcn), got, ctx=f"lm dim={dim} m={m} k={k} nprobe={nprobe}")
        same = rid == got[0]
        if m > 1:  # one sub-quantiser = at most 256 distinct distances per list: rows tie by the hundred
            assert same.mean() > 0.99
        np.testing.assert_array_equal(rsc[same], got[1][same])
        np.testing.assert_array_equal(rsc, got[1])  # the score LISTS agree whatever the order among ties


def test_lm_fewer_sub_quantisers_follow_rebuilds_and_new_codebooks(gpu, monkeypatch):
    """The 16-table view is derived state: a Build that replaces the lists, and codebooks set by the caller, must
    both refresh it."""
    dim, m = 128, 4
    base = orc.random_vectors(6_000, dim, 77)
    q = orc.random_vectors(64, dim, 78)
    ref, ix = _pair(gpu, base, dim, 8, m=m)
    assert_batch_equivalent(ref.search_batch(q, 10, nprobe=4), _s(ix, q, 10, nprobe=4), ctx="m=4 first build")
    more = orc.random_vectors(3_000, dim, 79)
    # IvfPqVectorIndex.Build re-trains from the buffer only and replaces the lists (IvfPqVectorIndex.cs:64,92)
    ref.add_batch(more, ids=np.arange(6_000, 9_000))
    ref.build()
    ix.add(more)
    ix.build()
    assert_batch_equivalent(ref.search_batch(q, 10, nprobe=4), _s(ix, q, 10, nprobe=4), ctx="m=4 second build")
    # a frozen copy with the first index's codebooks
    cent = ix.centroids()
    cb, _ = ix.codebooks()
    fz = gpu.GpuIndex(gpu.IVF_PQ, dim, gpu.L2, nlist=8, m=m, k=256)
    fz.set_codebooks(cent, cb)
    fz.add(more)
    fz.build()
    a = _s(fz, q, 10, nprobe=4)
    cb2 = (cb * np.float32(0.9) + np.float32(0.01)).astype(np.float32)  # other codewords: other codes, other distances
    fz.set_codebooks(cent, cb2)
    fz.add(more)
    fz.build()
    b = _s(fz, q, 10, nprobe=4)
    assert not np.array_equal(a[1], b[1])
    assert fz.last_search_kernel()[0] == "ivfpq_lm_scan_kernel"
    monkeypatch.setenv("PYROPE_PQ_LM", "0")  # the query-major kernels read the index's own [m] tables
    qm = gpu.GpuIndex(gpu.IVF_PQ, dim, gpu.L2, nlist=8, m=m, k=256)
    qm.set_codebooks(cent, cb2)
    qm.add(more)
    qm.build()
    c = _s(qm, q, 10, nprobe=4)
    assert qm.last_search_kernel()[0] != "ivfpq_lm_scan_kernel"
    b = (b[0] - 3_000, b[1], b[2])  # fz holds the second copy of `more`: rows 3000..5999
    assert_batch_equivalent(c, b, ctx="m=4 new codebooks: query-major vs list-major")
