"""GPU parity tests: the CUDA path, called through the C ABI (pyrope_b200._lib -> libpyrope_gpu.so),
against the CPU oracle on the same seeded inputs.  Configs C1-C3 are BASELINE.json's CPU-runnable
configs (synthetic data = System.Random seeds 42 / 1337, Pyrope.Benchmarks/Program.cs:225-263).

Bar (BASELINE.json north_star): distances within 1e-4 relative (RTOL in tests/parity.py); top-k id
lists identical except sub-tolerance ties; centroid assignments and PQ codes bit-exact given the
same codebooks — and here the device-side training is bit-exact too, so centroids/codebooks match.
"""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.parity import assert_batch_equivalent, recall_at_k

pytestmark = pytest.mark.gpu

N, NQ, DIM, K = 10_000, 100, 128, 10


@pytest.fixture(scope="module")
def gpu():
    import pyrope_b200 as pg
    pg.load()
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    return pg


@pytest.fixture(scope="module")
def data():
    base = orc.random_vectors(N, DIM, 42)
    queries = orc.random_vectors(NQ, DIM, 1337)
    return base, queries


def _gpu_search(ix, Q, k, **kw):
    sc, rows, cnt = ix.search(Q, k, **kw)
    return rows, sc, cnt


# ------------------------------------------------------------------------------------------ C1
@pytest.mark.parametrize("metric", [orc.L2, orc.IP, orc.COSINE])
def test_c1_flat_matches_oracle(gpu, data, metric):
    base, queries = data
    ref = orc.FlatIndex(DIM, metric)
    ref.add_batch(base)
    ix = gpu.GpuIndex(gpu.FLAT, DIM, metric)
    assert ix.add(base) == 0
    got = _gpu_search(ix, queries, K)
    assert_batch_equivalent(ref.search_batch(queries, K), got, ctx=f"C1 metric={metric}")
    assert ix.stats()["live"] == N
    assert ix.last_search_launches() >= 2


def test_c1_flat_matches_float64_bruteforce(gpu, data):
    """Independent of the oracle: exact top-10 under float64 arithmetic."""
    base, queries = data
    ix = gpu.GpuIndex(gpu.FLAT, DIM, gpu.L2)
    ix.add(base)
    sc, rows, cnt = ix.search(queries, K)
    b64, q64 = base.astype(np.float64), queries.astype(np.float64)
    d = ((q64 ** 2).sum(1)[:, None] - 2 * q64 @ b64.T + (b64 ** 2).sum(1)[None, :])
    truth = np.argsort(d, axis=1, kind="stable")[:, :K]
    assert (cnt == K).all()
    assert recall_at_k(truth, rows, K) == 1.0
    np.testing.assert_allclose(-sc, np.take_along_axis(d, truth, 1), rtol=1e-4)


def test_flat_topk_100_and_k_larger_than_n(gpu, data):
    base, queries = data
    ref = orc.FlatIndex(DIM, orc.IP)
    ref.add_batch(base[:3000])
    ix = gpu.GpuIndex(gpu.FLAT, DIM, gpu.INNER_PRODUCT)
    ix.add(base[:3000])
    assert_batch_equivalent(ref.search_batch(queries[:20], 100), _gpu_search(ix, queries[:20], 100), ctx="k=100")
    # k > n: every live row comes back
    ref2 = orc.FlatIndex(DIM, orc.L2)
    ref2.add_batch(base[:7])
    ix2 = gpu.GpuIndex(gpu.FLAT, DIM, gpu.L2)
    ix2.add(base[:7])
    got = _gpu_search(ix2, queries[:5], 16)
    assert (got[2] == 7).all()
    assert_batch_equivalent(ref2.search_batch(queries[:5], 16), got, ctx="k>n")
    # large k
    assert_batch_equivalent(ref.search_batch(queries[:3], 1000), _gpu_search(ix, queries[:3], 1000), ctx="k=1000")


@pytest.mark.parametrize("dim", [1, 2, 12, 37, 130])
def test_flat_odd_dimensions(gpu, dim):
    rng = np.random.default_rng(dim)
    base = rng.random((500, dim), dtype=np.float32)
    q = rng.random((9, dim), dtype=np.float32)
    for metric in (orc.L2, orc.IP, orc.COSINE):
        ref = orc.FlatIndex(dim, metric)
        ref.add_batch(base)
        ix = gpu.GpuIndex(gpu.FLAT, dim, metric)
        ix.add(base)
        assert_batch_equivalent(ref.search_batch(q, 5), _gpu_search(ix, q, 5), ctx=f"dim={dim} metric={metric}")


def test_flat_reference_unit_cases(gpu):
    """BruteForceVectorIndexTests.cs:10-65 through the C ABI."""
    ix = gpu.GpuIndex(gpu.FLAT, 2, gpu.COSINE)
    ix.add([[1, 0], [0, 1]])
    sc, rows, cnt = ix.search([1, 0.1], 1)
    assert cnt[0] == 1 and rows[0, 0] == 0
    ix = gpu.GpuIndex(gpu.FLAT, 2, gpu.INNER_PRODUCT)  # Upsert overwrites
    ix.add([1, 0])
    ix.update_row(0, [0, 2])
    sc, rows, cnt = ix.search([0, 1], 1)
    assert rows[0, 0] == 0 and sc[0, 0] > 1.0
    ix = gpu.GpuIndex(gpu.FLAT, 2, gpu.L2)  # Delete removes
    ix.add([1, 1])
    assert ix.delete_row(0) and not ix.delete_row(0)
    sc, rows, cnt = ix.search([1, 1], 1)
    assert cnt[0] == 0 and rows[0, 0] == -1
    with pytest.raises(gpu.PyropeGpuError, match="dimension"):  # wrong dimension
        ix.add([1.0])
    ix = gpu.GpuIndex(gpu.FLAT, 2, gpu.INNER_PRODUCT)  # MaxScans 0 -> empty
    ix.add([[1, 0], [0, 1]])
    sc, rows, cnt = ix.search([1, 0], 1, max_scans=0)
    assert cnt[0] == 0
    with pytest.raises(gpu.PyropeGpuError, match="topK"):  # topK <= 0 throws for FLAT
        ix.search([1, 0], 0)
    empty = gpu.GpuIndex(gpu.FLAT, 2, gpu.L2)  # empty index -> empty result, not an error
    sc, rows, cnt = empty.search([1, 0], 3)
    assert cnt[0] == 0


def test_flat_tombstones_upsert_and_maxscans(gpu, data):
    base, queries = data
    n = 2000
    rng = np.random.default_rng(7)
    ref = orc.FlatIndex(DIM, orc.L2)
    ix = gpu.GpuIndex(gpu.FLAT, DIM, gpu.L2)
    ref.add_batch(base[:n])
    ix.add(base[:n])
    dels = rng.choice(n, 300, replace=False)
    for r in dels:
        assert ref.delete(int(r)) and ix.delete_row(int(r))
    live = np.setdiff1d(np.arange(n), dels)
    for r in rng.choice(live, 40, replace=False):  # Upsert of a live id overwrites in place
        v = rng.random(DIM, dtype=np.float32)
        ref.upsert(int(r), v)
        ix.update_row(int(r), v)
    assert ix.stats()["live"] == ref.count() == n - 300
    assert_batch_equivalent(ref.search_batch(queries[:30], K), _gpu_search(ix, queries[:30], K), ctx="tombstones")
    for ms in (1, 5, 137, 1699, 1700, 5000):  # MaxScans counts live rows in insertion order
        assert_batch_equivalent(ref.search_batch(queries[:10], K, max_scans=ms),
                                _gpu_search(ix, queries[:10], K, max_scans=ms), ctx=f"max_scans={ms}")


# ------------------------------------------------------------------------------------------ C2
@pytest.fixture(scope="module")
def c2(gpu, data):
    base, queries = data
    ref = orc.IvfFlatIndex(DIM, orc.L2, nlist=100)
    ref.add_batch(base)
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_FLAT, DIM, gpu.L2, nlist=100)
    ix.add(base)
    ix.build()
    return ref, ix


def test_c2_ivfflat_training_and_assignment_bit_exact(c2):
    ref, ix = c2
    assert ix.is_built()
    np.testing.assert_array_equal(ref.centroids(), ix.centroids())  # k-means is bit-exact
    off, rows, _ = ix.lists()
    ref_lists = ref.lists()
    assert len(off) == len(ref_lists) + 1
    for c, ids in enumerate(ref_lists):
        np.testing.assert_array_equal(ids, rows[off[c]:off[c + 1]])  # same members, same order
    assert ix.stats() == {"live": N, "buffer": 0, "dim": DIM, "metric": 0}


def test_c2_ivfflat_search_matches_oracle(c2, data):
    ref, ix = c2
    _, queries = data
    assert_batch_equivalent(ref.search_batch(queries, K), _gpu_search(ix, queries, K), ctx="C2 default nprobe=3")
    for nprobe in (1, 7, 100, 1000):
        assert_batch_equivalent(ref.search_batch(queries[:25], K, nprobe=nprobe),
                                _gpu_search(ix, queries[:25], K, nprobe=nprobe), ctx=f"C2 nprobe={nprobe}")
    for ms in (0, 1, 50, 120, 100000):
        assert_batch_equivalent(ref.search_batch(queries[:10], K, max_scans=ms),
                                _gpu_search(ix, queries[:10], K, max_scans=ms), ctx=f"C2 max_scans={ms}")


def test_c2_recall_sanity(c2, data):
    ref, ix = c2
    base, queries = data
    flat = orc.FlatIndex(DIM, orc.L2)
    flat.add_batch(base)
    truth = flat.search_batch(queries, K)[0]
    sc, rows, cnt = ix.search(queries, K, nprobe=100)  # all lists probed == exact
    assert recall_at_k(truth, rows, K) == 1.0


@pytest.mark.parametrize("metric", [orc.IP, orc.COSINE])
def test_ivfflat_other_metrics_buffer_and_deletes(gpu, data, metric):
    base, queries = data
    n = 3000
    ref = orc.IvfFlatIndex(DIM, metric, nlist=20)
    ix = gpu.GpuIndex(gpu.IVF_FLAT, DIM, metric, nlist=20)
    ref.add_batch(base[:n])
    ix.add(base[:n])
    # search before build: exact scan of the buffer (IvfFlatVectorIndexTests.cs:52-66)
    assert_batch_equivalent(ref.search_batch(queries[:10], K), _gpu_search(ix, queries[:10], K), ctx="pre-build")
    ref.build()
    ix.build()
    np.testing.assert_array_equal(ref.centroids(), ix.centroids())
    # new rows after the build sit in the buffer and are searched exactly
    ref.add_batch(base[n:n + 200], ids=np.arange(n, n + 200))
    first = ix.add(base[n:n + 200])
    assert first == n
    for r in (5, 17, 2999, n + 3, n + 150):  # deletes in lists and in the buffer
        assert ref.delete(r) and ix.delete_row(r)
    assert ix.stats()["live"] == ref.count()
    assert_batch_equivalent(ref.search_batch(queries[:30], K, nprobe=4), _gpu_search(ix, queries[:30], K, nprobe=4),
                            ctx=f"metric={metric} buffer+lists")
    assert_batch_equivalent(ref.search_batch(queries[:10], K, nprobe=4, max_scans=300),
                            _gpu_search(ix, queries[:10], K, nprobe=4, max_scans=300), ctx="max_scans after deletes")
    # rebuild folds the buffer into the lists
    ref.build()
    ix.build()
    np.testing.assert_array_equal(ref.centroids(), ix.centroids())
    assert_batch_equivalent(ref.search_batch(queries[:30], K), _gpu_search(ix, queries[:30], K), ctx="after rebuild")


def test_ivfflat_reference_unit_cases(gpu):
    """IvfFlatVectorIndexTests.cs through the C ABI."""
    ix = gpu.GpuIndex(gpu.IVF_FLAT, 2, gpu.L2, nlist=2)
    ix.add([[1, 0]])
    assert ix.centroids() is None  # GetCentroids before build -> null
    ix = gpu.GpuIndex(gpu.IVF_FLAT, 2, gpu.L2, nlist=2)
    ix.add([[0.1, 0.1], [0.2, 0.2], [10.1, 10.1], [10.2, 10.2]])
    ix.build()
    c = ix.centroids()
    assert c is not None and c.shape == (2, 2)
    sc, rows, cnt = ix.search([0, 0], 2)
    assert cnt[0] == 2 and set(rows[0].tolist()) == {0, 1}  # the two 'a' points (Build_ClustersData :69-90)
    ix = gpu.GpuIndex(gpu.IVF_FLAT, 2, gpu.L2, nlist=3)
    ix.add([[0, 0], [5, 5], [10, 10]])
    ix.build()
    sc, rows, cnt = ix.search([0, 0], 3, nprobe=3)
    assert cnt[0] == 3
    empty = gpu.GpuIndex(gpu.IVF_FLAT, 2, gpu.L2)
    sc, rows, cnt = empty.search([0, 0], 1)
    assert cnt[0] == 0
    sc, rows, cnt = ix.search([0, 0], 0)  # IVF does not validate topK: empty result
    assert cnt[0] == 0


# ------------------------------------------------------------------------------------------ C3
@pytest.fixture(scope="module")
def c3(gpu, data):
    base, queries = data
    ref = orc.IvfPqIndex(DIM, orc.L2, m=4, k=256, nlist=100)
    ref.add_batch(base)
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_PQ, DIM, gpu.L2, nlist=100, m=4, k=256)
    ix.add(base)
    ix.build()
    return ref, ix


def test_c3_ivfpq_codebooks_assignments_codes_bit_exact(c3):
    ref, ix = c3
    np.testing.assert_array_equal(ref.centroids(), ix.centroids())
    cb, ks = ix.codebooks()
    np.testing.assert_array_equal(ref.pq().codebook(), cb)
    assert ks.tolist() == ref.pq().ksub()
    off, rows, codes = ix.lists()
    for c, (ids, rcodes) in enumerate(ref.lists()):
        np.testing.assert_array_equal(ids, rows[off[c]:off[c + 1]])
        np.testing.assert_array_equal(rcodes, codes[off[c]:off[c + 1]])


def test_c3_ivfpq_search_matches_oracle(c3, data):
    ref, ix = c3
    _, queries = data
    assert_batch_equivalent(ref.search_batch(queries, K), _gpu_search(ix, queries, K), ctx="C3 default nprobe=1")
    for nprobe in (3, 16, 100):
        assert_batch_equivalent(ref.search_batch(queries[:25], K, nprobe=nprobe),
                                _gpu_search(ix, queries[:25], K, nprobe=nprobe), ctx=f"C3 nprobe={nprobe}")
    assert_batch_equivalent(ref.search_batch(queries[:5], 100, nprobe=8), _gpu_search(ix, queries[:5], 100, nprobe=8),
                            ctx="C3 k=100")


def test_c3_generic_kernel_agrees_with_fast_kernel(gpu, data, c3, monkeypatch):
    base, queries = data
    ref, fast = c3
    monkeypatch.setenv("PYROPE_PQ_GENERIC", "1")
    ix = gpu.GpuIndex(gpu.IVF_PQ, DIM, gpu.L2, nlist=100, m=4, k=256)
    ix.set_codebooks(fast.centroids(), fast.codebooks()[0])
    ix.add(base)
    ix.build()
    off_a, rows_a, codes_a = fast.lists()
    off_b, rows_b, codes_b = ix.lists()
    np.testing.assert_array_equal(off_a, off_b)
    np.testing.assert_array_equal(codes_a, codes_b)  # frozen codebooks -> identical assignment+codes
    assert_batch_equivalent(ref.search_batch(queries[:40], K, nprobe=5), _gpu_search(ix, queries[:40], K, nprobe=5),
                            ctx="generic kernel")


@pytest.mark.parametrize("m", [8, 16, 32, 2])
def test_ivfpq_other_m(gpu, data, m):
    base, queries = data
    n = 4000
    ref = orc.IvfPqIndex(DIM, orc.L2, m=m, k=256, nlist=16)
    ref.add_batch(base[:n])
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_PQ, DIM, gpu.L2, nlist=16, m=m, k=256)
    ix.add(base[:n])
    ix.build()
    cb, ks = ix.codebooks()
    np.testing.assert_array_equal(ref.pq().codebook(), cb)
    off, rows, codes = ix.lists()
    for c, (ids, rcodes) in enumerate(ref.lists()):
        np.testing.assert_array_equal(rcodes, codes[off[c]:off[c + 1]])
    assert_batch_equivalent(ref.search_batch(queries[:20], K, nprobe=4), _gpu_search(ix, queries[:20], K, nprobe=4),
                            ctx=f"m={m}")


def test_ivfpq_reference_unit_case(gpu):
    """IvfPqVectorIndexTests.cs:41-67: dim 128, m=16, k=256, nlist=4, 100 Random(123) vectors."""
    dim = 128
    base = orc.random_vectors(100, dim, 123)
    ref = orc.IvfPqIndex(dim, orc.L2, m=16, k=256, nlist=4)
    ref.add_batch(base)
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_PQ, dim, gpu.L2, nlist=4, m=16, k=256)
    ix.add(base)
    ix.build()
    cb, ks = ix.codebooks()
    assert ks.tolist() == [100] * 16  # K clipped to the number of training points
    np.testing.assert_array_equal(ref.pq().codebook(), cb)
    q = np.full((1, dim), 0.5, np.float32)
    sc, rows, cnt = ix.search(q, 5)
    assert cnt[0] == 5
    assert_batch_equivalent(ref.search_batch(q, 5), (rows, sc, cnt), ctx="IvfPq unit case")
    # buffer rows after the build are scored exactly with the metric and merged with ADC results
    extra = orc.random_vectors(10, dim, 5)
    ref.add_batch(extra, ids=np.arange(100, 110))
    ix.add(extra)
    assert_batch_equivalent(ref.search_batch(q, 8, nprobe=4), _gpu_search(ix, q, 8, nprobe=4), ctx="buffer+adc")
    # IVF_PQ Delete only touches the buffer (IvfPqVectorIndex.cs:48-53)
    assert ix.delete_row(103) and ref.delete(103)
    assert not ix.delete_row(5) and not ref.delete(5)
    assert_batch_equivalent(ref.search_batch(q, 8, nprobe=4), _gpu_search(ix, q, 8, nprobe=4), ctx="after delete")


# ------------------------------------------------------------------------------------------ building blocks
def test_building_blocks_bit_exact(gpu, data):
    base, queries = data
    cent, _ = orc.kmeans_train(base[:2000], 37, orc.L2, 10, 42)
    gcent, _ = gpu.kmeans_train(base[:2000], 37, gpu.L2, 10, 42)
    np.testing.assert_array_equal(cent, gcent)
    for metric in (orc.L2, orc.IP, orc.COSINE):
        a = gpu.coarse_assign(base[:1500], cent, metric)
        b = np.array([orc.find_nearest_centroid(v, cent, metric) for v in base[:1500]], np.int32)
        np.testing.assert_array_equal(a, b)
    pq = orc.ProductQuantizer(DIM, 16, 256)
    pq.train(base[:3000] - 0.5)
    cb = pq.codebook()
    codes = gpu.pq_encode(cb, base[3000:3500] - 0.5)
    for i in range(500):
        np.testing.assert_array_equal(codes[i], pq.encode(base[3000 + i] - 0.5))
    tab = gpu.pq_distance_table(cb, queries[:4] - 0.5)
    for i in range(4):
        np.testing.assert_allclose(tab[i], pq.distance_table(queries[i] - 0.5), rtol=1e-5, atol=1e-7)


def test_device_entry_points_and_merge(gpu, data):
    torch = pytest.importorskip("torch")
    base, queries = data
    ix = gpu.GpuIndex(gpu.FLAT, DIM, gpu.L2)
    ix.add(base[:5000])
    ix2 = gpu.GpuIndex(gpu.FLAT, DIM, gpu.L2)
    ix2.add(base[5000:], labels=np.arange(5000, N))
    q = torch.from_numpy(queries).cuda()
    outs = []
    for h in (ix, ix2):
        s = torch.empty((NQ, K), dtype=torch.float32, device="cuda")
        r = torch.empty((NQ, K), dtype=torch.int64, device="cuda")
        c = torch.empty((NQ,), dtype=torch.int32, device="cuda")
        h.search_device(q.data_ptr(), NQ, K, s.data_ptr(), r.data_ptr(), c.data_ptr(),
                        stream=torch.cuda.current_stream().cuda_stream)
        outs.append((s, r))
    torch.cuda.synchronize()
    S = torch.stack([o[0] for o in outs]).contiguous()
    R = torch.stack([o[1] for o in outs]).contiguous()
    ms = torch.empty((NQ, K), dtype=torch.float32, device="cuda")
    mr = torch.empty((NQ, K), dtype=torch.int64, device="cuda")
    mc = torch.empty((NQ,), dtype=torch.int32, device="cuda")
    gpu._lib.topk_merge_device(NQ, 2, K, K, S.data_ptr(), R.data_ptr(), ms.data_ptr(), mr.data_ptr(), mc.data_ptr(),
                               stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = orc.FlatIndex(DIM, orc.L2)
    ref.add_batch(base)
    assert_batch_equivalent(ref.search_batch(queries, K), (mr.cpu().numpy(), ms.cpu().numpy(), mc.cpu().numpy()),
                            ctx="two shards merged")


# ---------------------------------------------------------------- SQ8 FLAT (BruteForceVectorIndex.EnableQuantization)
def _sq8_check(ref, got_rows, got_sc, got_cnt, Q, k, max_scans=None, ctx=""):
    """Integer scores tie a lot: compare the score lists exactly and check that every returned row really has the
    score it is listed with (ties may permute / swap ids at the boundary)."""
    for qi in range(Q.shape[0]):
        sc = ref.scores(Q[qi], max_scans)
        by_id = dict(sc)
        want = sorted((s for _, s in sc), reverse=True)[:k]
        n = int(got_cnt[qi])
        assert n == len(want), f"{ctx} q{qi}: count {n} != {len(want)}"
        np.testing.assert_array_equal(np.asarray(want, np.float32), got_sc[qi, :n], err_msg=f"{ctx} q{qi}")
        for j in range(n):
            assert by_id[int(got_rows[qi, j])] == got_sc[qi, j], f"{ctx} q{qi} rank {j}"


@pytest.mark.parametrize("metric", ["L2", "IP", "COSINE"])
def test_sq8_flat_matches_oracle(gpu, metric):
    from oracle import sq8_oracle as sq
    rng = np.random.default_rng(31)
    dim, n = 100, 5000                                   # dim not a multiple of 16: zero padding
    base = (rng.random((n, dim), dtype=np.float32) - 0.3) * 4
    base[17] = 0.25                                      # a flat vector: all bytes 0
    Q = (rng.random((37, dim), dtype=np.float32) - 0.3) * 4
    gm = {"L2": gpu.L2, "IP": gpu.INNER_PRODUCT, "COSINE": gpu.COSINE}[metric]
    ix = gpu.GpuIndex(gpu.FLAT, dim, gm)
    ref = sq.Sq8FlatIndex(dim, metric)
    ix.set_quantization(True)
    ix.add(base[:3000])
    for i in range(3000):
        ref.add(i, base[i])
    # rows written while the flag is off have no quantised form: counted by MaxScans, never returned
    ix.set_quantization(False); ref.enable = False
    ix.add(base[3000:3500])
    for i in range(3000, 3500):
        ref.add(i, base[i])
    ix.set_quantization(True); ref.enable = True
    ix.add(base[3500:])
    for i in range(3500, n):
        ref.add(i, base[i])
    for i in (5, 1000, 3100, 4999):
        assert ix.delete_row(i) and ref.delete(i)
    ix.update_row(42, base[4242]); ref.upsert(42, base[4242])
    ix.update_row(3200, base[1]); ref.upsert(3200, base[1])   # gains a quantised form now
    for k, ms in ((10, None), (1, None), (100, None), (10, 3300), (10, 0)):
        sc, rows, cnt = ix.search(Q, k, max_scans=-1 if ms is None else ms)
        _sq8_check(ref, rows, sc, cnt, Q, k, ms, ctx=f"sq8 {metric} k={k} max_scans={ms}")
        if ms != 0:
            assert ix.last_search_kernel()[0] == "sq8_scan_kernel"


def test_sq8_bytes_bit_exact(gpu):
    """ScalarQuantizer.Quantize on device, checked through distances to one-hot probes is indirect; check the
    bytes directly instead: a FLAT L2 search of the exact byte pattern of row r must return r with score 0."""
    from oracle import sq8_oracle as sq
    rng = np.random.default_rng(5)
    base = rng.standard_normal((300, 48)).astype(np.float32)
    ix = gpu.GpuIndex(gpu.FLAT, 48, gpu.L2)
    ix.set_quantization(True)
    ix.add(base)
    # querying with a row itself quantises to the same bytes: distance 0 exactly, and every other score must equal
    # the oracle's integer distance between the two byte vectors
    sc, rows, cnt = ix.search(base[:20], 5)
    qb = [sq.quantize(r)[0] for r in base]
    for i in range(20):
        assert sc[i, 0] == 0.0
        for j in range(5):
            assert -sq.l2sq_8bit(qb[i], qb[int(rows[i, j])]) == sc[i, j]
