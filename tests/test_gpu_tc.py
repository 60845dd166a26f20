"""GPU parity tests of the tensor-core FLAT path (flat_tc.cu: tcgen05 3xTF32 GEMM + fused top-k + exact
fp32 re-score), forced on with PYROPE_FLAT_TC=1 so that small shapes exercise it too.  Same bar as
tests/test_gpu_parity.py: distances within 1e-4 relative of the oracle, ids identical modulo ties."""
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


@pytest.fixture()
def force_tc(monkeypatch):
    monkeypatch.setenv("PYROPE_FLAT_TC", "1")


def _s(ix, Q, k, **kw):
    sc, rows, cnt = ix.search(Q, k, **kw)
    return rows, sc, cnt


@pytest.mark.parametrize("metric", [orc.L2, orc.IP, orc.COSINE])
def test_c1_flat_tc_matches_oracle(gpu, force_tc, metric):
    base = orc.random_vectors(10_000, 128, 42)
    q = orc.random_vectors(100, 128, 1337)
    ref = orc.FlatIndex(128, metric)
    ref.add_batch(base)
    ix = gpu.GpuIndex(gpu.FLAT, 128, metric)
    ix.add(base)
    assert_batch_equivalent(ref.search_batch(q, 10), _s(ix, q, 10), ctx=f"TC C1 metric={metric}")
    # the tensor-core path issues: query split, (operand prepare), tc kernel, re-score, merge
    assert ix.last_search_launches() >= 4


def test_flat_tc_dim768_ip_top100(gpu, force_tc):
    """C4's shape (inner product, d=768, TOPK 100) at a size the oracle finishes in seconds."""
    rng = np.random.default_rng(5)
    base = rng.random((20_000, 768), dtype=np.float32)
    q = rng.random((300, 768), dtype=np.float32)
    ref = orc.FlatIndex(768, orc.IP)
    ref.add_batch(base)
    ix = gpu.GpuIndex(gpu.FLAT, 768, gpu.INNER_PRODUCT)
    ix.add(base)
    assert_batch_equivalent(ref.search_batch(q, 100), _s(ix, q, 100), ctx="TC d=768 k=100")


@pytest.mark.parametrize("n,nq,dim,k", [(257, 1, 64, 5), (1000, 129, 36, 1), (5000, 130, 200, 33), (255, 7, 8, 16)])
def test_flat_tc_ragged_shapes(gpu, force_tc, n, nq, dim, k):
    rng = np.random.default_rng(n + nq)
    base = (rng.random((n, dim), dtype=np.float32) - 0.5)
    q = (rng.random((nq, dim), dtype=np.float32) - 0.5)
    for metric in (orc.L2, orc.IP, orc.COSINE):
        ref = orc.FlatIndex(dim, metric)
        ref.add_batch(base)
        ix = gpu.GpuIndex(gpu.FLAT, dim, metric)
        ix.add(base)
        assert_batch_equivalent(ref.search_batch(q, k), _s(ix, q, k), ctx=f"TC n={n} nq={nq} d={dim} k={k} m={metric}")


def test_flat_tc_tombstones_updates_maxscans(gpu, force_tc):
    base = orc.random_vectors(6000, 128, 42)
    q = orc.random_vectors(40, 128, 1337)
    rng = np.random.default_rng(3)
    ref = orc.FlatIndex(128, orc.L2)
    ix = gpu.GpuIndex(gpu.FLAT, 128, gpu.L2)
    ref.add_batch(base[:5000])
    ix.add(base[:5000])
    assert_batch_equivalent(ref.search_batch(q, 10), _s(ix, q, 10), ctx="before")
    dels = rng.choice(5000, 700, replace=False)
    for r in dels:
        assert ref.delete(int(r)) and ix.delete_row(int(r))
    live = np.setdiff1d(np.arange(5000), dels)
    for r in rng.choice(live, 50, replace=False):
        v = rng.random(128, dtype=np.float32)
        ref.upsert(int(r), v)
        ix.update_row(int(r), v)
    ref.add_batch(base[5000:], ids=np.arange(5000, 6000))  # incremental add after a search
    ix.add(base[5000:])
    assert_batch_equivalent(ref.search_batch(q, 10), _s(ix, q, 10), ctx="after deletes/upserts/adds")
    for ms in (1, 300, 4300, 5300, 100000):
        assert_batch_equivalent(ref.search_batch(q[:8], 10, max_scans=ms), _s(ix, q[:8], 10, max_scans=ms),
                                ctx=f"TC max_scans={ms}")


def test_near_duplicates_keep_fp32_distances(gpu, force_tc):
    """3xTF32 + |x|^2-2q.x cancels badly for near-duplicates; the exact re-score must repair it."""
    rng = np.random.default_rng(11)
    base = rng.random((4096, 128), dtype=np.float32)
    q = base[:16] + np.float32(1e-4) * rng.standard_normal((16, 128)).astype(np.float32)
    base[100:116] = q + np.float32(3e-4) * rng.standard_normal((16, 128)).astype(np.float32)
    ref = orc.FlatIndex(128, orc.L2)
    ref.add_batch(base)
    ix = gpu.GpuIndex(gpu.FLAT, 128, gpu.L2)
    ix.add(base)
    assert_batch_equivalent(ref.search_batch(q, 5), _s(ix, q, 5), ctx="near duplicates")


def test_ivf_coarse_probe_on_tensor_cores(gpu, force_tc):
    base = orc.random_vectors(10_000, 128, 42)
    q = orc.random_vectors(100, 128, 1337)
    ref = orc.IvfFlatIndex(128, orc.L2, nlist=100)
    ref.add_batch(base)
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_FLAT, 128, gpu.L2, nlist=100)
    ix.add(base)
    ix.build()
    np.testing.assert_array_equal(ref.centroids(), ix.centroids())
    for nprobe in (-1, 1, 10, 100):
        assert_batch_equivalent(ref.search_batch(q, 10, nprobe=nprobe), _s(ix, q, 10, nprobe=nprobe),
                                ctx=f"TC coarse nprobe={nprobe}")
    refp = orc.IvfPqIndex(128, orc.L2, m=16, k=256, nlist=64)
    refp.add_batch(base[:6000])
    refp.build()
    ixp = gpu.GpuIndex(gpu.IVF_PQ, 128, gpu.L2, nlist=64, m=16, k=256)
    ixp.add(base[:6000])
    ixp.build()
    assert_batch_equivalent(refp.search_batch(q, 10, nprobe=8), _s(ixp, q, 10, nprobe=8), ctx="TC coarse + ADC")


def test_tc_agrees_with_cuda_core_path(gpu, monkeypatch):
    base = orc.random_vectors(30_000, 128, 42)
    q = orc.random_vectors(256, 128, 1337)
    monkeypatch.setenv("PYROPE_FLAT_TC", "0")
    a = gpu.GpuIndex(gpu.FLAT, 128, gpu.L2)
    a.add(base)
    ra = _s(a, q, 10)
    monkeypatch.setenv("PYROPE_FLAT_TC", "1")
    b = gpu.GpuIndex(gpu.FLAT, 128, gpu.L2)
    b.add(base)
    rb = _s(b, q, 10)
    assert_batch_equivalent(ra, rb, ctx="tc vs cuda-core")
    assert recall_at_k(ra[0], rb[0], 10) > 0.999


def test_two_pass_threshold_matches_oracle_and_one_pass(gpu, force_tc, monkeypatch):
    """n >= 16384, d <= 256, k' >= 32: group-maxima pass -> per-query threshold -> filtered pass."""
    rng = np.random.default_rng(17)
    base = rng.random((20_000, 128), dtype=np.float32)
    base[5000:5040] = base[100]          # 41 identical rows: more ties than k' at the threshold
    q = rng.random((150, 128), dtype=np.float32)
    q[0] = base[100]
    for metric in (orc.L2, orc.IP, orc.COSINE):
        ref = orc.FlatIndex(128, metric)
        ref.add_batch(base)
        want = ref.search_batch(q, 20)
        ix = gpu.GpuIndex(gpu.FLAT, 128, metric)
        ix.add(base)
        got = _s(ix, q, 20)
        assert ix.last_search_launches() >= 6   # split, operand prep, pass A, select, pass B, re-score, merge
        assert_batch_equivalent(want, got, ctx=f"two-pass metric={metric}")
        monkeypatch.setenv("PYROPE_TC_ONEPASS", "1")
        one = gpu.GpuIndex(gpu.FLAT, 128, metric)
        one.add(base)
        assert_batch_equivalent(_s(one, q, 20), got, ctx=f"one-pass vs two-pass metric={metric}")
        monkeypatch.delenv("PYROPE_TC_ONEPASS")


@pytest.mark.parametrize("nq,metric", [(60, orc.L2), (300, orc.IP)])
def test_flat_tc_two_pass_one_term_many_splits(gpu, force_tc, nq, metric):
    """The two-pass threshold path (d <= 256, >= 16,384 rows, k' >= 32) with few query tiles, i.e. many row
    splits x two column halves: pass A group maxima, one-TF32 pass B with band pruning, exact re-score of every
    survivor.  Also run with near-duplicate rows, where a whole cluster sits inside the rounding band."""
    rng = np.random.default_rng(77 + nq)
    base = rng.random((70_000, 32), dtype=np.float32)
    base[5000:5400] = base[4999] + rng.random((400, 32), dtype=np.float32) * 1e-4   # 400 rows within the band of each other
    q = rng.random((nq, 32), dtype=np.float32)
    q[:8] = base[4999] + 1e-3
    ref = orc.FlatIndex(32, metric)
    ref.add_batch(base)
    ix = gpu.GpuIndex(gpu.FLAT, 32, gpu.L2 if metric == orc.L2 else gpu.INNER_PRODUCT)
    ix.add(base)
    for k in (40, 100):
        assert_batch_equivalent(ref.search_batch(q, k), _s(ix, q, k), ctx=f"two-pass one-term nq={nq} k={k}")
    assert ix.last_search_kernel()[0].startswith("flat_tc_kernel")


# ------------------------------------------------------------------------------------------------
# long rows (d > 256, >= 65,536 rows): ONE product per K slice on fp16 copies + rigorous band + exact re-score
# ------------------------------------------------------------------------------------------------
def _flat_pair(gpu, metric, base):
    ref = orc.FlatIndex(base.shape[1], metric)
    ref.add_batch(base)
    ix = gpu.GpuIndex(gpu.FLAT, base.shape[1], gpu.L2 if metric == orc.L2 else gpu.INNER_PRODUCT)
    ix.add(base)
    return ref, ix


@pytest.mark.parametrize("metric", [orc.L2, orc.IP])
def test_flat_one_term_fp16_pass_matches_oracle(gpu, metric):
    rng = np.random.default_rng(310)
    base = rng.random((70_000, 320), dtype=np.float32)
    base[300:340] = base[299] + rng.random((40, 320), dtype=np.float32) * 1e-4   # a cluster inside one rounding band
    q = rng.random((150, 320), dtype=np.float32)
    q[:4] = base[299] + 1e-3
    ref, ix = _flat_pair(gpu, metric, base)
    for k in (10, 100):
        rid, rsc, rcn = ref.search_batch(q, k)
        got = _s(ix, q, k)
        assert ix.last_search_kernel()[0] == "flat_tc_kernel (1xFP16 + band)"
        assert_batch_equivalent((rid, rsc, rcn), got, ctx=f"one-term fp16 k={k}")
        same = rid == got[0]
        np.testing.assert_array_equal(rsc[same], got[1][same])   # scores come from the exact re-score


def test_flat_one_term_fp16_range_edges(gpu):
    rng = np.random.default_rng(311)
    # every value below fp16's smallest normal
    tiny = (rng.random((66_000, 264), dtype=np.float32) * np.float32(2e-5)).astype(np.float32)
    ref, ix = _flat_pair(gpu, orc.IP, tiny)
    q = (rng.random((40, 264), dtype=np.float32) * np.float32(2e-5)).astype(np.float32)
    assert_batch_equivalent(ref.search_batch(q, 20), _s(ix, q, 20), ctx="fp16 underflow")
    # magnitudes eight orders apart inside a row
    mag = np.where(np.arange(264) % 3 == 0, np.float32(1.5e4), np.float32(1e-4)).astype(np.float32)
    mixed = (rng.random((66_000, 264), dtype=np.float32) * mag).astype(np.float32)
    ref, ix = _flat_pair(gpu, orc.L2, mixed)
    q = (rng.random((40, 264), dtype=np.float32) * mag).astype(np.float32)
    assert_batch_equivalent(ref.search_batch(q, 20), _s(ix, q, 20), ctx="fp16 mixed magnitudes")
    # one query beyond fp16: its band is infinite, the three-term pass redoes the batch
    q[7, 0] = np.float32(2.0e5)
    assert_batch_equivalent(ref.search_batch(q, 20), _s(ix, q, 20), ctx="query beyond fp16")
    # a table value beyond fp16: the tf32 one-term pass is kept
    mixed[123, 3] = np.float32(1.0e6)
    ref, ix = _flat_pair(gpu, orc.L2, mixed)
    q = (rng.random((40, 264), dtype=np.float32) * mag).astype(np.float32)
    assert_batch_equivalent(ref.search_batch(q, 20), _s(ix, q, 20), ctx="table beyond fp16")
    assert ix.last_search_kernel()[0] == "flat_tc_kernel (1xTF32 + band)"


def test_flat_one_term_fp16_follows_appends_and_updates(gpu):
    """The fp16 copy is derived state: rows appended between searches are converted on their own, an in-place update
    starts it over.  Both must show up in the next search."""
    rng = np.random.default_rng(312)
    base = rng.random((66_000, 264), dtype=np.float32)
    ref, ix = _flat_pair(gpu, orc.IP, base)
    q = rng.random((64, 264), dtype=np.float32)
    assert_batch_equivalent(ref.search_batch(q, 10), _s(ix, q, 10), ctx="before appends")
    assert ix.last_search_kernel()[0] == "flat_tc_kernel (1xFP16 + band)"
    more = (q[:20] * np.float32(1.5)).astype(np.float32)          # the new best matches of the first 20 queries
    ref.add_batch(more, ids=np.arange(66_000, 66_020))
    ix.add(more)
    got = _s(ix, q, 10)
    assert_batch_equivalent(ref.search_batch(q, 10), got, ctx="after appends")
    assert (got[0][:20, 0] == np.arange(66_000, 66_020)).all()
    upd = (q[30] * np.float32(3.0)).astype(np.float32)             # an in-place update of an old row
    ref.upsert(5, upd)
    ix.update_row(5, upd)
    got = _s(ix, q, 10)
    assert_batch_equivalent(ref.search_batch(q, 10), got, ctx="after an update")
    assert got[0][30, 0] == 5
