"""GPU parity tests of the list-major IVF_FLAT scan (ivf_lm.cu: pairs grouped by list, 16 queries per pass
over a list, butterfly reduce-scatter, exact re-score in the reference's order).  Batches of >= 64 queries
take it; the query-major kernel (ivf.cu) is the on-GPU cross-check."""
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


def _pair(gpu, base, dim, nlist, metric_o, metric_g):
    ref = orc.IvfFlatIndex(dim, metric_o, nlist=nlist)
    ref.add_batch(base)
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_FLAT, dim, metric_g, nlist=nlist)
    ix.add(base)
    ix.build()
    np.testing.assert_array_equal(ref.centroids(), ix.centroids())
    return ref, ix


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_ivfflat_lm_matches_oracle_bit_for_bit(gpu, metric):
    mo, mg = (orc.L2, gpu.L2) if metric == "l2" else (orc.IP, gpu.INNER_PRODUCT)
    base = orc.random_vectors(12_000, 128, 42)
    q = orc.random_vectors(200, 128, 1337)
    ref, ix = _pair(gpu, base, 128, 32, mo, mg)
    for k, nprobe in ((10, 3), (10, 8), (100, 4), (1, 1), (33, 32)):
        rid, rsc, rcn = ref.search_batch(q, k, nprobe=nprobe)
        gid, gsc, gcn = _s(ix, q, k, nprobe=nprobe)
        assert_batch_equivalent((rid, rsc, rcn), (gid, gsc, gcn), ctx=f"ivf lm {metric} k={k} nprobe={nprobe}")
        same = rid == gid
        assert same.mean() > 0.99
        np.testing.assert_array_equal(rsc[same], gsc[same])   # survivors re-scored in the reference's order
    assert ix.last_search_launches() >= 10


def test_ivfflat_lm_long_lists_small_dim_and_overflow(gpu):
    base = orc.random_vectors(16_000, 64, 5)
    q = orc.random_vectors(70, 64, 6)
    ref, ix = _pair(gpu, base, 64, 4, orc.L2, gpu.L2)      # ~4000 rows per list
    for k, nprobe in ((10, 4), (300, 2), (1000, 4)):        # k = 300 / 1000: queues overflow -> redo path
        assert_batch_equivalent(ref.search_batch(q, k, nprobe=nprobe), _s(ix, q, k, nprobe=nprobe),
                                ctx=f"ivf lm long k={k} nprobe={nprobe}")


def test_ivfflat_lm_deletes_buffer_and_query_major_agree(gpu, monkeypatch):
    base = orc.random_vectors(9_000, 128, 21)
    q = orc.random_vectors(96, 128, 22)
    ref, ix = _pair(gpu, base, 128, 16, orc.L2, gpu.L2)
    rng = np.random.default_rng(1)
    dels = [int(r) for r in rng.choice(9_000, 300, replace=False)]
    for r in dels:
        assert ref.delete(r) and ix.delete_row(r)
    extra = orc.random_vectors(50, 128, 23)
    ref.add_batch(extra, ids=np.arange(9_000, 9_050))
    ix.add(extra)
    got = _s(ix, q, 10, nprobe=5)
    assert_batch_equivalent(ref.search_batch(q, 10, nprobe=5), got, ctx="ivf lm deletes + buffer")
    # MaxScans budgets keep the reference's rank-order walk: they take the query-major kernel
    assert_batch_equivalent(ref.search_batch(q, 10, nprobe=5, max_scans=700), _s(ix, q, 10, nprobe=5, max_scans=700),
                            ctx="ivf budget")
    monkeypatch.setenv("PYROPE_PQ_LM", "0")
    qm = gpu.GpuIndex(gpu.IVF_FLAT, 128, gpu.L2, nlist=16)
    qm.set_codebooks(ix.centroids())
    qm.add(base)
    qm.build()
    for r in dels:
        assert qm.delete_row(r)
    qm.add(extra)
    other = _s(qm, q, 10, nprobe=5)
    assert qm.last_search_launches() < 10
    assert_batch_equivalent(other, got, ctx="query-major vs list-major")
    assert recall_at_k(other[0], got[0], 10) > 0.999


def test_ivfflat_lm_edge_shapes(gpu):
    rng = np.random.default_rng(10)
    centers = rng.random((3, 64), dtype=np.float32)
    base = np.concatenate([c + np.float32(1e-3) * rng.standard_normal((150, 64)).astype(np.float32) for c in centers])
    q = np.concatenate([base[:40] + np.float32(1e-3), rng.random((40, 64), dtype=np.float32)]).astype(np.float32)
    ref, ix = _pair(gpu, base, 64, 8, orc.L2, gpu.L2)
    for k, nprobe in ((10, 1), (10, 64), (400, 2), (1000, 8)):
        got = _s(ix, q, k, nprobe=nprobe)
        assert_batch_equivalent(ref.search_batch(q, k, nprobe=nprobe), got, ctx=f"ivf lm edge k={k} nprobe={nprobe}")
    for nq, k, nprobe in ((80, 10, 3), (64, 50, 2), (80, 10, 3)):   # scratch re-initialised between searches
        assert_batch_equivalent(ref.search_batch(q[:nq], k, nprobe=nprobe), _s(ix, q[:nq], k, nprobe=nprobe),
                                ctx=f"ivf lm repeat nq={nq} k={k}")


def test_ivf_lm_cosine_matches_oracle(gpu):
    """Cosine on the list-major IVF_FLAT scan: rows are ranked by q.x / |x| on the way, the survivors re-scored with
    VectorMath.Cosine on the stored norms (IvfFlatVectorIndex.cs:351-360) — incl. a zero row and a zero query."""
    rng = np.random.default_rng(17)
    base = (rng.random((12_000, 64), dtype=np.float32) - 0.5)
    base[123] = 0.0
    Q = (rng.random((200, 64), dtype=np.float32) - 0.5)
    Q[7] = 0.0
    ref = orc.IvfFlatIndex(64, orc.COSINE, nlist=24)
    ref.add_batch(base)
    ref.build()
    ix = gpu.GpuIndex(gpu.IVF_FLAT, 64, gpu.COSINE, nlist=24)
    ix.add(base)
    ix.build()
    np.testing.assert_array_equal(ref.centroids(), ix.centroids())
    for k, nprobe in ((10, 6), (100, 3)):
        sc, rows, cnt = ix.search(Q, k, nprobe=nprobe)
        assert ix.last_search_kernel()[0] == "ivf_lm_scan_kernel"
        assert_batch_equivalent(ref.search_batch(Q, k, nprobe=nprobe), (rows, sc, cnt), ctx=f"ivf lm cosine k={k} nprobe={nprobe}")


# ------------------------------------------------------------------------------------------------
# rows wider than 128 floats: queries staged in shared memory, 128-dimension chunks (ivf_lm_scan_wide_kernel)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,metric", [(256, "l2"), (768, "ip"), (136, "l2"), (384, "cos"), (1024, "l2")])
def test_ivfflat_lm_wide_rows(gpu, dim, metric):
    mo, mg = {"l2": (orc.L2, gpu.L2), "ip": (orc.IP, gpu.INNER_PRODUCT), "cos": (orc.COSINE, gpu.COSINE)}[metric]
    base = orc.random_vectors(6_000, dim, 50 + dim)
    q = orc.random_vectors(130, dim, 51 + dim)
    ref, ix = _pair(gpu, base, dim, 16, mo, mg)
    for k, nprobe in ((10, 3), (100, 6), (1, 16), (600, 2)):   # k = 600 > the seed sample: cold start + redo path
        rid, rsc, rcn = ref.search_batch(q, k, nprobe=nprobe)
        gid, gsc, gcn = _s(ix, q, k, nprobe=nprobe)
        assert ix.last_search_kernel()[0] == "ivf_lm_scan_kernel"
        assert_batch_equivalent((rid, rsc, rcn), (gid, gsc, gcn), ctx=f"ivf lm wide d={dim} {metric} k={k} nprobe={nprobe}")
        same = rid == gid
        assert same.mean() > 0.99
        np.testing.assert_array_equal(rsc[same], gsc[same])   # survivors re-scored in the reference's order


def test_ivfflat_lm_wide_rows_with_deletes_and_ragged_lists(gpu):
    dim = 320
    base = orc.random_vectors(5_000, dim, 77)
    q = orc.random_vectors(90, dim, 78)
    ref, ix = _pair(gpu, base, dim, 24, orc.L2, gpu.L2)     # ~208 rows per list: blocks of 32 rows end ragged, odd lengths
    rng = np.random.default_rng(3)
    for r in (int(x) for x in rng.choice(5_000, 200, replace=False)):
        assert ref.delete(r) and ix.delete_row(r)
    assert_batch_equivalent(ref.search_batch(q, 10, nprobe=8), _s(ix, q, 10, nprobe=8), ctx="ivf lm wide deletes")
