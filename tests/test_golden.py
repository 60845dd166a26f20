"""Committed golden vectors (tests/golden/configs_1_2_3.npz, made by tests/golden/make_golden.py) for
BASELINE.json configs 1-3: the oracle must keep reproducing them (CPU, always run) and the CUDA path must
match them through the C ABI (GPU)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.parity import assert_batch_equivalent

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "configs_1_2_3.npz"))
N, NQ, DIM, K = 10_000, 100, 128, 10


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def data():
    return orc.random_vectors(N, DIM, 42), orc.random_vectors(NQ, DIM, 1337)


def test_inputs_are_the_reference_generator(data):
    base, q = data
    assert sha(base) == str(G["base_sha256"]) and sha(q) == str(G["query_sha256"])
    np.testing.assert_array_equal(base[:2, :8], G["base_head"])


def test_oracle_reproduces_c1(data):
    base, q = data
    for name, metric in (("l2", orc.L2), ("ip", orc.IP), ("cos", orc.COSINE)):
        ix = orc.FlatIndex(DIM, metric)
        ix.add_batch(base)
        ids, sc, cnt = ix.search_batch(q[:25], K)
        np.testing.assert_array_equal(ids, G[f"c1_{name}_ids"][:25])
        np.testing.assert_array_equal(sc, G[f"c1_{name}_scores"][:25])


def test_oracle_reproduces_c2_c3_build(data):
    base, q = data
    ivf = orc.IvfFlatIndex(DIM, orc.L2, nlist=100)
    ivf.add_batch(base)
    ivf.build()
    assert sha(ivf.centroids()) == str(G["c2_centroids_sha256"])
    ids, sc, cnt = ivf.search_batch(q, K, nprobe=3)
    np.testing.assert_array_equal(ids, G["c2_np3_ids"])
    pq = orc.IvfPqIndex(DIM, orc.L2, m=4, k=256, nlist=100)
    pq.add_batch(base)
    pq.build()
    assert sha(pq.pq().codebook()) == str(G["c3_codebook_sha256"])
    assert sha(np.concatenate([c for _, c in pq.lists()])) == str(G["c3_codes_sha256"])
    ids, sc, cnt = pq.search_batch(q, K, nprobe=1)
    np.testing.assert_array_equal(ids, G["c3_np1_ids"])
    np.testing.assert_array_equal(sc, G["c3_np1_scores"])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    return pg


def _full(ids, sc):
    return ids, sc, np.full(len(ids), ids.shape[1], np.int32)


def _s(ix, Q, k, **kw):
    sc, rows, cnt = ix.search(Q, k, **kw)
    return rows, sc, cnt


@pytest.mark.gpu
def test_gpu_matches_golden_c1(gpu, data):
    base, q = data
    for name, metric in (("l2", gpu.L2), ("ip", gpu.INNER_PRODUCT), ("cos", gpu.COSINE)):
        ix = gpu.GpuIndex(gpu.FLAT, DIM, metric)
        ix.add(base)
        assert_batch_equivalent(_full(G[f"c1_{name}_ids"], G[f"c1_{name}_scores"]), _s(ix, q, K), ctx=f"golden C1 {name}")


@pytest.mark.gpu
def test_gpu_matches_golden_c2(gpu, data):
    base, q = data
    ix = gpu.GpuIndex(gpu.IVF_FLAT, DIM, gpu.L2, nlist=100)
    ix.add(base)
    ix.build()
    assert sha(ix.centroids()) == str(G["c2_centroids_sha256"])          # training bit-exact
    off, rows, _ = ix.lists()
    assign = np.empty(N, np.int32)
    for c in range(100):
        assign[rows[off[c]:off[c + 1]]] = c
    np.testing.assert_array_equal(assign, G["c2_assign"])                # assignment bit-exact
    for nprobe in (3, 10):
        assert_batch_equivalent(_full(G[f"c2_np{nprobe}_ids"], G[f"c2_np{nprobe}_scores"]), _s(ix, q, K, nprobe=nprobe),
                                ctx=f"golden C2 nprobe={nprobe}")


@pytest.mark.gpu
def test_gpu_matches_golden_c3_and_m16(gpu, data):
    base, q = data
    ix = gpu.GpuIndex(gpu.IVF_PQ, DIM, gpu.L2, nlist=100, m=4, k=256)
    ix.add(base)
    ix.build()
    cb, _ = ix.codebooks()
    assert sha(cb) == str(G["c3_codebook_sha256"])
    off, rows, codes = ix.lists()
    np.testing.assert_array_equal(np.diff(off).astype(np.int32), G["c3_list_sizes"])
    assert sha(codes) == str(G["c3_codes_sha256"]) and sha(rows) == str(G["c3_ids_sha256"])   # codes bit-exact
    for nprobe in (1, 8):
        assert_batch_equivalent(_full(G[f"c3_np{nprobe}_ids"], G[f"c3_np{nprobe}_scores"]), _s(ix, q, K, nprobe=nprobe),
                                ctx=f"golden C3 nprobe={nprobe}")
    ix16 = gpu.GpuIndex(gpu.IVF_PQ, DIM, gpu.L2, nlist=32, m=16, k=256)
    ix16.add(base)
    ix16.build()
    assert sha(ix16.lists()[2]) == str(G["m16_codes_sha256"])
    assert_batch_equivalent(_full(G["m16_np8_ids"], G["m16_np8_scores"]), _s(ix16, q, K, nprobe=8), ctx="golden m=16")
