"""The identity the list-major scan relies on for m < 16 (csrc/pq_lm.cu, lm_expand_*): an (m, d/m) product quantiser is a
(16, d/16) one whose codewords are the pieces of the original codewords and whose codes repeat every byte 16/m times.
Checked on the CPU against the oracle's ProductQuantizer.ComputeDistanceTable (ProductQuantizer.cs:98-120): the
16-table ADC distance equals the m-table one up to fp32 summation order (the GPU re-scores survivors in the original
quantiser, so only this approximate equality is needed)."""
import numpy as np
import pytest

from oracle import pyoracle as orc


def expand_codebook(cb: np.ndarray) -> np.ndarray:
    """[m][k][d/m] -> [16][k][d/16]: virtual table v = m * (16/m) + piece"""
    m, k, subr = cb.shape
    per = 16 // m
    sub16 = subr // per
    return cb.reshape(m, k, per, sub16).transpose(0, 2, 1, 3).reshape(16, k, sub16).copy()


def expand_codes(codes: np.ndarray, m: int) -> np.ndarray:
    return np.repeat(codes, 16 // m, axis=1)


@pytest.mark.parametrize("dim,m", [(128, 8), (128, 4), (64, 4), (64, 2), (128, 1)])
def test_sixteen_table_view_gives_the_same_adc_distance(dim, m):
    rng = np.random.default_rng(100 + m)
    k = 256
    cb = (rng.random((m, k, dim // m), dtype=np.float32) - 0.5).astype(np.float32)
    pq = orc.ProductQuantizer(dim, m, k)
    pq.set_codebook(cb)
    pq16 = orc.ProductQuantizer(dim, 16, k)
    pq16.set_codebook(expand_codebook(cb))
    codes = rng.integers(0, k, (500, m), dtype=np.uint8)
    codes16 = expand_codes(codes, m)
    assert codes16.shape == (500, 16)
    for _ in range(5):
        r = (rng.random(dim, dtype=np.float32) - 0.5).astype(np.float32)   # a residual query
        t = pq.distance_table(r)          # [m][k]
        t16 = pq16.distance_table(r)      # [16][k]
        d = t[np.arange(m)[None, :], codes].astype(np.float64).sum(axis=1)
        d16 = t16[np.arange(16)[None, :], codes16].astype(np.float64).sum(axis=1)
        np.testing.assert_allclose(d16, d, rtol=2e-6, atol=1e-7)
        # and the pieces really are pieces: every table of the original is the sum of its 16/m virtual tables
        per = 16 // m
        np.testing.assert_allclose(t16.reshape(m, per, k).astype(np.float64).sum(axis=1), t.astype(np.float64), rtol=2e-6, atol=1e-7)


def test_ivfflat_adopt_reproduces_the_built_index():
    """orc_ivfflat_adopt (what bench.py uses to hand a GPU-built IVF_FLAT index to the oracle at sizes where CPU k-means
    would take hours): an index rebuilt from another one's centroids and lists answers identically."""
    for metric in (orc.L2, orc.IP, orc.COSINE):
        base = orc.random_vectors(3_000, 24, 9)
        a = orc.IvfFlatIndex(24, metric, nlist=12)
        a.add_batch(base)
        a.build()
        lists = a.lists()
        off = np.zeros(len(lists) + 1, np.int64)
        off[1:] = np.cumsum([len(x) for x in lists])
        ids = np.concatenate(lists)
        b = orc.IvfFlatIndex(24, metric, nlist=12)
        b.adopt(a.centroids(), off, ids, base[ids])
        q = orc.random_vectors(40, 24, 10)
        ra, rb = a.search_batch(q, 10, nprobe=4), b.search_batch(q, 10, nprobe=4)
        for x, y in zip(ra, rb):
            np.testing.assert_array_equal(x, y)
