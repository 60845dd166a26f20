"""GPU parity at the SHAPE of BASELINE config 5 (reduced only in base size): nlist >= 16,384 so the coarse stage
takes the two-pass one-TF32 tensor-core probe + P+8 exact re-rank, nprobe 64, m 16 so the scan is the list-major
fixed-point ADC kernel with histogram thresholds — the combination the headline number is measured on
(IvfPqVectorIndex.cs:118-212).  The oracle adopts the GPU-built index (same centroids, codebooks, codes), so the
comparison isolates the SEARCH path; the codes themselves are checked against ProductQuantizer.Encode on a sample.
bench.py repeats the same comparison at full size inside every bench line (`parity`)."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.parity import assert_batch_equivalent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(request):
    import torch

    import pyrope_b200 as pg
    from pyrope_b200 import _lib
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    dim, n, nlist, m = 128, 2_000_000, 16384, 16
    ix = pg.GpuIndex(pg.IVF_PQ, dim, pg.L2, nlist=nlist, m=m, k=256)
    ix.set_train_params(4 * nlist, 2)  # bounded training (opt-in deviation): codebooks are inputs to the search
    ix.reserve(n)
    stage = torch.empty(n * dim, dtype=torch.float32, device="cuda")
    _lib.fill_uniform_device(stage.data_ptr(), n * dim, 42, 0, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ix.add_device(stage.data_ptr(), n)
    base_head = stage[:4096 * dim].view(4096, dim).cpu().numpy()
    del stage
    ix.build()
    nq = 3000
    q = torch.empty(nq * dim, dtype=torch.float32, device="cuda")
    _lib.fill_uniform_device(q.data_ptr(), q.numel(), 1337, 0, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    Q = q.view(nq, dim).cpu().numpy()
    ref = orc.IvfPqIndex(dim, orc.L2, m=m, k=256, nlist=nlist)
    off, rows, codes = ix.lists()
    cb, _ = ix.codebooks()
    cent = ix.centroids()
    ref.adopt(cent, cb, off, rows, codes)
    return dict(pg=pg, ix=ix, ref=ref, Q=Q, base_head=base_head, off=off, rows=rows, codes=codes, cb=cb, cent=cent)


def test_c5_shape_search_matches_oracle(big):
    ix, ref, Q = big["ix"], big["ref"], big["Q"]
    sc, rows, cnt = ix.search(Q, 10, nprobe=64)
    assert ix.last_search_kernel()[0] == "ivfpq_lm_scan_kernel"
    assert_batch_equivalent(ref.search_batch(Q, 10, nprobe=64), (rows, sc, cnt), ctx="c5 shape k=10 nprobe=64")


def test_c5_shape_topk100(big):
    ix, ref, Q = big["ix"], big["ref"], big["Q"][:500]
    sc, rows, cnt = ix.search(Q, 100, nprobe=64)
    assert_batch_equivalent(ref.search_batch(Q, 100, nprobe=64), (rows, sc, cnt), ctx="c5 shape k=100 nprobe=64")


def test_c5_shape_codes_and_assignment_bit_exact_on_a_sample(big):
    """Rows 0..4095: the list a row sits in must be KMeansUtils.FindNearestCentroid's choice and its code bytes
    ProductQuantizer.Encode's, both by the oracle, given the GPU-trained codebooks."""
    off, rows, codes, cent, cb = big["off"], big["rows"], big["codes"], big["cent"], big["cb"]
    head = big["base_head"]
    pos_of = np.full(int(rows.max()) + 1, -1, np.int64)
    pos_of[rows] = np.arange(len(rows))
    list_of_pos = np.searchsorted(off, np.arange(len(rows)), side="right") - 1
    pq = orc.ProductQuantizer(128, 16, 256)
    pq.set_codebook(cb)
    for r in range(0, 4096, 16):
        p = pos_of[r]
        assert p >= 0
        c = orc.find_nearest_centroid(head[r], cent, orc.L2)
        assert int(list_of_pos[p]) == c, f"row {r}: list {int(list_of_pos[p])} != oracle {c}"
        np.testing.assert_array_equal(codes[p], pq.encode(head[r] - cent[c]))
