"""The N > 1 path on CPU: two `gloo` ranks (127.0.0.1) run the sharding rules of pyrope_b200/shard.py — list
ownership l % world, per-rank top-k, all-gather, merge — with the CPU oracle standing in for each rank's
GPU engine, and the merged result must equal the unsharded oracle.  Also the slicing arithmetic."""
import os
import socket

import numpy as np
import pytest

from oracle import pyoracle as orc
from pyrope_b200 import shard
from tests.parity import assert_batch_equivalent

N, NQ, DIM, K, NLIST, NPROBE = 4000, 37, 64, 10, 16, 5


def test_slicing_arithmetic():
    for nq, world in ((10_000, 8), (37, 2), (5, 8), (1, 4)):
        got = []
        per0 = None
        for r in range(world):
            lo, hi, per = shard.query_slice(nq, r, world)
            per0 = per if per0 is None else per0
            assert per == per0 and 0 <= hi - lo <= per
            got += list(range(lo, hi))
        assert got == list(range(nq)) and per0 * world >= nq
    for n, world in ((10_000_000, 8), (10, 3)):
        blocks = [shard.row_block(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    assert [shard.list_owner(l, 4) for l in range(6)] == [0, 1, 2, 3, 0, 1]


def test_merge_topk_reference():
    s = np.array([[[5.0, 3.0, 0.0]], [[4.0, 3.0, 1.0]]], np.float32)
    r = np.array([[[10, 11, -1]], [[20, 21, 22]]], np.int64)
    ms, mr, mc = shard.merge_topk(s, r, 4)
    assert mr[0].tolist() == [10, 20, 11, 21] and ms[0].tolist() == [5.0, 4.0, 3.0, 3.0] and mc[0] == 4  # tie -> lower rank
    ms, mr, mc = shard.merge_topk(s[:, :, 2:], r[:, :, 2:], 3)
    assert mr[0].tolist() == [22, -1, -1] and mc[0] == 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = orc.random_vectors(N, DIM, 42)
    q = orc.random_vectors(NQ, DIM, 1337)
    # ---- IVF_PQ: every rank sees every row, keeps the lists it owns (pyrope_index_set_shard semantics)
    full = orc.IvfPqIndex(DIM, orc.L2, m=16, k=256, nlist=NLIST)
    full.add_batch(base)
    full.build()
    lists = full.lists()
    keep = [(ids, codes) if shard.list_owner(l, world) == rank else (ids[:0], codes[:0]) for l, (ids, codes) in enumerate(lists)]
    off = np.concatenate([[0], np.cumsum([len(i) for i, _ in keep])]).astype(np.int64)
    mine = orc.IvfPqIndex(DIM, orc.L2, m=16, k=256, nlist=NLIST)
    mine.adopt(full.centroids(), full.pq().codebook(), off, np.concatenate([i for i, _ in keep]),
               np.concatenate([c for _, c in keep]))
    ids, sc, cnt = mine.search_batch(q, K, nprobe=NPROBE)
    ids = np.where(np.arange(K)[None, :] < cnt[:, None], ids, -1)
    g_s = [torch.empty((NQ, K), dtype=torch.float32) for _ in range(world)]
    g_r = [torch.empty((NQ, K), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(g_s, torch.from_numpy(sc.astype(np.float32)))
    dist.all_gather(g_r, torch.from_numpy(ids.astype(np.int64)))
    ms, mr, mc = shard.merge_topk(torch.stack(g_s).numpy(), torch.stack(g_r).numpy(), K)
    # ---- FLAT: contiguous row blocks, global row numbers as labels
    lo, hi = shard.row_block(N, rank, world)
    fl = orc.FlatIndex(DIM, orc.IP)
    fl.add_batch(base[lo:hi], ids=np.arange(lo, hi))
    fids, fsc, fcnt = fl.search_batch(q, K)
    dist.all_gather(g_s, torch.from_numpy(fsc.astype(np.float32)))
    dist.all_gather(g_r, torch.from_numpy(fids.astype(np.int64)))
    fs, fr, fc = shard.merge_topk(torch.stack(g_s).numpy(), torch.stack(g_r).numpy(), K)
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), ms=ms, mr=mr, mc=mc, fs=fs, fr=fr, fc=fc)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_match_unsharded_oracle(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "merged.npz")
    base = orc.random_vectors(N, DIM, 42)
    q = orc.random_vectors(NQ, DIM, 1337)
    full = orc.IvfPqIndex(DIM, orc.L2, m=16, k=256, nlist=NLIST)
    full.add_batch(base)
    full.build()
    assert_batch_equivalent(full.search_batch(q, K, nprobe=NPROBE), (got["mr"], got["ms"], got["mc"]), ctx="gloo IVF_PQ")
    fl = orc.FlatIndex(DIM, orc.IP)
    fl.add_batch(base)
    assert_batch_equivalent(fl.search_batch(q, K), (got["fr"], got["fs"], got["fc"]), ctx="gloo FLAT")


@pytest.mark.gpu
def test_split_coarse_equals_fused_search_and_device_merge():
    """coarse_probe on query slices + probed search == plain search; shard.merge_topk == pyrope_topk_merge_device."""
    torch = pytest.importorskip("torch")
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    base = orc.random_vectors(12_000, 128, 42)
    qh = orc.random_vectors(101, 128, 1337)
    q = torch.from_numpy(qh).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for kind, kw, P in ((pg.IVF_PQ, dict(nlist=32, m=16, k=256), 6), (pg.IVF_FLAT, dict(nlist=32), 3)):
        ix = pg.GpuIndex(kind, 128, pg.L2, **kw)
        ix.add(base)
        ix.build()
        ref = ix.search(qh, 10, nprobe=P)
        world = 3
        per = shard.query_slice(101, 0, world)[2]
        probes = torch.full((world * per, P), -1, dtype=torch.int64, device="cuda")
        for r in range(world):
            lo, hi, _ = shard.query_slice(101, r, world)
            ix.coarse_probe_device(q[lo:hi].data_ptr(), hi - lo, P, probes[r * per:].data_ptr(), stream=st)
        s = torch.empty((101, 10), dtype=torch.float32, device="cuda")
        rw = torch.empty((101, 10), dtype=torch.int64, device="cuda")
        c = torch.empty((101,), dtype=torch.int32, device="cuda")
        ix.search_probed_device(q.data_ptr(), 101, 10, P, probes.data_ptr(), s.data_ptr(), rw.data_ptr(), c.data_ptr(), stream=st)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(rw.cpu().numpy(), ref[1])
        np.testing.assert_array_equal(s.cpu().numpy(), ref[0])
    # device merge vs the numpy reference
    rng = np.random.default_rng(0)
    S = np.sort(rng.random((4, 50, 8)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    R = rng.integers(0, 1 << 40, (4, 50, 8)).astype(np.int64)
    R[1, :, 6:] = -1
    dS, dR = torch.from_numpy(S).cuda(), torch.from_numpy(R).cuda()
    ms = torch.empty((50, 8), dtype=torch.float32, device="cuda")
    mr = torch.empty((50, 8), dtype=torch.int64, device="cuda")
    mc = torch.empty((50,), dtype=torch.int32, device="cuda")
    pg._lib.topk_merge_device(50, 4, 8, 8, dS.data_ptr(), dR.data_ptr(), ms.data_ptr(), mr.data_ptr(), mc.data_ptr(), stream=st)
    torch.cuda.synchronize()
    es, er, ec = shard.merge_topk(S, R, 8)
    np.testing.assert_array_equal(ms.cpu().numpy(), es)
    np.testing.assert_array_equal(mr.cpu().numpy(), er)
    np.testing.assert_array_equal(mc.cpu().numpy(), ec)
