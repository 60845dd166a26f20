"""The peer-memory all-gather of the sharded search (csrc/peer.cu).  One GPU is enough to exercise the protocol: two
ranks of one group live on the same device, on two streams, and hand each other their buffers directly
(pyrope_peer_group_attach) — the kernels, epochs, parity halves and flag words are the ones a multi-GPU run uses; the
real NVLink path runs in bench.py under torchrun (`details.collective`).

Ranks that SHARE a device must not allocate or free device memory while an exchange is in flight: cudaMalloc / cudaFree
may wait for the whole device, i.e. for this rank's own wait kernel, which waits for a peer whose host thread sits in
the same kind of call — a deadlock that separate devices cannot have.  Hence every buffer below is allocated, and every
library workspace grown, before the rank threads start."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Raw:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


@pytest.fixture(scope="module")
def gpu():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    return pg


def _in_threads(fn, world):
    """One host thread per rank, as in a real deployment: a rank's enqueue never waits behind a peer's."""
    import threading
    errs = []

    def body(r):
        try:
            fn(r)
        except Exception as ex:  # noqa: BLE001
            errs.append((r, ex))

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


def _groups(pg, world, slot_bytes, n_slots):
    gs = [pg.PeerGroup(world, r, slot_bytes, n_slots) for r in range(world)]
    bufs = [g.buffer() for g in gs]
    for g in gs:
        g.attach(bufs)
    return gs


@pytest.mark.parametrize("world", [2, 4])
def test_allgather_rounds_on_one_device(gpu, world):
    import torch
    nbytes = 256 * 1024 + 16
    gs = _groups(gpu, world, nbytes, 2)
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(0)
    for it in range(6):  # both parity halves, several times over; slot 1 every other round (its own epoch counter)
        slot = it & 1
        host = [rng.integers(0, 256, nbytes, dtype=np.uint8) for _ in range(world)]
        src = [torch.from_numpy(h).cuda() for h in host]
        torch.cuda.synchronize()
        outs = [torch.empty(world * nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()

        def run(r):
            with torch.cuda.stream(streams[r]):
                ptr = gs[r].allgather(slot, src[r].data_ptr(), nbytes, stream=streams[r].cuda_stream)
                outs[r].copy_(torch.as_tensor(_Raw(ptr, world * nbytes), device="cuda"))  # consumer on the same stream

        _in_threads(run, world)
        torch.cuda.synchronize()
        want = np.concatenate(host)
        for r in range(world):
            np.testing.assert_array_equal(outs[r].cpu().numpy(), want, err_msg=f"round {it} rank {r}")
    for g in gs:
        g.close()


def test_allgather_argument_checks(gpu):
    import torch
    gs = _groups(gpu, 2, 4096, 1)
    src = torch.zeros(8192, dtype=torch.uint8, device="cuda")
    with pytest.raises(gpu.PyropeGpuError):
        gs[0].allgather(0, src.data_ptr(), 24)         # not a multiple of 16
    with pytest.raises(gpu.PyropeGpuError):
        gs[0].allgather(0, src.data_ptr(), 8192)       # more than the slot holds
    with pytest.raises(gpu.PyropeGpuError):
        gs[0].allgather(3, src.data_ptr(), 64)         # no such slot
    with pytest.raises(gpu.PyropeGpuError):
        gs[0].allgather(0, src.data_ptr() + 4, 64)     # misaligned source
    lone = gpu.PeerGroup(2, 0, 4096, 1)
    with pytest.raises(gpu.PyropeGpuError):
        lone.allgather(0, src.data_ptr(), 64)          # peers never attached
    with pytest.raises(gpu.PyropeGpuError):
        gpu.PeerGroup(1, 0, 4096, 1)                   # a world of one has nothing to exchange
    lone.close()
    for g in gs:
        g.close()


def test_sharded_search_through_the_peer_exchange(gpu):
    """Two list shards of one IVF_PQ index on one device: probe lists of a coarse stage split by query and the local
    top-k lists travel through the peer group; the merged result equals the unsharded index's."""
    import torch
    from oracle import pyoracle as orc
    from pyrope_b200 import _lib
    from pyrope_b200.shard import query_slice
    dim, nlist, nq, k, P, world = 128, 32, 200, 10, 8, 2
    base = orc.random_vectors(20_000, dim, 5)
    q = orc.random_vectors(nq, dim, 6)
    one = gpu.GpuIndex(gpu.IVF_PQ, dim, gpu.L2, nlist=nlist, m=16, k=256)
    one.add(base)
    one.build()
    cent = one.centroids()
    cb, _ = one.codebooks()
    want = one.search(q, k, nprobe=P)
    shards = []
    for r in range(world):
        ix = gpu.GpuIndex(gpu.IVF_PQ, dim, gpu.L2, nlist=nlist, m=16, k=256)
        ix.set_shard(r, world)
        ix.set_codebooks(cent, cb)
        ix.add(base)
        ix.build()
        shards.append(ix)
    per = query_slice(nq, 0, world)[2]
    gs = _groups(gpu, world, max(per * P * 8, nq * k * 8), 3)
    streams = [torch.cuda.Stream() for _ in range(world)]
    Q = torch.from_numpy(q).cuda()
    res = []
    for r in range(world):  # every buffer of the exchange, and every library workspace, exists before the ranks start
        mine = torch.full((per, P), -1, dtype=torch.int64, device="cuda")
        sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        rw = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        cn = torch.empty((nq,), dtype=torch.int32, device="cuda")
        res.append((torch.empty_like(sc), torch.empty_like(rw), torch.empty_like(cn), mine, sc, rw, cn))
        full = torch.empty((nq, P), dtype=torch.int64, device="cuda")
        st0 = torch.cuda.current_stream().cuda_stream
        shards[r].coarse_probe_device(Q.data_ptr(), nq, P, full.data_ptr(), stream=st0)   # same shapes as below, or larger
        shards[r].search_probed_device(Q.data_ptr(), nq, k, P, full.data_ptr(), sc.data_ptr(), rw.data_ptr(), cn.data_ptr(), stream=st0)
        torch.cuda.synchronize()

    def run(r):
        with torch.cuda.stream(streams[r]):
            st = streams[r].cuda_stream
            lo, hi, _ = query_slice(nq, r, world)
            m_sc, m_rw, m_cn, mine, sc, rw, cn = res[r]
            shards[r].coarse_probe_device(Q[lo:hi].data_ptr(), hi - lo, P, mine.data_ptr(), stream=st)
            probes = gs[r].allgather(0, mine.data_ptr(), per * P * 8, stream=st)
            shards[r].search_probed_device(Q.data_ptr(), nq, k, P, probes, sc.data_ptr(), rw.data_ptr(), cn.data_ptr(), stream=st)
            g_s = gs[r].allgather(1, sc.data_ptr(), nq * k * 4, stream=st)
            g_r = gs[r].allgather(2, rw.data_ptr(), nq * k * 8, stream=st)
            _lib.topk_merge_device(nq, world, k, k, g_s, g_r, m_sc.data_ptr(), m_rw.data_ptr(), m_cn.data_ptr(), stream=st)

    _in_threads(run, world)
    torch.cuda.synchronize()
    for r in range(world):
        np.testing.assert_array_equal(res[r][1].cpu().numpy(), want[1], err_msg=f"rank {r} rows")
        np.testing.assert_array_equal(res[r][0].cpu().numpy(), want[0], err_msg=f"rank {r} scores")
        np.testing.assert_array_equal(res[r][2].cpu().numpy(), want[2], err_msg=f"rank {r} counts")
    for g in gs:
        g.close()
