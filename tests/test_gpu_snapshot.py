"""IVectorIndex.Snapshot / Load through the C ABI: a freshly created index loaded from a snapshot answers
exactly like the index that wrote it, for all three kinds, and stays writable afterwards."""
import numpy as np
import pytest

from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import pyrope_b200 as pg
    pg._lib.check(pg.load().pyrope_gpu_init(0))
    return pg


def _same(a, b):
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("kind", ["flat", "ivf_flat", "ivf_pq"])
def test_snapshot_load_round_trip(gpu, tmp_path, kind):
    base = orc.random_vectors(6000, 128, 42)
    q = orc.random_vectors(80, 128, 1337)
    mk = {"flat": lambda: gpu.GpuIndex(gpu.FLAT, 128, gpu.L2),
          "ivf_flat": lambda: gpu.GpuIndex(gpu.IVF_FLAT, 128, gpu.COSINE, nlist=16),
          "ivf_pq": lambda: gpu.GpuIndex(gpu.IVF_PQ, 128, gpu.L2, nlist=16, m=16, k=256)}[kind]
    a = mk()
    a.add(base[:5000])
    if kind != "flat":
        a.build()
    a.add(base[5000:])                      # FLAT: more rows; IVF: rows in the post-build buffer
    for r in (3, 77, 5500):
        if kind == "ivf_pq" and r < 5000:   # IVF_PQ deletes only buffered rows (IvfPqVectorIndex.cs:48-53)
            continue
        assert a.delete_row(r)
    kw = {} if kind == "flat" else {"nprobe": 4}
    want = a.search(q, 10, **kw)
    path = tmp_path / f"{kind}.bin"
    a.snapshot(path)
    b = mk()
    b.load(path)
    assert b.stats() == a.stats() and b.is_built() == a.is_built()
    _same(want, b.search(q, 10, **kw))
    if kind != "flat":
        np.testing.assert_array_equal(a.centroids(), b.centroids())
        _same(a.lists(), b.lists())
    # still writable: the same writes on both give the same answers
    extra = orc.random_vectors(20, 128, 5)
    for ix in (a, b):
        assert ix.add(extra) == 6000
        assert ix.delete_row(6001)
    _same(a.search(q, 10, **kw), b.search(q, 10, **kw))
    # wrong target / missing file
    other = gpu.GpuIndex(gpu.FLAT, 64, gpu.L2)
    with pytest.raises(gpu.PyropeGpuError):
        other.load(path)
    with pytest.raises(gpu.PyropeGpuError):
        b.load(tmp_path / "missing.bin")


# ---- a corrupt or truncated file must be refused as a whole: the live index keeps answering as before ---------------
def _mk(gpu, kind):
    return {"flat": lambda: gpu.GpuIndex(gpu.FLAT, 64, gpu.L2),
            "ivf_flat": lambda: gpu.GpuIndex(gpu.IVF_FLAT, 64, gpu.L2, nlist=8),
            "ivf_pq": lambda: gpu.GpuIndex(gpu.IVF_PQ, 64, gpu.L2, nlist=8, m=16, k=256)}[kind]()


@pytest.mark.parametrize("kind", ["flat", "ivf_flat", "ivf_pq"])
def test_truncated_and_bit_flipped_snapshots_leave_the_index_untouched(gpu, tmp_path, kind):
    import struct
    base = orc.random_vectors(1500, 64, 3)
    q = orc.random_vectors(16, 64, 4)
    a = _mk(gpu, kind)
    a.add(base[:1200])
    if kind != "flat":
        a.build()
    a.add(base[1200:])
    if kind != "ivf_pq":
        assert a.delete_row(7)
    assert a.delete_row(1300)
    path = tmp_path / "good.bin"
    a.snapshot(path)
    blob = path.read_bytes()

    live = _mk(gpu, kind)                 # the index a bad load must not disturb
    live.add(base[:300])
    kw = {} if kind == "flat" else {"nprobe": 4}
    want, stats = live.search(q, 5, **kw), live.stats()

    def refused(data, what):
        p = tmp_path / "bad.bin"
        p.write_bytes(data)
        with pytest.raises(gpu.PyropeGpuError):
            live.load(p)
        assert live.stats() == stats, what
        _same(want, live.search(q, 5, **kw))

    for cut in (4, 20, 60, len(blob) // 3, len(blob) // 2, len(blob) - 9, len(blob) - 1):
        refused(blob[:cut], f"truncated at {cut}")
    # header counters: next_row (offset 32), nslots (40), live (48), ndead (56)
    for off, val in ((32, -5), (40, 1 << 50), (40, 3), (48, 10 ** 9), (56, 12345)):
        bad = bytearray(blob)
        bad[off:off + 8] = struct.pack("<q", val)
        refused(bytes(bad), f"int64 at {off} = {val}")
    # the length field of the row vectors (offset 64): zero rows while nslots > 0, and a length past the end of the file
    for val in (0, 1 << 41):
        bad = bytearray(blob)
        bad[64:72] = struct.pack("<Q", val)
        refused(bytes(bad), f"row-vector byte count {val}")
    if kind != "flat":
        # list offsets: find the offsets vector (nc + 1 = 9 int64 starting with 0 and ending at list_total) and break it
        off_h, rows_h, _ = a.lists()
        needle = struct.pack("<Q", len(off_h)) + off_h.astype("<i8").tobytes()
        at = blob.find(needle)
        assert at > 0
        bad = bytearray(blob)
        bad[at + 8 + 8 * 3: at + 8 + 8 * 4] = struct.pack("<q", int(off_h[-1]) + 100)   # not monotone / past the end
        refused(bytes(bad), "list offsets not monotone")
        bad = bytearray(blob)
        bad[at + 8 + 8 * 8: at + 8 + 8 * 9] = struct.pack("<q", int(off_h[-1]) - 1)      # does not end at list_total
        refused(bytes(bad), "list offsets do not end at list_total")
        # a list row ordinal outside [0, next_row)
        rn = rows_h.astype("<i8").tobytes()
        at = blob.find(rn)
        assert at > 0
        bad = bytearray(blob)
        bad[at:at + 8] = struct.pack("<q", 10 ** 12)
        refused(bytes(bad), "list row ordinal out of range")
    # and the good file still loads
    live.load(path)
    assert live.stats() == a.stats()
    _same(a.search(q, 5, **kw), live.search(q, 5, **kw))


def test_vindex_load_checks_the_id_table_before_touching_the_index(gpu, tmp_path):
    """pyrope_vindex_load: a missing / corrupt .ids file is detected BEFORE the index file is loaded, so rows never end
    up under another process's labels; dead rows are labelled -1."""
    from pyrope_b200 import vector_index as vi
    base = orc.random_vectors(50, 16, 9)
    a = vi.BruteForceVectorIndex(16)
    for i in range(40):
        a.Add(f"k{i}", base[i])
    a.Delete("k3")
    p = str(tmp_path / "v.bin")
    a.Snapshot(p)
    b = vi.BruteForceVectorIndex(16)
    b.Add("mine", base[45])
    before = b.Search(base[45], 3)
    ids = (tmp_path / "v.bin.ids").read_bytes()
    (tmp_path / "v.bin.ids").write_bytes(ids[: len(ids) // 2])
    with pytest.raises(vi.ArgumentException):
        b.Load(p)
    assert b.Search(base[45], 3) == before and b.GetStats().Count == 1
    (tmp_path / "v.bin.ids").unlink()
    with pytest.raises(FileNotFoundError):
        b.Load(p)
    assert b.Search(base[45], 3) == before
    (tmp_path / "v.bin.ids").write_bytes(ids)
    b.Load(p)
    assert b.GetStats().Count == 39
    assert [r.Id for r in b.Search(base[5], 1)] == ["k5"]
    assert all(r.Id != "k3" for r in b.Search(base[3], 39))


def test_id_table_recycles_ordinals_under_churn(gpu):
    """The process-wide id table is reference counted: ids that no index holds any more give their ordinal back."""
    import ctypes as C
    from pyrope_b200 import vector_index as vi
    L = gpu.load()

    def size():
        live, slots = C.c_int64(0), C.c_int64(0)
        assert L.pyrope_vindex_id_table_size(C.byref(live), C.byref(slots)) == 0
        return live.value, slots.value

    ix = vi.BruteForceVectorIndex(8)
    v = np.ones(8, np.float32)
    live0, _ = size()
    for i in range(200):
        ix.Add(f"churn-{i}", v * i)
    assert size()[0] == live0 + 200
    slots_after_first = size()[1]
    for round_ in range(5):
        for i in range(200):
            assert ix.Delete(f"churn-{i}" if round_ == 0 else f"churn-{round_}-{i}")
        assert size()[0] == live0
        for i in range(200):
            ix.Add(f"churn-{round_ + 1}-{i}", v * i)
        assert size() == (live0 + 200, slots_after_first)      # ordinals re-used, table did not grow
    res = ix.Search(v * 7, 1)
    assert res[0].Id == "churn-5-7"
    ix.close()
    assert size()[0] == live0                                  # destroying the index releases its ids


def test_sq8_load_follows_the_loading_index_flag(gpu, tmp_path):
    """BruteForceVectorIndex.Load re-adds every row through InternalAdd (BruteForceVectorIndex.cs:93-97 -> :162-184): a row
    gets a quantised form iff the LOADING index has EnableQuantization on; the flag is not part of the file."""
    from oracle import sq8_oracle as sq
    rng = np.random.default_rng(2)
    base = rng.standard_normal((400, 32)).astype(np.float32)
    Q = rng.standard_normal((9, 32)).astype(np.float32)
    a = gpu.GpuIndex(gpu.FLAT, 32, gpu.L2)
    a.set_quantization(True)
    a.add(base[:200])
    a.set_quantization(False)
    a.add(base[200:])             # no quantised form on the writing side
    path = tmp_path / "sq8.bin"
    a.snapshot(path)
    on = gpu.GpuIndex(gpu.FLAT, 32, gpu.L2)
    on.set_quantization(True)
    on.load(path)                 # every row re-added with the flag on: all 400 visible to the quantised scan
    ref = sq.Sq8FlatIndex(32, "L2")
    for i in range(400):
        ref.add(i, base[i])
    sc, rows, cnt = on.search(Q, 10)
    assert on.last_search_kernel()[0] == "sq8_scan_kernel"
    for qi in range(len(Q)):
        want = sorted((s for _, s in ref.scores(Q[qi], None)), reverse=True)[:10]
        np.testing.assert_array_equal(np.asarray(want, np.float32), sc[qi])
    off = gpu.GpuIndex(gpu.FLAT, 32, gpu.L2)
    off.load(path)                # flag off while loading: rows have no quantised form ...
    off.set_quantization(True)    # ... and stay invisible to the quantised scan once it is switched on (:312-322)
    sc, rows, cnt = off.search(Q, 10)
    assert (cnt == 0).all()
