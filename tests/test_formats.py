"""CPU tests of the payload / dataset formats (csrc/formats.cu through the C ABI) against the reference's own cases
(VectorEncodingTests.cs, FvecsReaderTests.cs, the payloads of VectorCommandParserTests.cs) and against the
independent restatement in oracle/formats_oracle.py on generated payloads.  Floats must be bit-identical."""
import struct

import numpy as np
import pytest

from oracle import formats_oracle as fo
from pyrope_b200 import formats as fm


def _same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape and a.tobytes() == b.tobytes() or (np.isnan(a) == np.isnan(b)).all() and \
        np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


# ---- the reference's cases
def test_to_little_endian_bytes_encodes_float32():  # VectorEncodingTests.cs
    assert fm.ToLittleEndianBytes([1.0, 2.0]) == struct.pack("<f", 1.0) + struct.pack("<f", 2.0)
    assert fm.ToLittleEndianBytes([]) == b""
    with pytest.raises(ValueError):
        fm.ToLittleEndianBytes(None)


def _write_fvecs(path, rows):
    with open(path, "wb") as f:
        for r in rows:
            f.write(struct.pack("<i", len(r)))
            f.write(np.asarray(r, "<f4").tobytes())


def test_fvecs_read_all_and_limit(tmp_path):  # FvecsReaderTests.cs
    p = tmp_path / "a.fvecs"
    _write_fvecs(p, [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    got = fm.ReadFvecs(str(p))
    assert got.tolist() == [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]
    assert fm.ReadFvecs(str(p), limit=1).tolist() == [[1.0, 2.0, 3.0]]
    assert fm.ReadFvecs(str(p), limit=0).shape[0] == 0          # `limit is <= 0` -> yield break
    assert fm.ReadFvecs(str(p), skip=1).tolist() == [[4.0, 5.0, 6.0]]


def test_fvecs_errors_and_torn_tail(tmp_path):  # FvecsReader.cs:31-52
    p = tmp_path / "b.fvecs"
    _write_fvecs(p, [[1.0, 2.0]])
    with open(p, "ab") as f:
        f.write(b"\x02\x00")                                   # torn header: the read just ends
    assert fm.ReadFvecs(str(p)).tolist() == fo.read_fvecs(str(p))[0][None, :].tolist()
    with open(p, "ab") as f:
        f.write(b"\x00\x00" + struct.pack("<f", 1.0))           # header complete (d = 2), record short
    with pytest.raises(EOFError, match="Truncated"):
        fm.ReadFvecs(str(p))
    with pytest.raises(EOFError):
        fo.read_fvecs(str(p))
    q = tmp_path / "c.fvecs"
    q.write_bytes(struct.pack("<i", -3))
    with pytest.raises(ValueError, match="Invalid vector dimension -3"):
        fm.ReadFvecs(str(q))
    with pytest.raises(FileNotFoundError):
        fm.ReadFvecs(str(tmp_path / "missing.fvecs"))


def test_fvecs_random_matches_oracle(tmp_path):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((257, 24)).astype(np.float32)
    p = tmp_path / "r.fvecs"
    _write_fvecs(p, X)
    assert fm.ReadFvecs(str(p)).tobytes() == np.stack(fo.read_fvecs(str(p))).tobytes() == X.tobytes()
    assert fm.ReadFvecs(str(p), limit=100).tobytes() == X[:100].tobytes()


@pytest.mark.parametrize("payload,expect", [
    (b"[1.0,2.0,3.0]", [1.0, 2.0, 3.0]),        # VectorCommandParserTests.cs:19
    (b"[1.0]", [1.0]),                           # :48, :63
    (b"[ 1 , -2.5e0 ,\n3E+1 ]  ", [1.0, -2.5, 30.0]),
    (b"1,2,3", [1.0, 2.0, 3.0]),
    (b"1 2  3", [1.0, 2.0, 3.0]),
    (b" 0.5, -.25 ,+1e-3 ", [0.5, -0.25, 1e-3]),
    (b"NaN,Infinity,-infinity", [float("nan"), float("inf"), float("-inf")]),
    (b"1e40", [float("inf")]),                  # float.TryParse saturates since .NET Core 3.0
])
def test_parse_vector_text_forms(payload, expect):
    _same(fm.ParseVector(payload), expect)
    _same(fo.parse_vector(payload), expect)


def test_parse_vector_binary_and_errors():
    v = np.array([1.5, -2.25, 3e-5, 1e30], np.float32)
    raw = fm.ToLittleEndianBytes(v)
    assert fm.ParseVector(raw).tobytes() == v.tobytes() == fo.parse_vector(raw).tobytes()
    with pytest.raises(ValueError, match="empty"):
        fm.ParseVector(b"")
    for bad in (b"[1,2", b"abc", b"[1,,2]x"):
        if len(bad) % 4:
            with pytest.raises(fm.FormatException):
                fm.ParseVector(bad)
            with pytest.raises(fo.FormatError):
                fo.parse_vector(bad)
    # quirks reproduced on purpose: text that is not a vector but is a multiple of four bytes is taken as float32,
    # and four ASCII digits are a number, not a float32
    assert fm.ParseVector(b"abcd").tobytes() == b"abcd" == fo.parse_vector(b"abcd").tobytes()
    _same(fm.ParseVector(b"1234"), [1234.0])
    _same(fm.ParseVector(b"[1,2]xyz"), fo.parse_vector(b"[1,2]xyz"))  # trailing text: not JSON, not CSV, 8 raw bytes

def test_parse_vector_generated_payloads_match_oracle():
    rng = np.random.default_rng(7)
    pieces = ["1", "-2", "0.5", "1e3", "-1E-2", ".5", "5.", "+7", "01", "1e", "--1", "nan", "Infinity", "1.2.3", "", " ",
              "0x10", "1_0", "∞", "١"]
    n_checked = 0
    for _ in range(3000):
        k = int(rng.integers(1, 5))
        toks = [pieces[int(i)] for i in rng.integers(0, len(pieces), k)]
        sep = [",", " ", ", ", " ,", "\t"][int(rng.integers(0, 5))]
        body = sep.join(toks)
        text = body if rng.random() < 0.5 else "[" + body + "]"
        data = text.encode("utf-8")
        if not data:
            continue
        try:
            want = fo.parse_vector(data)
        except fo.FormatError:
            with pytest.raises(fm.FormatException):
                fm.ParseVector(data)
            continue
        _same(fm.ParseVector(data), want)
        n_checked += 1
    assert n_checked > 500
    for _ in range(300):  # raw float32 payloads of random bytes
        data = rng.integers(0, 256, int(rng.integers(1, 16)) * 4, dtype=np.uint8).tobytes()
        try:
            want = fo.parse_vector(data)
        except fo.FormatError:
            continue
        assert fm.ParseVector(data).tobytes() == want.tobytes()
