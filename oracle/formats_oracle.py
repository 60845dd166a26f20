"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's payload / dataset formats, used as the checker for
csrc/formats.cu.  Nothing under pyrope_b200/ imports this.

Follows Utils/VectorParsing.cs:10-101 (ParseVector: JSON array -> CSV -> raw float32), 
Benchmarks/Encoding/VectorEncoding.cs:8-16 and Benchmarks/Datasets/FvecsReader.cs:14-60.  Pinned on the reference's
own cases: VectorEncodingTests.cs, FvecsReaderTests.cs and the payloads of VectorCommandParserTests.cs.
Written independently of the C++ (regular expressions + Python's own number parsing) so that the two can disagree."""
from __future__ import annotations

import json
import re
import struct

import numpy as np

_JSON_NUM = re.compile(r"-?(?:0|[1-9][0-9]*)(?:\.[0-9]+)?(?:[eE][+-]?[0-9]+)?\Z")
_NET_FLOAT = re.compile(r"[+-]?(?:[0-9]+\.?[0-9]*|\.[0-9]+)(?:[eE][+-]?[0-9]+)?\Z")
_WS = " \t\n\v\f\r"


class FormatError(ValueError):
    pass


def _f32(x: float) -> np.float32:
    with np.errstate(over="ignore"):
        return np.float32(x)


def _try_json(text: str):
    if not text.strip(_WS) or text[0] != "[":
        return None
    try:
        val = json.loads(text, parse_constant=lambda c: (_ for _ in ()).throw(ValueError(c)))
    except ValueError:
        return None
    if not isinstance(val, list) or not val:
        return None
    # System.Text.Json: numbers only (no bool / null / nested / strings), strict number grammar
    toks = re.findall(r"[^\[\],\s]+", text)
    if len(toks) != len(val) or not all(_JSON_NUM.match(t) for t in toks):
        return None
    out = [_f32(float(t)) for t in toks]
    if not all(np.isfinite(out)):
        raise FormatError("out of bounds for a Single")
    return np.array(out, np.float32)


def _try_csv(text: str):
    if not text.strip(_WS):
        return None
    parts = [p.strip(_WS) for p in re.split(r"[, ]", text)]
    parts = [p for p in parts if p]
    if not parts:
        return None
    out = []
    for p in parts:
        low = p.lower().lstrip("+-")
        if low == "nan":
            out.append(np.float32("nan"))
        elif low == "infinity":
            out.append(np.float32("-inf") if p[0] == "-" else np.float32("inf"))
        elif _NET_FLOAT.match(p):
            out.append(_f32(float(p)))
        else:
            return None
    return np.array(out, np.float32)


def parse_vector(data: bytes) -> np.ndarray:
    """VectorParsing.ParseVector:10-35."""
    if not data:
        raise ValueError("Vector payload is empty.")
    text = data.decode("utf-8", "replace")
    v = _try_json(text)
    if v is not None:
        return v
    v = _try_csv(text)
    if v is not None:
        return v
    if len(data) % 4 == 0:
        return np.frombuffer(data, "<f4").copy()
    raise FormatError("Unsupported vector format.")


def to_little_endian_bytes(vec) -> bytes:
    """VectorEncoding.ToLittleEndianBytes:8-16."""
    return b"".join(struct.pack("<f", float(x)) for x in vec)


def read_fvecs(path: str, limit: int | None = None):
    """FvecsReader.Read:14-60 -> list of float32 arrays."""
    out = []
    if limit is not None and limit <= 0:
        return out
    with open(path, "rb") as f:
        while True:
            if limit is not None and len(out) >= limit:
                break
            hdr = f.read(4)
            if len(hdr) < 4:
                break
            (dim,) = struct.unpack("<i", hdr)
            if dim <= 0:
                raise ValueError(f"Invalid vector dimension {dim} in fvecs file.")
            body = f.read(dim * 4)
            if len(body) != dim * 4:
                raise EOFError("Truncated fvecs record.")
            out.append(np.frombuffer(body, "<f4").copy())
    return out
