"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's 8-bit scalar-quantised FLAT path, the checker
for csrc/sq8.cu.  Nothing under pyrope_b200/ imports this.

Follows Vector/ScalarQuantizer.cs:22-62 (Quantize: per-vector min / max, float32 arithmetic, Math.Round = round half
to even, clamp), :64-88 (Dequantize), Vector/VectorMath.cs:441-680 (L2Squared8Bit / DotProduct8Bit: exact integer
sums) and the quantised branch of BruteForceVectorIndex.Search (BruteForceVectorIndex.cs:297-336).  Pinned on the
reference's own cases: ScalarQuantizerTests.cs and VectorMathTests.cs:132-155."""
from __future__ import annotations

import numpy as np

F = np.float32


def quantize(v):
    """-> (bytes, min, max)"""
    v = np.asarray(v, F)
    if v.size == 0:
        return np.zeros(0, np.uint8), F(0), F(0)
    mn, mx = F(v.min()), F(v.max())
    rng = F(mx - mn)
    if rng == 0:
        return np.zeros(v.size, np.uint8), mn, mx
    scale = F(F(255.0) / rng)
    nz = ((v - mn).astype(F) * scale).astype(F)          # (vector[i] - min) * scale, in float32
    q = np.rint(nz.astype(np.float64))                    # Math.Round(double): ties to even
    return np.clip(q, 0, 255).astype(np.uint8), mn, mx


def dequantize(q, mn, mx):
    q = np.asarray(q, np.uint8)
    rng = F(F(mx) - F(mn))
    if rng == 0:
        return np.full(q.size, mn, F)
    scale = F(rng / F(255.0))
    return (F(mn) + (q.astype(F) * scale).astype(F)).astype(F)


def l2sq_8bit(a, b) -> int:
    d = np.asarray(a, np.int64) - np.asarray(b, np.int64)
    return int((d * d).sum())


def dot_8bit(a, b) -> int:
    return int((np.asarray(a, np.int64) * np.asarray(b, np.int64)).sum())


class Sq8FlatIndex:
    """BruteForceVectorIndex with EnableQuantization (ids are ints; insertion order = scan order)."""

    def __init__(self, dim, metric="L2"):
        self.dim, self.metric = dim, metric
        self.enable = True
        self.ids, self.q, self.dead, self.pos = [], [], [], {}

    def _q(self, v):
        return quantize(v)[0] if self.enable else None   # :167-181: empty when the flag is off

    def add(self, i, v):
        assert i not in self.pos
        self.pos[i] = len(self.ids)
        self.ids.append(i); self.q.append(self._q(v)); self.dead.append(False)

    def upsert(self, i, v):
        if i in self.pos:
            self.q[self.pos[i]] = self._q(v)             # :203-214
            self.dead[self.pos[i]] = False
        else:
            self.add(i, v)

    def delete(self, i):
        if i not in self.pos:
            return False
        self.dead[self.pos.pop(i)] = True
        return True

    def scores(self, query, max_scans=None):
        """-> list of (id, score) for every row the quantised loop scores, in scan order (:306-333)."""
        qq = quantize(query)[0]
        count = len(self.ids)
        limit = count if max_scans is None else min(max_scans, count)
        out, scanned = [], 0
        if limit <= 0:
            return out
        for s in range(count):
            if self.dead[s]:
                continue
            if scanned >= limit:
                break
            scanned += 1
            t = self.q[s]
            if t is None or len(t) != self.dim:
                continue
            if self.metric == "L2":
                sc = F(-l2sq_8bit(qq, t))
            else:                                          # inner product AND cosine: the raw byte dot product
                sc = F(dot_8bit(qq, t))
            out.append((self.ids[s], sc))
        return out
