"""Host-side sharding rules of the multi-GPU path: one process per GPU, no data-path collective
except the all-gather of per-rank results (and of probe lists, when the coarse stage is split by query).

  * FLAT: contiguous row blocks, global row numbers travel as labels;
  * IVF_*: centroids / PQ codebooks replicated, inverted list l lives on rank l % world
    (pyrope_index_set_shard); the coarse ranking is computed by rank r for queries [lo_r, hi_r) only;
  * merge: per query, the world x k candidates -> best k, higher score first, ties to the lower rank
    (DeltaVectorIndex.cs:95-121 without the id de-dupe, which sharding makes unnecessary).

Pure numpy/python: used by bench.py for the slicing arithmetic and by the gloo CPU tests as the reference of
pyrope_topk_merge_device."""
from __future__ import annotations

import numpy as np


def row_block(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [lo, hi) of a FLAT base owned by `rank`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def list_owner(list_id: int, world: int) -> int:
    return list_id % world


def query_slice(nq: int, rank: int, world: int) -> tuple[int, int, int]:
    """-> (lo, hi, per): rank ranks centroids for queries [lo, hi); every rank contributes `per` rows to the
    all-gather (the last ranks pad with -1 rows when world does not divide nq)."""
    per = (nq + world - 1) // world
    lo = min(nq, rank * per)
    hi = min(nq, lo + per)
    return lo, hi, per


def merge_topk(scores: np.ndarray, rows: np.ndarray, k: int):
    """scores/rows [world][nq][k_in] (rows < 0 = empty slot) -> (scores [nq][k], rows [nq][k], counts [nq])."""
    world, nq, kin = scores.shape
    out_s = np.zeros((nq, k), np.float32)
    out_r = np.full((nq, k), -1, np.int64)
    cnt = np.zeros(nq, np.int32)
    for q in range(nq):
        cand = [(-float(scores[w, q, j]), w, j) for w in range(world) for j in range(kin) if rows[w, q, j] >= 0]
        cand.sort()
        for i, (ns, w, j) in enumerate(cand[:k]):
            out_s[q, i] = scores[w, q, j]
            out_r[q, i] = rows[w, q, j]
        cnt[q] = min(k, len(cand))
    return out_s, out_r, cnt
