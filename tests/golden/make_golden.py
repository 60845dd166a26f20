#!/usr/bin/env python
"""Regenerate tests/golden/*.npz.

The reference (takurot/Pyrope) is C#/.NET and cannot run in this image, and it ships no golden top-k
lists for this path, so these vectors come from the CPU oracle (oracle/oracle.c), whose
own validity rests on the reference's unit-test cases and the System.Random known answers
(tests/test_oracle_reference_cases.py).  They freeze the BASELINE.json configs 1-3 at the reference's
synthetic inputs (System.Random seeds 42 / 1337, Pyrope.Benchmarks/Program.cs:219-263) so that any later
change to the oracle OR to the CUDA path shows up as a diff against a committed file.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as orc  # noqa: E402

N, NQ, DIM, K = 10_000, 100, 128, 10


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    base = orc.random_vectors(N, DIM, 42)
    q = orc.random_vectors(NQ, DIM, 1337)
    out = {"base_sha256": sha(base), "query_sha256": sha(q),
           "base_head": base[:2, :8].copy(), "query_head": q[:2, :8].copy()}

    # C1: FLAT L2 / IP / Cosine, TOPK 10 (BruteForceVectorIndex.Search)
    for name, metric in (("l2", orc.L2), ("ip", orc.IP), ("cos", orc.COSINE)):
        ix = orc.FlatIndex(DIM, metric)
        ix.add_batch(base)
        ids, sc, cnt = ix.search_batch(q, K)
        out[f"c1_{name}_ids"], out[f"c1_{name}_scores"] = ids, sc

    # C2: IVF_FLAT nlist=100, default nprobe 3 (IvfFlatVectorIndex Build + Search)
    ivf = orc.IvfFlatIndex(DIM, orc.L2, nlist=100)
    ivf.add_batch(base)
    ivf.build()
    cent = ivf.centroids()
    out["c2_centroids_sha256"] = sha(cent)
    out["c2_centroids_head"] = cent[:2, :8].copy()
    out["c2_assign"] = np.array([orc.find_nearest_centroid(v, cent, orc.L2) for v in base], np.int32)
    for nprobe in (3, 10):
        ids, sc, cnt = ivf.search_batch(q, K, nprobe=nprobe)
        out[f"c2_np{nprobe}_ids"], out[f"c2_np{nprobe}_scores"] = ids, sc

    # C3: IVF_PQ nlist=100 m=4 k=256, default nprobe 1 (IvfPqVectorIndex Build + Search)
    pq = orc.IvfPqIndex(DIM, orc.L2, m=4, k=256, nlist=100)
    pq.add_batch(base)
    pq.build()
    out["c3_centroids_sha256"] = sha(pq.centroids())
    out["c3_codebook_sha256"] = sha(pq.pq().codebook())
    lists = pq.lists()
    out["c3_list_sizes"] = np.array([len(i) for i, _ in lists], np.int32)
    out["c3_codes_sha256"] = sha(np.concatenate([c for _, c in lists]))
    out["c3_ids_sha256"] = sha(np.concatenate([i for i, _ in lists]))
    for nprobe in (1, 8):
        ids, sc, cnt = pq.search_batch(q, K, nprobe=nprobe)
        out[f"c3_np{nprobe}_ids"], out[f"c3_np{nprobe}_scores"] = ids, sc

    # m=16 variant at the C5 sub-vector shape (sub = 8): the list-major kernel's configuration
    pq16 = orc.IvfPqIndex(DIM, orc.L2, m=16, k=256, nlist=32)
    pq16.add_batch(base)
    pq16.build()
    l16 = pq16.lists()
    out["m16_codes_sha256"] = sha(np.concatenate([c for _, c in l16]))
    ids, sc, cnt = pq16.search_batch(q, K, nprobe=8)
    out["m16_np8_ids"], out["m16_np8_scores"] = ids, sc

    path = os.path.join(HERE, "configs_1_2_3.npz")
    np.savez_compressed(path, **{k: (np.array(v) if isinstance(v, str) else v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
