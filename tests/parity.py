"""Tolerance-aware comparison of top-k lists (north_star: distances within 1e-4 relative; id lists
identical except where adjacent distances differ by less than that tolerance)."""
import numpy as np

RTOL = 1e-4  # relative tolerance on distances/scores stated by BASELINE.json north_star


def _tol(a, b, rtol, atol):
    return rtol * max(abs(float(a)), abs(float(b))) + atol


def assert_topk_equivalent(ref_ids, ref_scores, got_ids, got_scores, rtol=RTOL, atol=1e-6, ctx=""):
    ref_ids, got_ids = np.asarray(ref_ids), np.asarray(got_ids)
    ref_scores, got_scores = np.asarray(ref_scores, np.float64), np.asarray(got_scores, np.float64)
    assert len(ref_ids) == len(got_ids), f"{ctx}: result count {len(got_ids)} != reference {len(ref_ids)}"
    n = len(ref_ids)
    for i in range(n):
        assert abs(got_scores[i] - ref_scores[i]) <= _tol(got_scores[i], ref_scores[i], rtol, atol), \
            f"{ctx}: score[{i}] {got_scores[i]!r} vs reference {ref_scores[i]!r}"
    for i in range(n - 1):
        assert got_scores[i] >= got_scores[i + 1], f"{ctx}: scores not sorted descending at {i}"
    ref_pos = {int(r): j for j, r in enumerate(ref_ids)}
    for i in range(n):
        g = int(got_ids[i])
        if g == int(ref_ids[i]):
            continue
        if g in ref_pos:  # swapped with a near-tie inside the list
            j = ref_pos[g]
            assert abs(ref_scores[j] - ref_scores[i]) <= 2 * _tol(ref_scores[j], ref_scores[i], rtol, atol), \
                f"{ctx}: id {g} at rank {i} but reference has it at rank {j} with a non-tied score"
        else:  # boundary tie: a candidate the reference dropped at the k-th place
            assert abs(got_scores[i] - ref_scores[n - 1]) <= 2 * _tol(got_scores[i], ref_scores[n - 1], rtol, atol), \
                f"{ctx}: id {g} at rank {i} is not in the reference list and not tied with its last score"
    assert len(set(int(x) for x in got_ids)) == n, f"{ctx}: duplicate ids in result"


def assert_batch_equivalent(ref, got, ctx=""):
    """ref/got: (ids [nq][k], scores [nq][k], counts [nq])"""
    rid, rsc, rcn = ref
    gid, gsc, gcn = got
    assert len(rcn) == len(gcn)
    for q in range(len(rcn)):
        assert int(rcn[q]) == int(gcn[q]), f"{ctx} q{q}: count {int(gcn[q])} != {int(rcn[q])}"
        c = int(rcn[q])
        assert_topk_equivalent(rid[q][:c], rsc[q][:c], gid[q][:c], gsc[q][:c], ctx=f"{ctx} q{q}")


def recall_at_k(truth_ids, got_ids, k):
    hit = 0
    for t, g in zip(truth_ids, got_ids):
        hit += len(set(int(x) for x in t[:k]) & set(int(x) for x in g[:k]))
    return hit / (len(truth_ids) * k)
