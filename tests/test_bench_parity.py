"""Host logic of bench.py's `parity` block (no GPU): the comparison it embeds in every bench line follows
tests/parity.py — scores within 1e-4 relative, ids identical except across near-ties."""
import numpy as np

import bench


def _res(ids, sc):
    ids, sc = np.asarray(ids, np.int64), np.asarray(sc, np.float32)
    return ids, sc, np.full(len(ids), ids.shape[1], np.int32)


def test_identical_results_are_green():
    r = _res([[1, 2, 3], [4, 5, 6]], [[-1.0, -2.0, -3.0], [-1.5, -2.5, -3.5]])
    out = bench.compare_topk(r, r, "self")
    assert out["mismatch"] == 0 and out["queries"] == 2 and out["identical_id_lists"] == 2 and out["max_rel_err"] == 0.0


def test_near_tie_swap_is_accepted_and_a_wrong_id_is_not():
    ref = _res([[1, 2, 3]], [[-1.0, -2.0, -2.00001]])
    swapped = _res([[1, 3, 2]], [[-1.0, -2.0, -2.00001]])
    assert bench.compare_topk(ref, swapped, "x")["mismatch"] == 0
    wrong = _res([[1, 2, 9]], [[-1.0, -2.0, -2.5]])
    out = bench.compare_topk(ref, wrong, "x")
    assert out["mismatch"] == 1 and "first_mismatch" in out


def test_score_outside_tolerance_and_count_mismatch():
    ref = _res([[1, 2]], [[-1.0, -2.0]])
    off = _res([[1, 2]], [[-1.0, -2.001]])
    assert bench.compare_topk(ref, off, "x")["mismatch"] == 1
    ids, sc, cnt = _res([[1, 2]], [[-1.0, -2.0]])
    assert bench.compare_topk(ref, (ids, sc, np.array([1], np.int32)), "x")["mismatch"] == 1


def test_both_arms_print_the_same_config_object():
    w = bench.workload("c5", 1.0)
    assert bench.config_of(w) == {"workload": bench.describe(w), "reduced": False}
    assert bench.host_cores() >= 1
