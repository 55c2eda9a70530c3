"""The comparator behind bench.py's `parity` record and the full-scale GPU tests (sgic_b200/verify.py), on CPU:
it must accept exactly what north_star allows (score noise, swaps among ties at the k-th score) and flag the rest."""
import numpy as np

from oracle.flat_ip import flat_ip_search
from sgic_b200.verify import compare_topk


def _case(seed=0, n=5000, d=32, nq=6, k=10, margin=8):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    s = xq.astype(np.float64) @ xb.astype(np.float64).T
    order = np.lexsort((np.broadcast_to(np.arange(n), s.shape), -s), axis=1)[:, :k + margin]
    ref_i = order
    ref_s = np.take_along_axis(s, order, 1)
    D, I = flat_ip_search(xb, xq, k)
    return D, I, ref_s, ref_i


def test_exact_answer_passes_and_reports_the_noise():
    D, I, ref_s, ref_i = _case()
    rec = compare_topk(D, I, ref_s, ref_i, 10, score_tol=1e-5)
    assert rec["ok"] and rec["queries"] == 6 and rec["ids_outside_ties"] == 0 and rec["max_score_err"] < 1e-6


def test_wrong_ids_wrong_scores_and_disorder_are_flagged():
    D, I, ref_s, ref_i = _case()
    bad = I.copy()
    bad[2, 4] = ref_i[2, -1]                      # a candidate clearly below the k-th score
    assert compare_topk(D, bad, ref_s, ref_i, 10, score_tol=1e-5)["ids_outside_ties"] >= 1
    bad = I.copy()
    bad[1, 3] = 4999 if 4999 not in ref_i[1] else 4998     # an id the reference never saw, with a winner's score
    rec = compare_topk(D, bad, ref_s, ref_i, 10, score_tol=1e-5)
    assert not rec["ok"] and rec["ids_outside_ties"] >= 1
    off = D.copy()
    off[0, 0] += 1e-3
    rec = compare_topk(off, I, ref_s, ref_i, 10, score_tol=1e-5)
    assert not rec["ok"] and rec["max_score_err"] > 9e-4
    sw = D.copy()
    sw[3, [0, 5]] = sw[3, [5, 0]]
    assert compare_topk(sw, I, ref_s, ref_i, 10, score_tol=1.0)["unsorted_rows"] == 1
    dup = I.copy()
    dup[4, 1] = dup[4, 0]
    assert not compare_topk(D, dup, ref_s, ref_i, 10, score_tol=1e-5)["ok"]


def test_swaps_among_ties_at_the_kth_score_are_accepted_and_padding_checked():
    k = 4
    ref_s = np.array([[0.9, 0.8, 0.7, 0.5, 0.5, 0.5, 0.1]])
    ref_i = np.array([[10, 11, 12, 3, 7, 9, 2]])
    D = np.array([[0.9, 0.8, 0.7, 0.5]], dtype=np.float32)
    for last in (3, 7, 9):                         # any member of the tie group is a valid 4th answer
        assert compare_topk(D, np.array([[10, 11, 12, last]]), ref_s, ref_i, k, score_tol=1e-6)["ok"]
    assert not compare_topk(D, np.array([[10, 11, 12, 2]]), ref_s, ref_i, k, score_tol=1e-6)["ok"]
    assert not compare_topk(D, np.array([[10, 11, 3, 7]]), ref_s, ref_i, k, score_tol=1e-6)["ok"]   # 12 is a clear winner
    # fewer rows than k: -1 padding expected
    ref_s, ref_i = np.array([[0.9, 0.8, 0.0, 0.0]]), np.array([[1, 0, -1, -1]])
    D = np.array([[0.9, 0.8, -3.4028235e38, -3.4028235e38]], dtype=np.float32)
    assert compare_topk(D, np.array([[1, 0, -1, -1]]), ref_s, ref_i, 4, score_tol=1e-6)["ok"]
    assert compare_topk(D, np.array([[1, 0, 5, -1]]), ref_s, ref_i, 4, score_tol=1e-6)["padding_errors"] == 1
