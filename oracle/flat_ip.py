"""numpy restatement of FAISS ``IndexFlatIP`` search semantics (test oracle).

Follows the call sites in the reference (src/search.py:113-120 ``do_search`` →
``index.search(q, k)``; src/build.py:93-94; src/compress.py:97,107) and the
upstream FAISS algorithm they reach (``IndexFlat::search`` →
``knn_inner_product``; not vendored, restated from the published source, see
SURVEY.md §8c):

* scores are fp32 inner products of fp32 rows;
* per query the k largest scores are returned **sorted descending**;
* ids are int64 row numbers; missing slots are ``I=-1``, ``D=-FLT_MAX``;
* a candidate replaces the current k-th best only if strictly greater, so among
  exact ties the earlier row survives (heap path, k < 100).

The oracle resolves every tie as (score desc, id asc).  FAISS's order *among*
equal scores is implementation-defined, so the parity checker
(:func:`check_topk`) accepts any permutation inside a tie group and any choice
of members at the k-th score boundary.
"""
from __future__ import annotations

import numpy as np

NEG_FLT_MAX = np.float32(-3.4028234663852886e38)


def _topk_rows(scores: np.ndarray, k: int, id0: int = 0):
    """Top-k of each row of ``scores`` ordered (score desc, id asc)."""
    nq, n = scores.shape
    kk = min(k, n)
    D = np.full((nq, k), NEG_FLT_MAX, dtype=scores.dtype)
    I = np.full((nq, k), -1, dtype=np.int64)
    if kk == 0:
        return D, I
    for qi in range(nq):
        s = scores[qi]
        if kk < n:
            # threshold = kk-th largest value; keep everything >= it, then order
            thr = np.partition(s, n - kk)[n - kk]
            cand = np.nonzero(s >= thr)[0]
        else:
            cand = np.arange(n)
        order = np.lexsort((cand, -s[cand].astype(np.float64)))[:kk]
        sel = cand[order]
        D[qi, :kk] = s[sel]
        I[qi, :kk] = sel + id0
    return D, I


def _merge(Da, Ia, Db, Ib, k):
    D = np.concatenate([Da, Db], axis=1)
    I = np.concatenate([Ia, Ib], axis=1)
    nq = D.shape[0]
    Do = np.full((nq, k), NEG_FLT_MAX, dtype=D.dtype)
    Io = np.full((nq, k), -1, dtype=np.int64)
    for qi in range(nq):
        valid = np.nonzero(I[qi] >= 0)[0]
        order = np.lexsort((I[qi, valid], -D[qi, valid].astype(np.float64)))[:k]
        sel = valid[order]
        Do[qi, : sel.size] = D[qi, sel]
        Io[qi, : sel.size] = I[qi, sel]
    return Do, Io


def flat_ip_search(xb: np.ndarray, xq: np.ndarray, k: int, *, dtype=np.float32,
                   block: int = 262144):
    """Exact inner-product k-NN, FAISS ``IndexFlatIP.search`` semantics.

    ``dtype=np.float32`` is the O-ref oracle (what FAISS computes on the fp32
    rows); ``dtype=np.float64`` is the O-exact oracle used on the fp16/bf16
    rounded values the GPU actually stores (SURVEY.md §8c "two tiers").
    """
    assert k > 0, "k must be positive"  # faiss: FAISS_THROW_IF_NOT(k > 0)
    xb = np.ascontiguousarray(xb, dtype=dtype)
    xq = np.ascontiguousarray(xq, dtype=dtype)
    assert xq.ndim == 2 and xb.ndim == 2 and xq.shape[1] == xb.shape[1]
    nq, n = xq.shape[0], xb.shape[0]
    D = np.full((nq, k), NEG_FLT_MAX, dtype=dtype)
    I = np.full((nq, k), -1, dtype=np.int64)
    for j0 in range(0, n, block):
        s = xq @ xb[j0:j0 + block].T
        Db, Ib = _topk_rows(s, k, id0=j0)
        D, I = _merge(D, I, Db, Ib, k)
    return D.astype(np.float32), I


def check_topk(D_got, I_got, xb, xq, k, *, score_tol: float, dtype=np.float64,
               tie_tol: float | None = None):
    """Parity checker used by the GPU tests.

    * every returned score is within ``score_tol`` of the oracle score **of the
      returned id** (so a wrong id with a plausible score fails);
    * scores are sorted descending;
    * the returned id set equals the oracle's top-k set except for candidates
      whose oracle score is within ``tie_tol`` of the oracle's k-th score
      (north_star: "identical except for ties at the k-th score boundary");
    * missing slots (ntotal < k) are ``-1`` / ``-FLT_MAX``.

    Raises AssertionError with a description; returns the number of boundary
    substitutions it accepted.
    """
    if tie_tol is None:
        tie_tol = score_tol
    xb = np.asarray(xb)
    xq = np.asarray(xq)
    D_got = np.asarray(D_got)
    I_got = np.asarray(I_got)
    nq, n = xq.shape[0], xb.shape[0]
    assert D_got.shape == (nq, k) and I_got.shape == (nq, k), (D_got.shape, I_got.shape)
    assert D_got.dtype == np.float32 and I_got.dtype == np.int64
    kk = min(k, n)
    swaps = 0
    xb64 = xb.astype(dtype)
    for qi in range(nq):
        s = xb64 @ xq[qi].astype(dtype)  # oracle scores of every row
        ids = I_got[qi]
        assert np.all(ids[kk:] == -1), f"q{qi}: padding ids {ids[kk:]}"
        assert np.all(D_got[qi, kk:] == NEG_FLT_MAX), f"q{qi}: padding scores"
        got = ids[:kk]
        assert np.all((got >= 0) & (got < n)), f"q{qi}: id out of range {got}"
        assert np.unique(got).size == kk, f"q{qi}: duplicate ids {got}"
        err = np.abs(D_got[qi, :kk].astype(dtype) - s[got])
        assert err.max(initial=0.0) <= score_tol, f"q{qi}: score err {err.max():.3e} > {score_tol}"
        dd = D_got[qi, :kk]
        assert np.all(dd[:-1] >= dd[1:]), f"q{qi}: not sorted descending"
        if kk == 0:
            continue
        # oracle top-kk
        if kk < n:
            kth = np.partition(s, n - kk)[n - kk]
        else:
            kth = s.min()
        must = np.nonzero(s > kth + tie_tol)[0]      # clearly inside
        may = np.nonzero(s >= kth - tie_tol)[0]      # inside or on the boundary
        missing = np.setdiff1d(must, got)
        assert missing.size == 0, f"q{qi}: missing clear top-k ids {missing[:8]}"
        extra = np.setdiff1d(got, may)
        assert extra.size == 0, f"q{qi}: ids below the k-th boundary {extra[:8]}"
        exact = np.lexsort((np.arange(n), -s))[:kk]
        swaps += int(np.setdiff1d(got, exact).size)
    return swaps
