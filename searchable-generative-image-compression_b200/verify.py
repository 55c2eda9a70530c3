"""Independent re-score of search answers at full scale (SURVEY.md §4.3 tier T3).

The parity tests compare the kernels with the CPU oracle at sizes the oracle finishes in seconds; at the sizes
``bench.py`` times (100M rows and more) the answer is checked here instead: the rows are read straight out of the
index's HBM allocation (a zero-copy ``torch`` view through ``__cuda_array_interface__``), scored against the
sampled queries by a plain chunked ``torch.matmul`` in fp32 followed by ``torch.topk`` — no kernel of this
package takes part — and the candidates are then re-scored in fp64.  What comes out is the record ``bench.py``
prints as ``parity``: how many queries were checked, the largest score error against fp64 on the stored values,
and how many returned ids are wrong beyond ties at the k-th score (must be 0).

The reference call being checked is ``index.search(q, k)`` (src/search.py:115): the k largest inner products,
sorted descending, ids = row numbers.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

__all__ = ["database_view", "rescore_topk", "compare_topk", "verify_search"]


class _DevArray:
    """Minimal ``__cuda_array_interface__`` carrier for a raw device pointer."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def database_view(index):
    """The index's rows as a ``(ntotal, d)`` CUDA tensor that aliases the HBM database (fp16 rows: exact view;
    bf16 rows: viewed as int16 and re-interpreted).  Valid until the next add / reserve reallocates."""
    import torch
    ptr = index._lib.sgic_index_data_dev(index._h)
    n, d = index.ntotal, index.d
    if n == 0 or not ptr:
        return torch.empty((0, d), dtype=torch.float16, device=f"cuda:{index.device}")
    with torch.cuda.device(index.device):
        if index.dtype == "bf16":
            t = torch.as_tensor(_DevArray(ptr, (n, d), "<i2"), device=f"cuda:{index.device}")
            return t.view(torch.bfloat16)
        return torch.as_tensor(_DevArray(ptr, (n, d), "<f2"), device=f"cuda:{index.device}")


def rescore_topk(index, q, k: int, *, margin: int = 16, chunk_rows: int = 1 << 20, id_base: int = 0):
    """Reference top-``k + margin`` of the fp32 queries ``q`` (CUDA tensor, (nq, d)) over every row the index
    holds: chunked fp32 ``matmul`` + ``topk`` + merge, on the stored (rounded) values and the queries rounded
    to the storage dtype the same way the kernels round them.  Returns ``(scores, ids)`` CUDA tensors of shape
    (nq, min(k + margin, ntotal)), best first, ties by ascending id, ids offset by ``id_base``."""
    import torch
    db = database_view(index)
    n, d = db.shape
    nq = q.shape[0]
    kk = min(k + margin, n)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        qr = q.to(db.dtype).float()                     # the rounding add() / the kernels apply to queries
        best_s = torch.full((nq, 0), 0.0, device=q.device)
        best_i = torch.zeros((nq, 0), dtype=torch.int64, device=q.device)
        for r0 in range(0, n, chunk_rows):
            blk = db[r0:r0 + chunk_rows].float()
            s = qr @ blk.t()
            ks = min(kk, s.shape[1])
            cs, ci = torch.topk(s, ks, dim=1)
            best_s = torch.cat([best_s, cs], dim=1)
            best_i = torch.cat([best_i, ci + r0], dim=1)
            if best_s.shape[1] > 4 * kk:                # keep the running candidate list short
                ts, tj = torch.topk(best_s, kk, dim=1)
                best_s, best_i = ts, torch.gather(best_i, 1, tj)
            del blk, s
        if kk == 0:
            return best_s, best_i
        ts, tj = torch.topk(best_s, min(kk, best_s.shape[1]), dim=1)
        ti = torch.gather(best_i, 1, tj)
        # exact order: fp64 scores of the surviving candidates, ties by ascending id
        rows = db[ti.reshape(-1)].double().view(nq, ti.shape[1], d)
        s64 = torch.einsum("qd,qcd->qc", qr.double(), rows)
        order = np.lexsort((ti.cpu().numpy(), -s64.cpu().numpy()), axis=1)
        order = torch.from_numpy(order).to(q.device)
        return torch.gather(s64, 1, order), torch.gather(ti, 1, order) + id_base
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def compare_topk(D, I, ref_s, ref_i, k: int, *, score_tol: float, tie_tol: float = 2e-6) -> Dict:
    """Answer ``(D, I)`` (host arrays (nq, k)) against a reference candidate list ``(ref_s, ref_i)`` ((nq, >= k)
    fp64 scores best first, ids; from :func:`rescore_topk` possibly merged over shards).

    * every returned score is within ``score_tol`` of the reference score of the SAME id (ids the reference list
      does not hold count as wrong unless their own score ties with the k-th);
    * rows are sorted descending;
    * the id sets agree except for candidates within ``tie_tol`` of the reference's k-th score.
    """
    D, I = np.asarray(D), np.asarray(I)
    ref_s, ref_i = np.asarray(ref_s, dtype=np.float64), np.asarray(ref_i)
    nq = D.shape[0]
    max_err, wrong, unsorted_rows, pad_err = 0.0, 0, 0, 0
    for r in range(nq):
        valid = int((ref_i[r] >= 0).sum())
        kv = min(k, valid)
        if kv < k:                                     # fewer rows than k: faiss padding
            pad_err += int((I[r, kv:] != -1).sum())
        if kv == 0:
            continue
        got_i, got_s = I[r, :kv], D[r, :kv].astype(np.float64)
        if np.any(np.diff(got_s) > 0):
            unsorted_rows += 1
        ref_map = {int(i): float(s) for i, s in zip(ref_i[r], ref_s[r])}
        kth = ref_s[r, kv - 1]
        want = set(int(i) for i in ref_i[r, :kv])
        for i, s in zip(got_i, got_s):
            i = int(i)
            if i in ref_map:
                max_err = max(max_err, abs(s - ref_map[i]))
                if i not in want and ref_map[i] < kth - tie_tol:
                    wrong += 1
            elif s < kth - tie_tol - score_tol:        # not even a candidate of the reference and clearly below
                wrong += 1
        have = set(int(i) for i in got_i)
        for i in want - have:
            if ref_map[i] > kth + tie_tol:             # a clear winner is missing
                wrong += 1
        if len(have) != kv:
            wrong += kv - len(have)                    # duplicates
    return {"queries": int(nq), "k": int(k), "max_score_err": float(max_err), "ids_outside_ties": int(wrong),
            "unsorted_rows": int(unsorted_rows), "padding_errors": int(pad_err),
            "ok": bool(wrong == 0 and unsorted_rows == 0 and pad_err == 0 and max_err <= score_tol)}


def verify_search(index, q, D, I, k: int, *, sample: Optional[np.ndarray] = None, score_tol: float = 3e-5,
                  id_base: int = 0) -> Tuple[Dict, Tuple]:
    """Re-score the sampled query rows of ``q`` (CUDA fp32 (nq, d)) on ``index`` (one GPU's rows) and compare with
    the answer ``(D, I)`` (CUDA or host).  Returns ``(record, (ref_s, ref_i))``."""
    import torch
    nq = q.shape[0]
    sel = np.arange(nq) if sample is None else np.asarray(sample)
    st = torch.from_numpy(sel).to(q.device)
    ref_s, ref_i = rescore_topk(index, q[st].contiguous(), k, id_base=id_base)
    Dh = D[st].cpu().numpy() if hasattr(D, "cpu") else np.asarray(D)[sel]
    Ih = I[st].cpu().numpy() if hasattr(I, "cpu") else np.asarray(I)[sel]
    rec = compare_topk(Dh, Ih, ref_s.cpu().numpy(), ref_i.cpu().numpy(), k, score_tol=score_tol)
    return rec, (ref_s, ref_i)
