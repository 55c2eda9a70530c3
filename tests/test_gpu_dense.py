"""GPU parity tests of the dense regime (K4: tcgen05/TMEM contraction with the fused top-k epilogue —
what replaces FAISS's exhaustive_inner_product_blas behind src/search.py:115 for batches of queries)
against the CPU oracle, through the faiss-compatible surface (C ABI).  Both kernels (single CTA and
CTA pairs) and both epilogues (thread-private lists for k <= 32, reservoirs for larger k) are covered."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.flat_ip import check_topk, flat_ip_search


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def f16(x):
    return x.astype(np.float16).astype(np.float64)


def make_index(xb, dtype="fp16", **kw):
    from sgic_b200 import faiss_compat as faiss
    idx = faiss.IndexFlatIP(xb.shape[1], dtype=dtype, device=0, **kw)
    idx.add(xb)
    return idx


# dense_mode: 0 = automatic (up to 128 queries: the transposed kernel when k <= 32 and shared memory allows, else
# the queries-on-M single-CTA kernel; CTA pairs beyond), 1 = queries-on-M single-CTA kernel, 2 = pairs with the
# query tile streamed, 3 = pairs forced (query tile resident), 4 = pairs with the database tile resident (below)
CASES = [
    (256, 64, 5, 4),          # smallest: one tile, d = one K chunk
    (1000, 512, 8, 10),
    (5000, 512, 128, 10),     # exactly one CTA's TMEM lanes
    (5000, 512, 129, 10),     # one query spills into a second tile
    (3001, 520, 17, 10),      # ragged N and d not a multiple of the 64-element K chunk
    (70000, 512, 300, 10),
    (20000, 768, 64, 100),    # ViT-L/14 width, reservoir epilogue
    (40000, 512, 200, 32),    # largest thread-private k
    (40000, 512, 200, 33),    # smallest reservoir k
    (9000, 256, 130, 1),
    (300, 512, 140, 32),      # k < n < one tile
    (300, 512, 140, 100),     # k close to n, reservoirs never fill
    (50, 512, 9, 64),         # k > n: -1 / -FLT_MAX padding out of the dense path
    (100000, 128, 4096, 5),   # a full query block
    (20000, 512, 64, 1024),   # k at the supported maximum (reservoirs of 2048 keys)
    (60000, 256, 129, 500),
    (40000, 512, 300, 200),   # reservoirs of 512 keys: the second register-resident select
    (40000, 768, 40, 256),    # same, single-CTA kernel, k at the top of that range
    (33000, 512, 2, 10),      # transposed kernel: smallest batch, ragged last tile
    (33000, 768, 33, 32),     # transposed: 48 query columns, largest thread-private k
    (33000, 1024, 64, 7),     # transposed: 64 columns, d = 1024
    (100, 64, 3, 10),         # fewer rows than one sub-tile
    (33000, 512, 9, 100),     # transposed with k > 32: lists of 100 keys per (warp, query)
    (33000, 512, 3, 1024),    # transposed at the largest k that fits (n_pad = 16)
]


@pytest.mark.parametrize("mode", [0, 1, 3, 2])
@pytest.mark.parametrize("n,d,nq,k", CASES)
def test_dense_matches_oracle(n, d, nq, k, mode):
    rng = np.random.default_rng(n + d + nq + k)
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    idx = make_index(xb)
    idx.set_option("dense_mode", mode)
    D, I = idx.search(xq, k)
    assert D.shape == (nq, k) and I.shape == (nq, k) and D.dtype == np.float32 and I.dtype == np.int64
    sel = np.arange(nq) if nq <= 48 else np.sort(rng.choice(nq, 48, replace=False))
    # O-exact: fp64 on the fp16-rounded values the GPU holds; O-ref: the fp32 vectors FAISS would hold
    check_topk(D[sel], I[sel], f16(xb), f16(xq[sel]), k, score_tol=3e-5, tie_tol=1e-6)
    check_topk(D[sel], I[sel], xb, xq[sel], k, score_tol=1e-3)
    idx.close()


# dense_mode 4: the pair kernel that keeps a DATABASE tile resident and streams the query block past it
# (scan_dense2b_kernel; picked automatically for several query tiles over a large shard).  Needs > 256 queries
# and d <= 512; per-(CTA, query) reservoirs of 32 keys (k <= 16), 64 keys (k <= 32) or 2*next_pow2(k) beyond.
CASES_B = [
    (70000, 512, 300, 10),
    (100000, 128, 4096, 5),    # a full query block (16 query tiles), two K chunks
    (40000, 512, 600, 32),     # largest k with 64-key reservoirs
    (40000, 512, 257, 16),     # largest k with 32-key reservoirs; one query in the second tile
    (40000, 320, 513, 17),     # five K chunks: the resident ring rotates through its eight slots
    (3001, 512, 300, 10),      # fewer database tiles than CTA pairs, ragged last tile
    (300, 512, 300, 100),      # k close to n: large reservoirs that never fill
    (30000, 256, 1000, 100),   # four K chunks (two tiles resident at once), bisection compaction
    (50, 512, 300, 64),        # k > n: -1 / -FLT_MAX padding
    (150000, 448, 1025, 1),    # seven K chunks, k = 1, five query tiles
    (20000, 512, 300, 1024),   # k at the supported maximum
    (200000, 512, 2048, 10),   # every pair owns several tiles: the slots are refilled under the last query tile
    (30000, 512, 600, 130),    # 512-key reservoirs per (CTA, query)
]


@pytest.mark.parametrize("n,d,nq,k", CASES_B)
def test_dense_database_resident_matches_oracle(n, d, nq, k):
    rng = np.random.default_rng(n + d + nq + k + 4)
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    idx = make_index(xb)
    idx.set_option("dense_mode", 4)
    D, I = idx.search(xq, k)
    assert D.shape == (nq, k) and I.shape == (nq, k)
    # first / last queries of every CTA's 128-query half and a random sample
    edge = [q for q in (0, 127, 128, 255, 256, 383, 384, nq - 1) if q < nq]
    sel = np.unique(np.concatenate([edge, rng.choice(nq, 40, replace=False)])).astype(np.int64)
    check_topk(D[sel], I[sel], f16(xb), f16(xq[sel]), k, score_tol=3e-5, tie_tol=1e-6)
    check_topk(D[sel], I[sel], xb, xq[sel], k, score_tol=1e-3)
    # size-independent property: the query-resident pair kernel must give the same answer
    idx.set_option("dense_mode", 3)
    D3, I3 = idx.search(xq, k)
    np.testing.assert_allclose(D, D3, atol=2e-5, rtol=0)
    assert (I == I3).mean() > 0.999
    idx.close()


def test_dense_database_resident_duplicates_and_bf16():
    rng = np.random.default_rng(43)
    base = unit(rng, 9000, 512)
    xb = np.concatenate([base, base, base])
    xq = base[:400].copy()
    for dtype in ("fp16", "bf16"):
        idx = make_index(xb, dtype)
        idx.set_option("dense_mode", 4)
        D, I = idx.search(xq, 6)
        for r in range(400):
            assert list(I[r, :3]) == [r, r + 9000, r + 18000]
            assert D[r, 0] == D[r, 1] == D[r, 2]
        idx.close()


def test_dense_bf16():
    import torch
    rng = np.random.default_rng(17)
    xb, xq = unit(rng, 30000, 512), unit(rng, 40, 512)
    idx = make_index(xb, "bf16")
    D, I = idx.search(xq, 10)
    r = lambda x: torch.from_numpy(x).to(torch.bfloat16).to(torch.float64).numpy()
    check_topk(D, I, r(xb), r(xq), 10, score_tol=3e-5, tie_tol=1e-6)
    check_topk(D, I, xb, xq, 10, score_tol=4e-3)


@pytest.mark.parametrize("k", [6, 40])
def test_dense_exact_duplicates_resolve_to_lowest_ids(k):
    """Exact ties (FaissDB re-adds every vector on each run, src/compress.py:296-306) must come back in
    (score desc, id asc) order from the tensor-core path too — bit-identical ids to the fp64 oracle."""
    rng = np.random.default_rng(23)
    base = unit(rng, 700, 512)
    xb = np.concatenate([base, base, base])
    xq = base[:20].copy()
    idx = make_index(xb)
    D, I = idx.search(xq, k)
    for r in range(20):
        assert list(I[r, :3]) == [r, r + 700, r + 1400]
        assert D[r, 0] == D[r, 1] == D[r, 2]
    check_topk(D, I, f16(xb), f16(xq), k, score_tol=3e-5, tie_tol=1e-6)


@pytest.mark.parametrize("k", [1, 2, 3, 40])
def test_dense_ties_across_items_full_block(k):
    """A full 4096-query block runs 8 rounds of (query tile, slice) items; later items inherit the score bound the
    earlier ones published (gthr).  Exact copies of a row sit in different slices, so the k-th boundary is a tie
    between lists: the answer must still be the lowest row numbers, identical with the inheritance switched off."""
    rng = np.random.default_rng(47)
    base = unit(rng, 9000, 512)
    xb = np.concatenate([base, base, base])
    xq = base[:4096].copy()
    idx = make_index(xb)
    D, I = idx.search(xq, k)
    want = np.stack([np.arange(4096) + 9000 * j for j in range(3)], axis=1)[:, :min(k, 3)]
    assert np.array_equal(I[:, :min(k, 3)], want)
    idx.set_option("dense_gthr", 0)
    D0, I0 = idx.search(xq, k)
    assert np.array_equal(I, I0) and np.array_equal(D, D0)
    sel = np.arange(0, 4096, 97)
    check_topk(D[sel], I[sel], f16(xb), f16(xq[sel]), k, score_tol=3e-5, tie_tol=1e-6)
    idx.close()


@pytest.mark.parametrize("nq,k,mode", [(150, 300, 0), (150, 600, 0), (300, 1000, 0), (40, 700, 1), (300, 300, 4), (300, 200, 0)])
def test_large_k_with_triplicated_rows(nq, k, mode):
    """Every row exists three times, so every score in a reservoir ties with two others and the k-th place falls
    inside a tie group for most k: the select that works on score words only (1024 / 2048-key reservoirs) must
    settle those ties by row number exactly like the sorted-list paths do."""
    rng = np.random.default_rng(nq + k)
    base = unit(rng, 700, 512)
    xb = np.concatenate([base, base, base])
    xq = unit(rng, nq, 512)
    idx = make_index(xb)
    idx.set_option("dense_mode", mode)
    D, I = idx.search(xq, k)
    idx.close()
    for r in range(0, nq, 7):
        ids = I[r]
        assert len(set(ids.tolist())) == k and ids.min() >= 0
        pairs = list(zip((-D[r]).tolist(), ids.tolist()))
        assert pairs == sorted(pairs)                                   # (score desc, id asc)
        groups = {}
        for i in ids.tolist():
            groups.setdefault(i % 700, []).append(i)
        for g, members in groups.items():                                # copies enter lowest row number first
            assert sorted(members) == [g + 700 * j for j in range(len(members))]
    sel = np.arange(0, nq, 11)
    check_topk(D[sel], I[sel], f16(xb), f16(xq[sel]), k, score_tol=3e-5, tie_tol=1e-6)


def test_dense_agrees_with_streaming_kernel():
    """Size-independent property: the same queries answered one at a time by K3 (CUDA-core streaming scan) and
    as one batch by K4 (tensor cores) give the same ids; scores agree to fp32 accumulation-order noise."""
    rng = np.random.default_rng(29)
    n, d, nq, k = 400_000, 512, 96, 10
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    idx = make_index(xb, retain_fp32=False)
    D4, I4 = idx.search(xq, k)
    idx.set_option("dense_min_nq", 1 << 30)   # force K3 for every batch size
    D3 = np.empty_like(D4)
    I3 = np.empty_like(I4)
    for i in range(nq):
        D3[i:i + 1], I3[i:i + 1] = idx.search(xq[i:i + 1], k)
    np.testing.assert_allclose(D4, D3, atol=2e-5, rtol=0)
    agree = (I4 == I3).mean()
    assert agree > 0.999, agree  # an id may swap only between scores closer than the accumulation noise
    assert np.all(np.diff(D4, axis=1) <= 0)


def test_dense_self_queries_full_block():
    """Database rows used as queries come back at rank 1 with score ~1 (round trip through both operands'
    fp16 rounding), for a whole 4096-query block over a database larger than L2."""
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.synth import fill_index_random
    n, d, nq, k = 3_000_000, 512, 4096, 10
    idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
    fill_index_random(idx, n)
    rng = np.random.default_rng(31)
    starts = np.sort(rng.choice(n // 512 - 1, 8, replace=False)) * 512
    ids = np.concatenate([np.arange(s0, s0 + 512) for s0 in starts])
    q = np.concatenate([idx.reconstruct_n(int(s0), 512) for s0 in starts])
    D, I = idx.search(q, k)
    assert np.array_equal(I[:, 0], ids)
    assert np.all(np.abs(D[:, 0] - 1.0) < 2e-3)
    assert np.all(D[:, 1:] < 0.5)  # random unit vectors in 512 dimensions are nearly orthogonal
    assert np.all(D[:, :-1] >= D[:, 1:])


def test_mixed_batch_sizes_share_the_workspace():
    """A resident service alternates batch sizes on one index: the streaming kernel's grid counter and the dense
    kernels' partial lists live in the same workspace and must not disturb each other (K3 -> K4 -> K3 -> K4t)."""
    rng = np.random.default_rng(41)
    xb, xq = unit(rng, 60_000, 512), unit(rng, 300, 512)
    idx = make_index(xb)
    ref = {}
    for nq in (1, 300, 1, 7, 1, 64, 2, 1):
        D, I = idx.search(xq[:nq], 10)
        check_topk(D[:8], I[:8], f16(xb), f16(xq[:min(nq, 8)]), 10, score_tol=3e-5, tie_tol=1e-6)
        if nq in ref:
            assert np.array_equal(ref[nq][1], I) and np.array_equal(ref[nq][0], D)
        ref[nq] = (D, I)


def test_sample_pass_does_not_change_large_k_answers():
    """k > 32: lists may start from the k-th best score of a SAMPLE of the rows instead of -inf (dense_seed).  The bound
    is valid whatever the sample looks like — including a sample that holds the best rows, duplicates that tie with the
    k-th score, and fewer than k good rows — so answers must be bit-identical with the pass forced on and off."""
    rng = np.random.default_rng(53)
    n, d, nq, k = 300_000, 512, 300, 100
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    xb[:60] = xq[0] * 0.9 + 0.1 * xb[:60]                 # the sample (first rows) holds query 0's best rows ...
    xb[200_000:200_040] = xb[:40]                           # ... and 40 of them again far outside it: exact ties
    xb[250_000:250_200] = xq[1] * 0.95 + 0.05 * xb[250_000:250_200]   # query 1's best rows lie outside the sample
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    idx = make_index(xb, retain_fp32=False)
    idx.set_option("dense_seed", 0)
    D0, I0 = idx.search(xq, k)
    idx.set_option("dense_seed", 2)
    launches = idx.stat("launches")
    D2, I2 = idx.search(xq, k)
    assert idx.stat("launches") - launches >= 6             # the sample search + seed kernel ran in front of the real one
    assert np.array_equal(I0, I2) and np.array_equal(D0, D2)
    sel = np.array([0, 1, 2, 150, 299])
    check_topk(D2[sel], I2[sel], f16(xb), f16(xq[sel]), k, score_tol=3e-5, tie_tol=1e-6)
    assert set(I2[0].tolist()) == set(range(60)) | set(range(200_000, 200_040))            # the 60 planted rows + the 40 copies
    for j in range(40):                                     # ties in row order
        a, b = np.where(I2[0] == j)[0][0], np.where(I2[0] == 200_000 + j)[0][0]
        assert a < b and D2[0, a] == D2[0, b]
    idx.set_option("dense_seed", 1)
