"""Full-scale parity (SURVEY.md §4.3 tier T3): the answers at the sizes ``bench.py`` times.

* C2 (1M x 512, batch 1, k = 10) and C3 (10M x 768, batch 4096, k = 100): complete comparison with the C
  restatement of FAISS (``oracle/flat_ip.c``) on the values the GPU holds (fp16-rounded rows and queries), the
  rows brought back chunk by chunk with ``reconstruct_n``.
* C4 (100M x 512 = 102.4 GB, one GPU): a shard beyond the 90 GB switch, so the batch kernels are auto-dispatched
  to ``scan_dense2b_kernel`` (K4b) and K3 runs its 3.1M-tile schedule with the shared tail.  Every batch size the
  bench times (1, 128, 256, 1024, 4096) is checked by the independent torch re-score of ``sgic_b200.verify`` (chunked
  fp32 matmul + topk over the HBM rows, candidates re-scored in fp64), and the C oracle confirms the winners on
  the candidate superset.

The reference call is ``index.search(q, k)`` — src/search.py:115.  Tolerances: 3e-5 against fp64 / the fp32
oracle on the stored values (fp32 accumulation-order noise), ids identical except for ties at the k-th score.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.flat_ip_c import flat_ip_search_c


def f16(x):
    return x.astype(np.float16).astype(np.float32)


def c2df_style_queries(index, rows):
    """Database rows pushed through the reference quantiser + dequantiser (compress.py:77 -> search.py:20-22):
    what ``query-c2df`` sends; the expected top-1 is the row itself at ~0.9987."""
    from oracle import c2df_ref
    out = []
    for r in rows:
        v = index.reconstruct_n(int(r), 1)[0]
        v = v / np.linalg.norm(v)
        q = np.clip(np.round((v * 0.5 + 0.5) * 255.0), 0, 255).astype(np.uint8)
        out.append(c2df_ref.dequantize_clip_u8(q))
    return np.stack(out).astype(np.float32)


def oracle_topk_chunked(index, xq, k, chunk_rows=1 << 20):
    """oracle/flat_ip.c over every row of the index, chunk by chunk; per-chunk answers merged by
    (score desc, id asc) — the exact top-k of the whole database as the FAISS restatement scores it."""
    n = index.ntotal
    Ds, Is = [], []
    for r0 in range(0, n, chunk_rows):
        xb = index.reconstruct_n(r0, min(chunk_rows, n - r0))
        D, I = flat_ip_search_c(xb, xq, min(k, xb.shape[0]))
        Ds.append(D)
        Is.append(np.where(I >= 0, I + r0, I))
    D, I = np.concatenate(Ds, axis=1), np.concatenate(Is, axis=1)
    order = np.lexsort((I, -D.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(D, order, 1), np.take_along_axis(I, order, 1)


def assert_parity(D, I, Do, Io, k, tol=3e-5):
    from sgic_b200.verify import compare_topk
    rec = compare_topk(D, I, Do, Io, k, score_tol=tol, tie_tol=tol)
    assert rec["ok"], rec
    return rec


def test_c2_1m_x_512_full_compare_with_c_oracle():
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.synth import fill_index_random, random_unit_queries
    n, d, k = 1_000_000, 512, 10
    idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
    fill_index_random(idx, n)
    xq = np.concatenate([random_unit_queries(16, d), c2df_style_queries(idx, [3, 77_777, 500_000, 999_999])])
    xq16 = f16(xq)
    Do, Io = oracle_topk_chunked(idx, xq16, k + 16, chunk_rows=250_000)
    # batch 1 (K3, one launch per query) — the configuration C2 is quoted on
    for i in range(xq.shape[0]):
        D, I = idx.search(xq[i:i + 1], k)
        assert idx.stat("last_kernel") == 0
        assert_parity(D, I, Do[i:i + 1], Io[i:i + 1], k)
    assert list(Io[16:, 0]) == [3, 77_777, 500_000, 999_999] and np.all(np.abs(Do[16:, 0] - 0.9987) < 2e-3)
    # the same queries as one batch (K4t) and as a padded 300-query batch (K4 pairs)
    D, I = idx.search(xq, k)
    assert_parity(D, I, Do, Io, k)
    big = np.concatenate([xq, random_unit_queries(280, d, seed=9)])
    D, I = idx.search(big, k)
    assert_parity(D[:20], I[:20], Do, Io, k)
    idx.close()


def test_c3_10m_x_768_batch_4096_k100_against_c_oracle():
    import torch
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.synth import fill_index_random, random_unit_queries
    from sgic_b200.verify import verify_search
    n, d, nq, k = 10_000_000, 768, 4096, 100
    idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
    fill_index_random(idx, n)
    xq = random_unit_queries(nq, d)
    D, I = idx.search(xq, k)
    assert idx.stat("last_kernel") == 4                 # CTA pairs, query tile streamed (d > 512)
    assert np.all(np.diff(D, axis=1) <= 0) and I.min() >= 0 and I.max() < n
    rng = np.random.default_rng(3)
    sel = np.sort(rng.choice(nq, 64, replace=False))
    sel[0], sel[-1] = 0, nq - 1
    # torch re-score of the sample over the HBM rows (what bench.py's verify block does) ...
    q_dev = torch.from_numpy(xq).cuda()
    rec, (ref_s, ref_i) = verify_search(idx, q_dev, D, I, k, sample=sel)
    assert rec["ok"] and rec["queries"] == 64, rec
    # ... and the C restatement of FAISS on 32 of them over all 10M rows (sgemm-blocked path, reservoir at k = 100)
    sub = sel[::2]
    Do, Io = oracle_topk_chunked(idx, f16(xq[sub]), k + 16, chunk_rows=1 << 20)
    assert_parity(D[sub], I[sub], Do, Io, k)
    # the two oracles agree with each other on the winners
    assert np.array_equal(ref_i.cpu().numpy()[::2, :k // 2], Io[:, :k // 2]) or \
        assert_parity(Do[:, :k].astype(np.float32), Io[:, :k], ref_s.cpu().numpy()[::2], ref_i.cpu().numpy()[::2], k)
    idx.close()


def test_c4_100m_x_512_every_batch_regime_rescored():
    """102.4 GB on one GPU: K3's 3.1M-tile schedule, the mid-batch kernels and the auto-dispatched K4b."""
    import torch
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.synth import fill_index_random, random_unit_queries
    from sgic_b200.verify import verify_search
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(0)
    if free < 125 << 30:
        pytest.skip(f"needs ~110 GB of free HBM, {free >> 30} GB available")
    n, d, k = 100_000_000, 512, 10
    idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
    fill_index_random(idx, n, chunk_rows=500_000)
    assert idx.ntotal == n
    rng = np.random.default_rng(17)
    planted = np.sort(rng.choice(n, 12, replace=False))
    planted[0], planted[-1] = 0, n - 1                   # first and last row of the shard
    xq_all = np.concatenate([c2df_style_queries(idx, planted), random_unit_queries(4096 - 12, d)])
    kernels = {}
    for nq in (1, 2, 128, 256, 1024, 4096):
        xq = xq_all[:nq]
        q_dev = torch.from_numpy(xq).cuda()
        D, I = idx.search_torch(q_dev, k)
        torch.cuda.synchronize()
        kernels[nq] = idx.stat("last_kernel")
        sel = np.unique(np.concatenate([np.arange(min(nq, 12)), rng.choice(nq, min(nq, 20), replace=False)]))
        rec, (ref_s, ref_i) = verify_search(idx, q_dev, D, I, k, sample=sel)
        assert rec["ok"], (nq, rec)
        Dh, Ih = D.cpu().numpy(), I.cpu().numpy()
        if nq >= 12:
            assert np.array_equal(Ih[:12, 0], planted) and np.all(np.abs(Dh[:12, 0] - 0.9987) < 2e-3)
        # host-buffer entry point: same answer
        D2, I2 = idx.search(xq[:min(nq, 64)], k)       # (a 64-query batch may run on another kernel: near-ties may swap)
        assert (I2 == Ih[:min(nq, 64)]).mean() > 0.995 and np.abs(D2 - Dh[:min(nq, 64)]).max() < 3e-5
        if nq == 4096:
            # the C oracle confirms the winners on the candidate superset of 16 queries
            for r, qi in enumerate(sel[:16]):
                cand = np.unique(np.concatenate([ref_i[r].cpu().numpy(), Ih[qi]]))
                xb = np.concatenate([idx.reconstruct_n(int(c), 1) for c in cand])
                Do, Io = flat_ip_search_c(xb, f16(xq[qi:qi + 1]), k + 8)
                assert_parity(Dh[qi:qi + 1], Ih[qi:qi + 1], Do, cand[Io], k)
    assert kernels[1] == 0, kernels                      # K3
    assert kernels[1024] == 5 and kernels[4096] == 5, kernels   # K4b, chosen by the dispatcher (>= 90 GB shard)
    idx.close()
