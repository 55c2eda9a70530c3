"""GPU parity tests of the search path (K3 streaming kernel + merge) against the CPU oracle.
Everything goes through the faiss-compatible surface, i.e. through the C ABI."""
import numpy as np
import pytest
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

pytestmark = pytest.mark.gpu

from oracle.flat_ip import check_topk, flat_ip_search, NEG_FLT_MAX


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def rounded(x, dtype):
    if dtype == "fp16":
        return x.astype(np.float16).astype(np.float64)
    import torch
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float64).numpy()


def make_index(xb, dtype="fp16"):
    from sgic_b200 import faiss_compat as faiss
    idx = faiss.IndexFlatIP(xb.shape[1], dtype=dtype, device=0)
    idx.add(xb)
    return idx


@pytest.mark.parametrize("n,d,nq,k", [
    (1, 512, 1, 1),
    (7, 512, 1, 10),        # k > ntotal: -1 / -FLT_MAX padding
    (33, 512, 2, 10),
    (1000, 512, 1, 10),
    (4097, 512, 3, 10),     # N not a multiple of the tile
    (20000, 512, 1, 100),
    (20000, 768, 4, 10),
    (5000, 256, 1, 10),
    (5000, 1024, 2, 10),
    (3000, 2048, 1, 10),
    (9999, 64, 1, 5),
    (50000, 512, 7, 10),    # > 4 queries: several passes
    (20000, 512, 1, 1024),
])
@pytest.mark.parametrize("force_k3", [True, False])
def test_search_matches_oracle(n, d, nq, k, force_k3):
    """force_k3: every batch size through the CUDA-core streaming kernel (its 2- and 4-query variants); otherwise
    the library's own regime choice (batches of 3+ go to the tensor-core kernels)."""
    if nq == 1 and not force_k3:
        pytest.skip("one query is always K3")
    rng = np.random.default_rng(n * 31 + d + nq + k)
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    idx = make_index(xb)
    if force_k3:
        idx.set_option("dense_min_nq", 1 << 30)
    assert idx.ntotal == n and idx.d == d
    D, I = idx.search(xq, k)
    # O-exact: oracle on the values the GPU holds (fp16-rounded db and queries)
    check_topk(D, I, rounded(xb, "fp16"), rounded(xq, "fp16"), k, score_tol=2e-5, tie_tol=1e-6)
    # O-ref: the fp32 vectors FAISS would hold; north_star tolerance 1e-3
    check_topk(D, I, xb, xq, k, score_tol=1e-3)


def test_bf16_storage():
    rng = np.random.default_rng(5)
    xb, xq = unit(rng, 30000, 512), unit(rng, 2, 512)
    idx = make_index(xb, "bf16")
    D, I = idx.search(xq, 10)
    check_topk(D, I, rounded(xb, "bf16"), rounded(xq, "bf16"), 10, score_tol=2e-5, tie_tol=1e-6)
    check_topk(D, I, xb, xq, 10, score_tol=4e-3)  # bf16 worst case 2^-8


def test_exact_duplicates_resolve_to_lowest_ids():
    """FaissDB re-adds every .npy on each run (src/compress.py:296-306), so exact ties are real."""
    rng = np.random.default_rng(11)
    base = unit(rng, 500, 512)
    xb = np.concatenate([base, base, base])           # every row three times
    xq = base[:3].copy()
    idx = make_index(xb)
    D, I = idx.search(xq, 6)
    for r in range(3):
        assert list(I[r, :3]) == [r, r + 500, r + 1000]   # (score desc, id asc)
        assert D[r, 0] == D[r, 1] == D[r, 2]
    Dref, Iref = flat_ip_search(rounded(xb, "fp16"), rounded(xq, "fp16"), 6, dtype=np.float64)
    assert np.array_equal(I, Iref)


def test_incremental_add_and_empty_index():
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(3)
    idx = faiss.IndexFlatIP(512, device=0)
    q = unit(rng, 1, 512)
    D, I = idx.search(q, 4)                           # ntotal == 0
    assert np.all(I == -1) and np.all(D == NEG_FLT_MAX)
    xb = unit(rng, 3000, 512)
    for s in range(0, 3000, 700):                     # grows the HBM allocation several times
        idx.add(xb[s:s + 700])
    for i in range(5):                                # one-row adds, as FaissDB.add does
        idx.add(xb[i][None, :])
    assert idx.ntotal == 3005
    full = np.concatenate([xb, xb[:5]])
    D, I = idx.search(q, 10)
    check_topk(D, I, rounded(full, "fp16"), rounded(q, "fp16"), 10, score_tol=2e-5, tie_tol=1e-6)


def test_argument_errors():
    from sgic_b200 import faiss_compat as faiss
    idx = faiss.IndexFlatIP(512, device=0)
    with pytest.raises(AssertionError):
        idx.add(np.zeros((2, 511), dtype=np.float32))
    with pytest.raises(AssertionError):
        idx.search(np.zeros((1, 512), dtype=np.float32), 0)
    with pytest.raises(RuntimeError):
        faiss.IndexFlatIP(513, device=0)
    with pytest.raises(RuntimeError):
        faiss.read_index("/nonexistent/index.faiss")


@pytest.mark.parametrize("n,nq,k", [(5000, 3, 2500), (5000, 2, 5000), (3000, 1, 7000), (40000, 2, 1025), (2049, 1, 2048)])
def test_k_beyond_1024_runs_in_passes(n, nq, k):
    """FAISS takes any k (do_search asks for min(topk, ntotal), src/search.py:114).  Beyond 1024 the answer is
    enumerated in passes of <= 1024, each bounded by the last key of the previous one: the concatenation must be
    the exact ordered top-k — including runs of exact ties that straddle a pass boundary — padded when k > n."""
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(n + k)
    base = unit(rng, n // 2, 512)
    xb = np.concatenate([base, base, unit(rng, n - 2 * (n // 2), 512)])        # every score appears twice: ties everywhere
    xq = unit(rng, nq, 512)
    idx = faiss.IndexFlatIP(512, device=0)
    idx.add(xb)
    D, I = idx.search(xq, k)
    assert D.shape == (nq, k) and I.shape == (nq, k)
    kk = min(k, n)
    assert np.all(I[:, kk:] == -1) and np.all(D[:, kk:] == np.float32(-3.4028234663852886e38))
    for r in range(nq):
        ids = I[r, :kk]
        assert len(set(ids.tolist())) == kk and ids.min() >= 0 and ids.max() < n     # no repeats, no gaps
        assert np.all(np.diff(D[r, :kk]) <= 0)
        pairs = list(zip((-D[r, :kk]).tolist(), ids.tolist()))
        assert pairs == sorted(pairs)                                              # (score desc, id asc) throughout
    check_topk(D, I, rounded(xb, "fp16"), rounded(xq, "fp16"), k, score_tol=3e-5, tie_tol=1e-6)
    # the first 1024 are what a k = 1024 search returns (a batch may take the tensor-core kernel there: same ids,
    # scores equal up to the accumulation order)
    D1, I1 = idx.search(xq, 1024)
    assert np.array_equal(I[:, :1024], I1)
    np.testing.assert_allclose(D[:, :1024], D1, atol=2e-5, rtol=0)


def test_database_grows_in_place_without_a_second_copy():
    """The database sits in one reserved virtual range that physical chunks are mapped into (cuMemMap): growing keeps
    the base address, copies nothing and never needs old + new buffers side by side — appends work past half of HBM
    without a reserve()."""
    import ctypes as C
    from sgic_b200 import _native, faiss_compat as faiss
    rng = np.random.default_rng(77)
    d, step, rounds = 512, 60_000, 10
    xb = rng.standard_normal((step * rounds, d)).astype(np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    idx = faiss.IndexFlatIP(d, device=0, retain_fp32=False)
    lib = _native.lib()
    ptrs, caps = [], []
    for r in range(rounds):
        idx.add(xb[r * step:(r + 1) * step])
        ptrs.append(lib.sgic_index_data_dev(idx._h))
        caps.append(idx.stat("capacity"))
    assert len(set(ptrs)) == 1, "the database moved while growing"
    assert caps == sorted(caps) and caps[-1] >= step * rounds and len(set(caps)) >= 2      # grew in several steps
    got = idx.reconstruct_n(0, step * rounds)
    assert np.array_equal(got, xb.astype(np.float16).astype(np.float32))                   # nothing lost on the way
    idx.reserve(caps[-1] + 1_000_000)
    assert lib.sgic_index_data_dev(idx._h) == ptrs[0] and idx.stat("capacity") >= caps[-1] + 1_000_000
    D, I = idx.search(xb[-1:], 1)
    assert I[0, 0] == step * rounds - 1
    idx.reset()
    assert idx.ntotal == 0 and lib.sgic_index_data_dev(idx._h) == ptrs[0]
    idx.close()
    # the cudaMalloc scheme (SGIC_VMM=0) stays available: same answers in a fresh process
    import os, subprocess, sys
    code = ("import numpy as np, sys; sys.path.insert(0, %r); from sgic_b200 import faiss_compat as faiss;"
            "rng=np.random.default_rng(1); x=rng.standard_normal((5000,64)).astype('float32');"
            "i=faiss.IndexFlatIP(64, device=0);"
            "[i.add(x[j*500:(j+1)*500]) for j in range(10)];"
            "D,I=i.search(x[4321:4322],1); assert I[0,0]==4321; print('ok')") % str(ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SGIC_VMM="0"), capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
