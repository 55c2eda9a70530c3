"""Seeded random shapes through every regime (K3, K4t, K4 single CTA / pairs, K4b, bounded passes): ragged n, any d
that is a multiple of 8, batches across the kernel boundaries (1, 2-3, <=128, 129-256, >256), k from 1 to beyond
1024, fp16 and bf16 — each answer checked against the fp64 oracle on the values the GPU stores (O-exact) and against
the fp32 oracle with the north-star tolerance (O-ref)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.flat_ip import check_topk


def _cases(n_cases=36, seed=2026):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_cases):
        n = int(rng.choice([1, 7, 255, 256, 257, 1000, 4097, 20000, 50001]))
        d = int(rng.choice([8, 64, 72, 128, 320, 512, 520, 768, 1024, 1280]))
        nq = int(rng.choice([1, 2, 3, 17, 128, 129, 256, 257, 600]))
        k = int(rng.choice([1, 2, 10, 16, 17, 32, 33, 100, 300, 1024, 1500]))
        mode = int(rng.choice([0, 0, 4, 3, 1]))
        dtype = "bf16" if rng.random() < 0.25 else "fp16"
        out.append((i, n, d, nq, k, mode, dtype))
    return out


@pytest.mark.parametrize("i,n,d,nq,k,mode,dtype", _cases())
def test_random_shape(i, n, d, nq, k, mode, dtype):
    import torch
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(1000 + i)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    if n >= 20:
        xb[n // 2: n // 2 + 5] = xb[:5]                      # a few exact duplicates
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    xq[0] = xb[0]
    idx = faiss.IndexFlatIP(d, dtype=dtype, device=0)
    idx.add(xb)
    idx.set_option("dense_mode", mode)
    D, I = idx.search(xq, k)
    idx.close()
    assert D.shape == (nq, k) and I.shape == (nq, k)
    sel = np.unique(np.concatenate([[0, nq - 1], rng.choice(nq, min(nq, 12), replace=False)]))
    if dtype == "fp16":
        r = lambda x: x.astype(np.float16).astype(np.float64)
        tol_ref = 1e-3
    else:
        r = lambda x: torch.from_numpy(x).to(torch.bfloat16).to(torch.float64).numpy()
        tol_ref = 6e-3                                         # bf16 keeps 8 bits of mantissa
    check_topk(D[sel], I[sel], r(xb), r(xq[sel]), k, score_tol=3e-5, tie_tol=1e-6)
    check_topk(D[sel], I[sel], xb, xq[sel], k, score_tol=tol_ref)
    kk = min(k, n)
    assert np.all(I[:, kk:] == -1) and np.all(I[:, :kk] >= 0)
    assert np.all(np.diff(D[:, :kk], axis=1) <= 0)
