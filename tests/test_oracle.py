"""The oracle itself: numpy restatement vs fp64 brute force, C restatement vs numpy, and the
known-answer tests the shipped fixtures give."""
import numpy as np
import pytest

from oracle import c2df_ref
from oracle.flat_ip import NEG_FLT_MAX, check_topk, flat_ip_search
from oracle.flat_ip_c import flat_ip_search_c


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def brute(xb, xq, k):
    s = xq.astype(np.float64) @ xb.astype(np.float64).T
    n = xb.shape[0]
    D = np.full((xq.shape[0], k), NEG_FLT_MAX, np.float32)
    I = np.full((xq.shape[0], k), -1, np.int64)
    for i in range(xq.shape[0]):
        order = np.lexsort((np.arange(n), -s[i]))[:k]
        D[i, :order.size] = s[i, order]
        I[i, :order.size] = order
    return D, I


@pytest.mark.parametrize("n,d,nq,k", [(1, 8, 1, 1), (5, 16, 3, 10), (1000, 64, 1, 10), (3000, 128, 19, 10),
                                      (3000, 128, 20, 10), (5000, 64, 33, 100), (2500, 32, 7, 1), (300, 512, 256, 100)])
def test_numpy_oracle_vs_fp64_bruteforce(n, d, nq, k):
    rng = np.random.default_rng(n + d + nq + k)
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    D, I = flat_ip_search(xb, xq, k)
    check_topk(D, I, xb, xq, k, score_tol=2e-6, tie_tol=2e-6)
    D64, I64 = flat_ip_search(xb, xq, k, dtype=np.float64)
    Db, Ib = brute(xb, xq, k)
    assert np.array_equal(I64, Ib)
    assert np.allclose(D64, Db, atol=1e-6)


@pytest.mark.parametrize("n,d,nq,k", [(1, 8, 1, 1), (7, 16, 2, 10), (2000, 64, 1, 10), (2000, 128, 19, 10),
                                      (5000, 128, 20, 10), (5000, 64, 40, 100), (5000, 64, 3, 100), (9000, 512, 64, 10)])
def test_c_oracle_matches_numpy_oracle(n, d, nq, k):
    """Both FAISS code paths (seq for nq<20, blocked for nq>=20; heap for k<100, reservoir above)."""
    rng = np.random.default_rng(1000 + n + d + nq + k)
    xb, xq = unit(rng, n, d), unit(rng, nq, d)
    Dc, Ic = flat_ip_search_c(xb, xq, k)
    check_topk(Dc, Ic, xb, xq, k, score_tol=2e-6, tie_tol=2e-6)
    Dn, In = flat_ip_search(xb, xq, k)
    assert np.allclose(Dc, Dn, atol=2e-6)
    assert (Ic == In).mean() > 0.999     # order may differ only inside near-ties


def test_c_oracle_rowpar_variant_agrees():
    rng = np.random.default_rng(5)
    xb, xq = unit(rng, 20000, 64), unit(rng, 2, 64)
    D1, I1 = flat_ip_search_c(xb, xq, 10)
    D2, I2 = flat_ip_search_c(xb, xq, 10, rowpar=True)
    assert np.array_equal(I1, I2) and np.allclose(D1, D2, atol=1e-6)


def test_duplicates_earlier_row_wins_and_padding():
    rng = np.random.default_rng(2)
    base = unit(rng, 50, 32)
    xb = np.concatenate([base, base])
    for fn in (flat_ip_search, flat_ip_search_c):
        D, I = fn(xb, base[:4], 2)
        assert np.array_equal(I[:, 0], np.arange(4)) and np.array_equal(I[:, 1], np.arange(4) + 50)
        D, I = fn(xb[:3], base[:1], 5)
        assert np.all(I[0, 3:] == -1) and np.all(D[0, 3:] == NEG_FLT_MAX)
    with pytest.raises(AssertionError):
        flat_ip_search(xb, base[:1], 0)
    with pytest.raises(RuntimeError):
        flat_ip_search_c(xb, base[:1], 0)


# ---------------------------------------------------------------- KATs on the shipped fixtures
def test_kat1_quantiser_matches_shipped_stream(golden):
    npy = np.load(golden / "apple.npy")
    q, z = c2df_ref.decode_clip((golden / "apple.c2df").read_bytes())
    assert q.shape == (512,) and int(q.sum()) == 65377 and list(q[:8]) == [133, 131, 119, 129, 125, 125, 132, 119]
    assert np.array_equal(c2df_ref.quantize_u8(npy), q)


def test_kat2_index_payload_is_renormalised_npy(golden):
    npy = np.load(golden / "apple.npy")
    x = c2df_ref.read_ixfi(golden / "index.faiss")
    assert x.shape == (1, 512)
    v = npy.copy()[None, :]
    v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12     # FaissDB.add, compress.py:105-107
    assert np.array_equal(x, v.astype("float32"))
    assert (golden / "ids.txt").read_text() == "../IO/bitstreams/apple.c2df\n"


def test_kat3_kat4_scores(golden):
    xb = c2df_ref.read_ixfi(golden / "index.faiss")
    npy = np.load(golden / "apple.npy")
    D, I = flat_ip_search(xb, npy[None, :], 1)
    assert I[0, 0] == 0 and abs(float(D[0, 0]) - 1.0) < 1e-6
    _, z = c2df_ref.decode_clip((golden / "apple.c2df").read_bytes())
    D, I = flat_ip_search(xb, z[None, :], 1)
    assert I[0, 0] == 0 and abs(float(D[0, 0]) - 0.9987136) < 2e-6
    Dc, Ic = flat_ip_search_c(xb, z[None, :], 1)
    assert Ic[0, 0] == 0 and abs(float(Dc[0, 0]) - 0.9987136) < 2e-6


def test_kat5_fp16_bf16_rounding_budget(golden):
    import torch
    xb = c2df_ref.read_ixfi(golden / "index.faiss")
    npy = np.load(golden / "apple.npy")
    _, z = c2df_ref.decode_clip((golden / "apple.c2df").read_bytes())
    for q, exact in ((npy, 1.0), (z, 0.9987136)):
        h = float(xb.astype(np.float16).astype(np.float64)[0] @ q.astype(np.float16).astype(np.float64))
        b = float(torch.from_numpy(xb).bfloat16().double().numpy()[0] @ torch.from_numpy(q).bfloat16().double().numpy())
        assert abs(h - exact) < 1e-3 and abs(b - exact) < 1e-3   # north_star budget holds for both


def test_oracle_decode_matches_reference_generated_vectors(golden):
    g = np.load(golden / "c2df_golden.npz")
    offs, dims = g["good_offsets"], g["good_dims"]
    blob = g["good_blob"].tobytes()
    pos = 0
    for i, d in enumerate(dims):
        q, z = c2df_ref.decode_clip(blob[offs[i]:offs[i + 1]])
        assert np.array_equal(q, g["good_codes"][pos:pos + d])
        assert np.array_equal(z, g["good_vecs"][pos:pos + d])        # bit-exact vs the reference's own code
        pos += d
    _, z = c2df_ref.decode_clip((golden / "apple.c2df").read_bytes())
    assert np.array_equal(z, g["apple_vec_from_c2df"])


def test_codes_to_f32_is_bit_identical_to_the_reference_arithmetic(golden):
    """sgic_codes_to_f32 (what write_index uses for an index that retains its u8 codes) against the reference's
    own expression run by numpy, row by row as build.py:82 does: identical bits, every width."""
    from oracle import c2df_ref
    from sgic_b200 import faiss_compat
    rng = np.random.default_rng(11)
    for d in (8, 16, 64, 104, 128, 136, 256, 512, 520, 768, 1024, 1280, 2048):
        q = rng.integers(0, 256, (300, d), dtype=np.uint8)
        q[0] = 0
        q[1] = 255
        q[2] = 128                                   # z ~ 0.0039: tiny norm, eps guard not hit but close to it
        q[3] = rng.integers(120, 136, d)             # CLIP-like narrow codes
        got = faiss_compat.codes_to_f32(q)
        want = np.stack([c2df_ref.dequantize_clip_u8(r) for r in q])
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32)), d
    # the shipped pair: codes of apple.c2df -> the vector search.py derives from it (golden, reference-run)
    g = np.load(golden / "c2df_golden.npz")
    _, z = c2df_ref.decode_clip((golden / "apple.c2df").read_bytes())
    codes = np.clip(np.round((np.load(golden / "apple.npy") * 0.5 + 0.5) * 255.0), 0, 255).astype(np.uint8)
    assert np.array_equal(faiss_compat.codes_to_f32(codes[None, :])[0], g["apple_vec_from_c2df"])
    assert np.array_equal(z, g["apple_vec_from_c2df"])
