"""Host-side mirrors of the reference's glue (c2df container, codec helpers, zstd binding)
against vectors produced by the reference's own code (tests/golden/make_golden.py)."""
import json

import numpy as np
import pytest

from sgic_b200 import c2df, zstd


@pytest.fixture(scope="module")
def g(golden):
    return np.load(golden / "c2df_golden.npz")


def _retrieval():
    # retrieval imports faiss_compat which only binds the library lazily: importable without a GPU
    from sgic_b200 import retrieval
    return retrieval


def test_unpack_every_entry_type_matches_reference(g):
    enc, header = c2df.unpack_c2df(g["types_blob"].tobytes())
    want = json.loads(g["types_json"].tobytes().decode())
    dtypes = json.loads(g["types_dtypes"].tobytes().decode())
    assert list(enc) == list(want)                     # insertion order preserved
    for k, v in enc.items():
        if isinstance(v, np.ndarray):
            assert [str(v.dtype), list(v.shape)] == dtypes[k]
            assert v.tolist() == want[k]
        elif isinstance(v, bytes):
            assert v.hex() == want[k]
        else:
            assert v == want[k] and type(v) is type(want[k])
    assert header == json.loads(g["types_header"].tobytes().decode())


def test_pack_is_byte_identical_to_reference(g):
    import torch
    type_case = {
        "a_bytes": b"\x00\x01\xfe\xff", "b_str": "héllo ✓", "c_int": -12345678901, "d_float": 3.141592653589793,
        "e_json": {"k": [1, 2.5, "x", None, True]}, "f_list": [1, 2, 3], "g_np": np.arange(12, dtype=np.float32).reshape(3, 4),
        "h_none": None, "i_bool_t": True, "j_bool_f": False, "k_shape": [7, 8, 9], "l_length": 77.0,
        "token_length": 31, "m_tensor": torch.arange(6, dtype=torch.int16).reshape(2, 3), "n_empty": b"",
        "o_u8": np.array([1, 2, 3], dtype=np.uint8),
    }
    hdr = {"version": 7, "note": "üñí", "nested": {"a": [1, 2]}}
    assert c2df.pack_c2df(type_case, hdr) == g["types_blob"].tobytes()


def test_shipped_fixture_layout(golden):
    raw = (golden / "apple.c2df").read_bytes()
    enc, header = c2df.unpack_c2df(golden / "apple.c2df")
    assert header == {"version": 2, "model_id": "ViT-B-32:laion2b_s34b_b79k", "embed_dim": 512,
                      "quant_type": "u8_symmetric_-1_1", "image_hw": [1000, 859], "padding": [0, 165, 0, 24]}
    assert list(enc) == ["z_bit_stream", "h_bit_stream", "img_shape", "feat_shape", "stack_shape", "token_length",
                         "z_indices_shape", "clip_stream", "clip_meta"]
    assert len(enc["clip_stream"]) == 331 and raw[2016:2016 + 331] == enc["clip_stream"]
    assert enc["clip_meta"]["dim"] == 512 and enc["token_length"] == 512
    # round trip through our writer reproduces the file bit for bit
    assert c2df.pack_c2df(enc, header) == raw


def test_decode_clip_from_c2df_bit_exact_vs_reference(g, golden):
    r = _retrieval()
    offs, dims = g["good_offsets"], g["good_dims"]
    blob = g["good_blob"].tobytes()
    pos = 0
    for i, d in enumerate(dims):
        z, header = r.decode_clip_from_c2df(blob[offs[i]:offs[i + 1]])
        assert z.dtype == np.float32 and z.shape == (d,)
        assert np.array_equal(z, g["good_vecs"][pos:pos + d])
        assert header["model_id"] == "ViT-B-32:laion2b_s34b_b79k"
        pos += d
    z, _ = r.decode_clip_from_c2df(golden / "apple.c2df")
    assert np.array_equal(z, g["apple_vec_from_c2df"])
    q = r.encode_c2df_query(golden / "apple.c2df")
    assert q.shape == (1, 512) and q.dtype == np.float32


def test_decode_errors_have_the_reference_exception_classes(g):
    r = _retrieval()
    offs = g["bad_offsets"]
    blob = g["bad_blob"].tobytes()
    for i, (name, cls) in enumerate(zip(g["bad_names"], g["bad_classes"])):
        b = blob[offs[i]:offs[i + 1]]
        if cls == "OK":
            z, _ = r.decode_clip_from_c2df(b)
            assert z.shape == (512,)
            continue
        with pytest.raises(Exception) as ei:
            r.decode_clip_from_c2df(b)
        got = type(ei.value).__name__
        same = {"error": {"error", "IndexError"}}.get(str(cls), {str(cls)})   # struct.error at EOF
        assert got in same, f"{name}: reference raises {cls}, got {got}"


def test_quantiser_and_zstd_round_trip(golden):
    from sgic_b200 import index_build
    r = _retrieval()
    npy = np.load(golden / "apple.npy")
    stream, meta = index_build.quantize_u8_and_compress(npy)
    assert meta == {"model_id": "ViT-B-32:laion2b_s34b_b79k", "dim": 512, "quant": "u8_symmetric_-1_1",
                    "codec": "zstd", "zstd_level": 19}
    enc, _ = c2df.unpack_c2df(golden / "apple.c2df")
    assert zstd.decompress(stream) == zstd.decompress(enc["clip_stream"])      # KAT-1 through our code
    with pytest.raises(zstd.ZstdError):
        zstd.decompress(b"not a frame at all")
    with pytest.raises(TypeError):
        zstd.decompress("abc")
    z = r.dequantize_clip_u8(np.frombuffer(zstd.decompress(stream), dtype=np.uint8))
    assert abs(float(np.linalg.norm(z)) - 1.0) < 1e-6
    assert np.array_equal(r.l2n(np.zeros((1, 4), np.float32)), np.zeros((1, 4), np.float32))   # eps guard
