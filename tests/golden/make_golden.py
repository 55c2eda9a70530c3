"""Generates tests/golden/c2df_golden.npz by RUNNING THE REFERENCE'S OWN CODE in the build
container (where /root/reference exists).  Re-run:  python tests/golden/make_golden.py

What is executed from the reference, unmodified:
  * src/filemaker.py        pack_c2df / unpack_c2df            (imports fine: numpy + torch)
  * src/search.py           l2n, dequantize_clip_u8, decode_clip_from_c2df, do_search
search.py has top-level imports of faiss / zstandard / open_clip / PIL which are not
installed; they are satisfied with stub modules so that the module body can execute:
  * `zstandard` → a 20-line shim over the system libzstd (same one-shot semantics);
  * `faiss`, `open_clip` → empty stubs (nothing generated here calls into them).
The quantiser (compress.py:77) cannot be imported (compress.py needs omegaconf, torchvision,
the codec model …); it is restated below in one line and pinned by KAT-1 against the shipped
apple.npy / apple.c2df pair.
"""
import ctypes as C
import io
import json
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent / "c2df_golden.npz"

# ---- stubs for the absent third-party modules ------------------------------------------------
_z = C.CDLL("libzstd.so.1")
_z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
_z.ZSTD_getFrameContentSize.argtypes = [C.c_char_p, C.c_size_t]
_z.ZSTD_decompress.restype = C.c_size_t
_z.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
_z.ZSTD_compress.restype = C.c_size_t
_z.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
_z.ZSTD_compressBound.restype = C.c_size_t
_z.ZSTD_compressBound.argtypes = [C.c_size_t]
_z.ZSTD_isError.restype = C.c_uint
_z.ZSTD_isError.argtypes = [C.c_size_t]


class ZstdError(Exception):
    pass


class ZstdDecompressor:
    def decompress(self, data):
        data = bytes(data)
        n = _z.ZSTD_getFrameContentSize(data, len(data))
        if n >= 2 ** 64 - 2:
            raise ZstdError("could not determine content size in frame header")
        buf = C.create_string_buffer(max(int(n), 1))
        r = _z.ZSTD_decompress(buf, n, data, len(data))
        if _z.ZSTD_isError(r):
            raise ZstdError("decompression error")
        return buf.raw[:r]


class ZstdCompressor:
    def __init__(self, level=3):
        self.level = level

    def compress(self, data):
        data = bytes(data)
        cap = _z.ZSTD_compressBound(len(data))
        buf = C.create_string_buffer(cap)
        r = _z.ZSTD_compress(buf, cap, data, len(data), self.level)
        return buf.raw[:r]


zmod = types.ModuleType("zstandard")
zmod.ZstdDecompressor, zmod.ZstdCompressor, zmod.ZstdError = ZstdDecompressor, ZstdCompressor, ZstdError
sys.modules["zstandard"] = zmod
fa = types.ModuleType("faiss")
fa.Index = object
sys.modules["faiss"] = fa
sys.modules["open_clip"] = types.ModuleType("open_clip")
sys.path.insert(0, str(REF))
import filemaker as ref_fm   # noqa: E402  the reference's container code
import search as ref_search  # noqa: E402  the reference's query glue


def quantize(z):  # src/compress.py:77 (restated; pinned by KAT-1)
    return np.clip(np.round((z * 0.5 + 0.5) * 255.0), 0, 255).astype(np.uint8)


def blob_for(vec, rng, *, dim_override=None, extra=None, drop=(), meta_override=None, stream_override=None,
             big=False):
    q = quantize(vec)
    stream = ZstdCompressor(level=19).compress(q.tobytes()) if stream_override is None else stream_override
    meta = {"model_id": "ViT-B-32:laion2b_s34b_b79k", "dim": int(vec.shape[0] if dim_override is None else dim_override),
            "quant": "u8_symmetric_-1_1", "codec": "zstd", "zstd_level": 19}
    if meta_override is not None:
        meta = meta_override
    enc = {}
    if big:  # the full reference layout: image bitstreams and shape entries in front of the clip entries
        enc["z_bit_stream"] = rng.integers(0, 256, 769, dtype=np.uint8).tobytes()
        enc["h_bit_stream"] = rng.integers(0, 256, 807, dtype=np.uint8).tobytes()
        enc["img_shape"] = [1000, 859]
        enc["feat_shape"] = np.array([1, 12, 16, 16])
        enc["stack_shape"] = (4, 4)
        enc["token_length"] = 512
        enc["z_indices_shape"] = [1, 1, 32, 16]
    enc["clip_stream"] = stream
    enc["clip_meta"] = meta
    if extra:
        enc.update(extra)
    for k in drop:
        enc.pop(k, None)
    header = {"version": 2, "model_id": "ViT-B-32:laion2b_s34b_b79k", "embed_dim": int(vec.shape[0]),
              "quant_type": "u8_symmetric_-1_1", "image_hw": [1000, 859], "padding": [0, 165, 0, 24]}
    return ref_fm.pack_c2df(enc, header)


def main():
    rng = np.random.default_rng(20261018)
    out = {}

    # (1) pack/unpack cases over every entry type (filemaker.py:20-73)
    import torch
    type_case = {
        "a_bytes": b"\x00\x01\xfe\xff", "b_str": "héllo ✓", "c_int": -12345678901, "d_float": 3.141592653589793,
        "e_json": {"k": [1, 2.5, "x", None, True]}, "f_list": [1, 2, 3], "g_np": np.arange(12, dtype=np.float32).reshape(3, 4),
        "h_none": None, "i_bool_t": True, "j_bool_f": False, "k_shape": [7, 8, 9], "l_length": 77.0,
        "token_length": 31, "m_tensor": torch.arange(6, dtype=torch.int16).reshape(2, 3), "n_empty": b"",
        "o_u8": np.array([1, 2, 3], dtype=np.uint8),
    }
    hdr = {"version": 7, "note": "üñí", "nested": {"a": [1, 2]}}
    out["types_blob"] = np.frombuffer(ref_fm.pack_c2df(type_case, hdr), dtype=np.uint8)
    enc, h = ref_fm.unpack_c2df(bytes(out["types_blob"]))
    out["types_json"] = np.frombuffer(json.dumps(
        {k: (v.tolist() if isinstance(v, np.ndarray) else (v.hex() if isinstance(v, bytes) else v)) for k, v in enc.items()},
        ensure_ascii=False).encode(), dtype=np.uint8)
    out["types_dtypes"] = np.frombuffer(json.dumps(
        {k: [str(v.dtype), list(v.shape)] for k, v in enc.items() if isinstance(v, np.ndarray)}).encode(), dtype=np.uint8)
    out["types_header"] = np.frombuffer(json.dumps(h, ensure_ascii=False).encode(), dtype=np.uint8)

    # (2) good files: reference decode_clip_from_c2df (search.py:24-41) on blobs made by the reference packer
    dims = [512] * 10 + [768] * 4 + [64, 8]
    blobs, vecs, codes = [], [], []
    for i, d in enumerate(dims):
        v = rng.standard_normal(d).astype(np.float32)
        if i % 3 == 0:
            v += 2.0 * np.ones(d, dtype=np.float32) / np.sqrt(d)  # "cone": correlated coordinates
        v /= np.linalg.norm(v)
        b = blob_for(v, rng, big=(i % 2 == 0))
        z, header = ref_search.decode_clip_from_c2df(io.BytesIO(b).getvalue())
        blobs.append(b)
        vecs.append(z)
        codes.append(quantize(v))
    out["good_blob"] = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    out["good_offsets"] = np.cumsum([0] + [len(b) for b in blobs]).astype(np.int64)
    out["good_dims"] = np.array(dims, dtype=np.int32)
    out["good_vecs"] = np.concatenate([v.ravel() for v in vecs]).astype(np.float32)
    out["good_codes"] = np.concatenate(codes).astype(np.uint8)

    # (3) malformed files and the exception class the reference raises for each
    v = rng.standard_normal(512).astype(np.float32)
    v /= np.linalg.norm(v)
    good = blob_for(v, rng, big=True)
    raw_frame = ZstdCompressor(level=19).compress(quantize(v).tobytes())
    bad = {
        "bad_magic": b"XXXX" + good[4:],
        "empty": b"",
        "truncated_header": good[:20],
        "truncated_mid_entry": good[:1000],
        "no_clip_stream": blob_for(v, rng, drop=("clip_stream",)),
        "no_clip_meta": blob_for(v, rng, drop=("clip_meta",)),
        "dim_zero": blob_for(v, rng, dim_override=0),
        "dim_negative": blob_for(v, rng, dim_override=-5),
        "dim_missing": blob_for(v, rng, meta_override={"model_id": "x"}),
        "meta_none": blob_for(v, rng, extra={"clip_meta": None}),
        "dim_mismatch": blob_for(v, rng, dim_override=256),
        "zstd_garbage": blob_for(v, rng, stream_override=b"\x28\xb5\x2f\xfd" + bytes(40)),
        "zstd_not_a_frame": blob_for(v, rng, stream_override=b"hello world, not zstd"),
        "zstd_truncated": blob_for(v, rng, stream_override=raw_frame[:-7]),
        "stream_is_str": blob_for(v, rng, extra={"clip_stream": "abc"}),
        "dim_as_string": blob_for(v, rng, meta_override={"dim": "512"}),
        "dim_as_float": blob_for(v, rng, meta_override={"dim": 512.0}),
    }
    names, classes, bblobs = [], [], []
    for name, b in bad.items():
        try:
            ref_search.decode_clip_from_c2df(b)
            cls = "OK"
        except BaseException as e:  # noqa: BLE001
            cls = type(e).__name__
        names.append(name)
        classes.append(cls)
        bblobs.append(b)
    out["bad_blob"] = np.frombuffer(b"".join(bblobs), dtype=np.uint8)
    out["bad_offsets"] = np.cumsum([0] + [len(b) for b in bblobs]).astype(np.int64)
    out["bad_names"] = np.array(names)
    out["bad_classes"] = np.array(classes)

    # (4) the shipped fixture through the reference's own decode (KAT-4 input)
    z, _ = ref_search.decode_clip_from_c2df(Path("/root/reference/IO/bitstreams/apple.c2df"))
    out["apple_vec_from_c2df"] = z.astype(np.float32)

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape, str(v.dtype)) for k, v in out.items()})
    print(dict(zip(names, classes)))


if __name__ == "__main__":
    main()
