"""GPU tests of the ingest side (K1 dequant+normalise, K2 fp32 pack, batched .c2df ingest) and
of the file formats / glue either side of the search call, all through the C ABI."""
import io
import json
import contextlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import c2df_ref
from oracle.flat_ip import check_topk


def fp16_ulp_diff(a, b):
    """Distance in fp16 ulps between two fp32 arrays holding fp16-representable values."""
    ia = a.astype(np.float16).view(np.int16).astype(np.int32)
    ib = b.astype(np.float16).view(np.int16).astype(np.int32)
    ia = np.where(ia < 0, -32768 - ia, ia)
    ib = np.where(ib < 0, -32768 - ib, ib)
    return np.abs(ia - ib)


def test_add_f32_rounds_to_nearest_even_fp16_and_bf16():
    import torch
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 512)).astype(np.float32)
    idx = faiss.IndexFlatIP(512, device=0)
    idx.add(x)
    assert np.array_equal(idx.reconstruct_n(0, 1000), x.astype(np.float16).astype(np.float32))   # bit-exact
    assert np.array_equal(idx.reconstruct(17), x[17].astype(np.float16).astype(np.float32))
    idb = faiss.IndexFlatIP(512, device=0, dtype="bf16")
    idb.add(x)
    assert np.array_equal(idb.reconstruct_n(0, 1000), torch.from_numpy(x).bfloat16().float().numpy())


def test_add_u8_matches_dequantize_clip_u8():
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(1)
    for d in (512, 768, 64):
        q = rng.integers(0, 256, size=(4000, d), dtype=np.uint8)
        q[0] = 0; q[1] = 255; q[2] = 128                       # extreme rows
        idx = faiss.IndexFlatIP(d, device=0)
        idx.add_u8(q)
        got = idx.reconstruct_n(0, q.shape[0])
        want = c2df_ref.dequantize_clip_u8(q)                  # fp32, reference operation order
        ulp = fp16_ulp_diff(got, want)
        # fp32 sum-of-squares order differs from numpy's pairwise sum: allow 1 fp16 ulp on a few elements
        assert ulp.max() <= 1 and (ulp > 0).mean() < 2e-3, (ulp.max(), (ulp > 0).mean())
        assert np.abs(got - want).max() < 6e-4


def test_add_c2df_batch_and_skip_semantics(golden):
    from sgic_b200 import faiss_compat as faiss
    g = np.load(golden / "c2df_golden.npz")
    # good files of mixed dimension followed by every malformed file
    blob = np.concatenate([g["good_blob"], g["bad_blob"]])
    offs = np.concatenate([g["good_offsets"], g["bad_offsets"][1:] + g["good_offsets"][-1]])
    idx = faiss.IndexFlatIP(512, device=0)
    added, status = idx.add_c2df(blob, offs, n_threads=3)
    dims = g["good_dims"]
    n_good512 = int((dims == 512).sum())
    n_ok_bad = int((g["bad_classes"] == "OK").sum())          # dim given as "512" / 512.0 decode fine upstream
    assert added == n_good512 + n_ok_bad == idx.ntotal
    assert list(status[:len(dims)]) == [0 if d == 512 else 7 for d in dims]
    rows = idx.reconstruct_n(0, n_good512)
    pos = 0
    k = 0
    for d in dims:
        if d == 512:
            want = g["good_vecs"][pos:pos + d]
            assert fp16_ulp_diff(rows[k], want).max() <= 1
            k += 1
        pos += d


def test_write_index_is_bit_identical_to_the_shipped_faiss_file(golden, tmp_path):
    """KAT-2 through the product: FaissDB.add(npy) + persist == IO/faiss/{index.faiss, ids.txt}."""
    from sgic_b200.index_build import FaissDB
    db = FaissDB(str(tmp_path), 512, device=0)
    db.add(np.load(golden / "apple.npy"), "../IO/bitstreams/apple.c2df")
    db.persist()
    assert (tmp_path / "index.faiss").read_bytes() == (golden / "index.faiss").read_bytes()
    assert (tmp_path / "ids.txt").read_bytes() == (golden / "ids.txt").read_bytes()
    # resume: reopen, append, persist (compress.py:94-101) — no de-duplication, as upstream
    db2 = FaissDB(str(tmp_path), 512, device=0)
    assert db2.index.ntotal == 1 and db2.ids == ["../IO/bitstreams/apple.c2df"]
    db2.add(np.load(golden / "apple.npy"), "dup")
    db2.persist()
    x = c2df_ref.read_ixfi(tmp_path / "index.faiss")
    assert x.shape == (2, 512) and np.array_equal(x[0], x[1])


def test_kat3_kat4_kat6_through_load_index_and_do_search(golden, tmp_path):
    from sgic_b200 import retrieval
    for name in ("index.faiss", "ids.txt"):
        (tmp_path / name).write_bytes((golden / name).read_bytes())
    index, paths, meta = retrieval.load_index(tmp_path)
    assert index.ntotal == 1 and index.d == 512 and paths == ["../IO/bitstreams/apple.c2df"] and meta == {"dim": 512}
    res = retrieval.do_search(np.load(golden / "apple.npy")[None, :], index, paths, topk=10)   # KAT-6: k clamps to 1
    assert len(res) == 1 and res[0][0] == "../IO/bitstreams/apple.c2df" and abs(res[0][1] - 1.0) < 1e-3
    res = retrieval.do_search(retrieval.encode_c2df_query(golden / "apple.c2df"), index, paths, topk=10)
    assert res[0][0] == "../IO/bitstreams/apple.c2df" and abs(res[0][1] - 0.9987136) < 1e-3
    with pytest.raises(FileNotFoundError):
        retrieval.load_index(tmp_path / "nope")


def _write_corpus(dirpath, vecs, rng):
    from sgic_b200 import c2df as c2
    from sgic_b200.index_build import quantize_u8_and_compress
    (dirpath / "sub").mkdir(parents=True)
    names = []
    for i, v in enumerate(vecs):
        stream, meta = quantize_u8_and_compress(v)
        enc = {"z_bit_stream": rng.integers(0, 256, 100, dtype=np.uint8).tobytes(), "token_length": 512,
               "clip_stream": stream, "clip_meta": meta}
        header = {"version": 2, "model_id": "ViT-B-32:laion2b_s34b_b79k", "embed_dim": int(v.shape[0])}
        p = dirpath / ("sub" if i % 3 == 0 else "") / f"img_{i:04d}.c2df"
        p.write_bytes(c2.pack_c2df(enc, header))
        names.append(p)
    return names


def test_build_index_from_c2df_dir_and_cli(tmp_path, capsys):
    from sgic_b200 import index_build, retrieval
    rng = np.random.default_rng(7)
    vecs = rng.standard_normal((40, 512)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    src = tmp_path / "bitstreams"
    src.mkdir()
    files = _write_corpus(src, vecs, rng)
    (src / "broken.c2df").write_bytes(b"XXXXnot a container")
    (src / "noclip.c2df").write_bytes(__import__("sgic_b200").c2df.pack_c2df({"token_length": 1}, {"version": 2}))
    out = tmp_path / "faiss"
    index_build.build_index_from_c2df_dir(src, out)
    log = capsys.readouterr().out
    assert "[SKIP] broken.c2df" in log and "[SKIP] noclip.c2df" in log and "[OK] Index process completed!: N=40, dim=512" in log
    want_paths = [str(p) for p in sorted(src.glob("**/*.c2df")) if p.name not in ("broken.c2df", "noclip.c2df")]
    assert json.loads((out / "paths.json").read_text()) == want_paths
    assert (out / "ids.txt").read_text() == "\n".join(want_paths)                  # no trailing newline (build.py:100)
    assert json.loads((out / "meta.json").read_text()) == {"dim": 512, "model_id": "ViT-B-32:laion2b_s34b_b79k"}
    assert (out / "faiss.index").read_bytes() == (out / "index.faiss").read_bytes()
    # both naming schemes load; the preferred one carries meta.json
    index, paths, meta = retrieval.load_index(out)
    assert index.ntotal == 40 and meta["model_id"].startswith("ViT-B-32")
    # query with one of the bitstreams: it must come back first with score ~1
    target = want_paths[5]
    res = retrieval.do_search(retrieval.encode_c2df_query(target), index, paths, topk=5)
    assert res[0][0] == target and abs(res[0][1] - 1.0) < 1e-3 and len(res) == 5
    # row order == sorted path order: compare every stored row with the oracle decode
    rows = index.reconstruct_n(0, 40)
    ref_rows = []
    for i, p in enumerate(want_paths):
        _, z = c2df_ref.decode_clip(open(p, "rb").read())
        assert np.abs(rows[i] - z).max() < 6e-4
        ref_rows.append(z)
    # the IxFI files are BYTE-identical to what build.py:91-99 writes: fp32 rows of the reference decode (the u8
    # codes are retained on the host and re-expanded with numpy's operation order, not read back from fp16 HBM)
    c2df_ref.write_ixfi(tmp_path / "ref.index", np.stack(ref_rows).astype(np.float32))
    assert (out / "faiss.index").read_bytes() == (tmp_path / "ref.index").read_bytes()
    # CLI: stdout is exactly the JSON document webapp.py:249 parses
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        retrieval.main(["query-c2df", "--index_dir", str(out), "--c2df", target, "--topk", "3"])
    doc = json.loads(buf.getvalue())
    assert [set(e) for e in doc] == [{"path", "score"}] * 3 and doc[0]["path"] == target
    assert buf.getvalue() == json.dumps(doc, ensure_ascii=False, indent=2) + "\n"
    # legacy naming only
    legacy = tmp_path / "legacy"
    legacy.mkdir()
    for n in ("index.faiss", "ids.txt"):
        (legacy / n).write_bytes((out / n).read_bytes())
    index2, paths2, meta2 = retrieval.load_index(legacy)
    assert paths2 == want_paths and meta2 == {"dim": 512, "model_id": "ViT-B-32:laion2b_s34b_b79k"}


def test_build_ixfi_is_byte_identical_through_both_decode_routes(tmp_path):
    """A corpus large enough for several device-decode slabs plus files the device profile rejects (raw-stored
    frames go through libzstd on the host): codes are retained on either route and in file order."""
    from sgic_b200 import c2df as c2, faiss_compat as faiss, zstd
    from sgic_b200.index_build import quantize_u8_and_compress
    rng = np.random.default_rng(21)
    n, d = 70_000, 64                                  # > one slab of 65 536 files
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    blobs, codes = [], []
    for i, v in enumerate(vecs):
        stream, meta = quantize_u8_and_compress(v)
        if i % 1000 == 7:                               # level-1 frame of noise: a raw block, host route
            q = rng.integers(0, 256, d, dtype=np.uint8)
            stream = zstd.compress(q.tobytes(), 1)
        codes.append(np.frombuffer(zstd.decompress(stream), dtype=np.uint8))
        blobs.append(c2.pack_c2df({"clip_stream": stream, "clip_meta": meta}, {"version": 2}))
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(b) for b in blobs], out=offs[1:])
    index = faiss.IndexFlatIP(d, device=0, retain_fp32=False, retain_codes=True)
    added, status = index.add_c2df(np.frombuffer(b"".join(blobs), dtype=np.uint8), offs)
    assert added == n and not status.any()
    assert index.stat("retained_code_rows") == n
    codes = np.stack(codes)
    assert np.array_equal(index.codes(), codes)
    faiss.write_index(index, str(tmp_path / "a.index"))
    want = np.stack([c2df_ref.dequantize_clip_u8(q) for q in codes[:3000]])
    got = c2df_ref.read_ixfi(tmp_path / "a.index")
    assert got.shape == (n, d) and np.array_equal(got[:3000].view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got, faiss.codes_to_f32(codes))
    # an fp32 add afterwards invalidates the code copy: the writer falls back to the HBM rows
    index.add(vecs[:2])
    assert index.stat("retained_code_rows") == -1
    with pytest.raises(RuntimeError, match="u8 codes"):
        index.codes(0, 1)


def test_build_fails_as_a_whole_on_mixed_dimensions(tmp_path):
    """np.concatenate in build.py:91 raises when a decodable file has another dimension; nothing is written."""
    from sgic_b200 import index_build
    rng = np.random.default_rng(8)
    src = tmp_path / "bitstreams"
    src.mkdir()
    v = rng.standard_normal((4, 512)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    _write_corpus(src, v, rng)
    w = rng.standard_normal((1, 768)).astype(np.float32)
    w /= np.linalg.norm(w)
    _write_corpus(src / "other", w, rng)
    with pytest.raises(ValueError, match="must match exactly"):
        index_build.build_index_from_c2df_dir(src, tmp_path / "faiss")
    assert not (tmp_path / "faiss" / "faiss.index").exists()


def test_calls_on_different_streams_are_ordered_by_the_library(tmp_path):
    """add_torch on a side stream WITHOUT synchronising, then host-buffer search / reconstruct / save on the
    index's own stream, then a device search on yet another stream: every call must see the rows (include/sgic.h,
    "Streams")."""
    import torch
    from sgic_b200 import faiss_compat as faiss
    dev = torch.device("cuda", 0)
    d, n = 512, 600_000
    g = torch.Generator(device=dev).manual_seed(5)
    index = faiss.IndexFlatIP(d, device=0, retain_fp32=False, capacity=2 * n)
    side, side2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    xs = []
    for _ in range(2):
        x = torch.randn((n, d), device=dev, generator=g)
        x /= x.norm(dim=1, keepdim=True)
        xs.append(x)
    torch.cuda.synchronize(dev)
    for rep, x in enumerate(xs):
        with torch.cuda.stream(side):
            big = torch.randn((4096, 4096), device=dev)
            for _ in range(20):                       # keep the side stream busy so the add is still queued
                big = big @ big * 1e-3
            index.add_torch(x)
        q = x[n - 1:n].cpu().numpy()
        D, I = index.search(q, 1)                     # own stream, host buffers: no sync in between
        assert I[0, 0] == (rep + 1) * n - 1 and abs(D[0, 0] - 1.0) < 2e-3
        rows = index.reconstruct_n(index.ntotal - 1, 1)
        assert np.abs(rows[0] - q[0]).max() < 1e-3
    x = xs[1]
    faiss.write_shard(index, str(tmp_path / "s.sgi2"))
    back = faiss.read_index(str(tmp_path / "s.sgi2"), device=0)
    assert back.ntotal == 2 * n and np.abs(back.reconstruct_n(2 * n - 1, 1)[0] - x[n - 1].cpu().numpy()).max() < 1e-3
    back.close()
    with torch.cuda.stream(side):
        index.add_torch(x[:1000].clone())
    with torch.cuda.stream(side2):
        Dd, Id = index.search_torch(x[999:1000].contiguous(), 2)
    side2.synchronize()
    assert sorted(Id[0].tolist()) == [n + 999, 2 * n + 999]


def test_from_npy_dir_matches_sorted_glob_order(tmp_path):
    from sgic_b200 import index_build
    rng = np.random.default_rng(9)
    vecs = rng.standard_normal((25, 512)).astype(np.float32)
    (tmp_path / "clip_vecs").mkdir()
    (tmp_path / "bitstreams").mkdir()
    names = [f"im{(i * 7) % 25:03d}" for i in range(25)]
    for i, v in enumerate(vecs):
        np.save(tmp_path / "clip_vecs" / f"{names[i]}.npy", v)
        if names[i] not in ("im003", "im017"):          # compress.py:304: vectors without a bitstream are not indexed
            (tmp_path / "bitstreams" / f"{names[i]}.c2df").write_bytes(b"C2DF")
    bit = str(tmp_path / "bitstreams")
    db = index_build.from_npy_dir(tmp_path / "clip_vecs", bit, tmp_path / "faiss", device=0)
    assert db.index.ntotal == 23
    assert db.ids[0] == f"{bit}/im000.c2df" and db.ids == sorted(db.ids) and f"{bit}/im003.c2df" not in db.ids
    x = c2df_ref.read_ixfi(tmp_path / "faiss" / "index.faiss")
    order = [i for i in sorted(range(25), key=lambda i: names[i]) if names[i] not in ("im003", "im017")]
    want = vecs[order] / (np.linalg.norm(vecs[order], axis=1, keepdims=True) + 1e-12)
    assert np.array_equal(x, want.astype(np.float32))         # fp32 rows kept bit-exact on disk
    # on-disk v2 of clip_vecs: one packed (N, d) .npy + stems; the index built from it is byte-identical
    packed = index_build.pack_npy_dir(tmp_path / "clip_vecs")
    arr = np.load(packed)
    assert arr.shape == (25, 512) and arr.dtype == np.float32
    assert np.array_equal(arr, vecs[sorted(range(25), key=lambda i: names[i])])
    assert sorted(p.name for p in (tmp_path / "clip_vecs").glob("*.npy")) == sorted(n + ".npy" for n in names)  # untouched
    db2 = index_build.from_npy_dir(tmp_path / "clip_vecs", bit, tmp_path / "faiss2", device=0)
    assert db2.ids == db.ids
    assert (tmp_path / "faiss2" / "index.faiss").read_bytes() == (tmp_path / "faiss" / "index.faiss").read_bytes()
    for f in (tmp_path / "clip_vecs").glob("*.npy"):           # the packed copy alone is enough
        f.unlink()
    db3 = index_build.from_npy_dir(tmp_path / "clip_vecs", bit, tmp_path / "faiss3", device=0)
    assert (tmp_path / "faiss3" / "index.faiss").read_bytes() == (tmp_path / "faiss" / "index.faiss").read_bytes()
    # a stale packed copy (a vector file appeared since) is ignored
    np.save(tmp_path / "clip_vecs" / "zz_new.npy", vecs[0])
    (tmp_path / "bitstreams" / "zz_new.c2df").write_bytes(b"C2DF")
    db4 = index_build.from_npy_dir(tmp_path / "clip_vecs", bit, tmp_path / "faiss4", device=0)
    assert db4.index.ntotal == 1 and db4.ids == [f"{bit}/zz_new.c2df"]
    # opting out of the existence check indexes everything (additive)
    db5 = index_build.from_npy_dir(tmp_path / "clip_vecs", "nowhere", tmp_path / "faiss5", require_bitstream=False, device=0)
    assert db5.index.ntotal == 1


def _c2df_blobs(vecs, rng, full_size=True):
    """Reference-style .c2df files (src/compress.py:250-275: codec streams + clip_stream + clip_meta)."""
    from sgic_b200 import c2df
    from sgic_b200.index_build import quantize_u8_and_compress
    blobs = []
    for z in vecs:
        payload, meta = quantize_u8_and_compress(z)
        enc = {}
        if full_size:
            enc["z_bit_stream"] = bytes(rng.integers(0, 256, 700, dtype=np.uint8))
            enc["h_bit_stream"] = bytes(rng.integers(0, 256, 800, dtype=np.uint8))
            enc["img_shape"] = [1, 3, 256, 256]
            enc["token_length"] = 256
        enc["clip_stream"] = payload
        enc["clip_meta"] = meta
        blobs.append(c2df.pack_c2df(enc, {"version": 2, "note": "synthetic"}))
    return blobs


@pytest.mark.parametrize("d", [512, 768, 128])
def test_device_zstd_decode_equals_host_decode(d):
    """K0: clip_stream frames decoded on the device give bit-identical rows and statuses to
    libzstd on the host (src/search.py:35), frame by frame — frames with match sequences included — and the
    malformed files are skipped identically."""
    from sgic_b200 import faiss_compat as faiss, c2df
    rng = np.random.default_rng(d)
    n = 3000
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    vecs[::3] = vecs[::3] / np.linalg.norm(vecs[::3], axis=1, keepdims=True) + 1.0 / np.sqrt(d)   # CLIP-cone third
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    blobs = _c2df_blobs(vecs, rng, full_size=(d == 512))
    # malformed / foreign files sprinkled in: they must be skipped identically on both routes
    blobs[10] = b"XXXXnot a container"
    blobs[500] = c2df.pack_c2df({"token_length": 1}, {"version": 2})
    blobs[501] = blobs[501][:len(blobs[501]) // 2]
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(b) for b in blobs], out=offs[1:])
    blob = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    out = {}
    for mode in (1, 0):
        idx = faiss.IndexFlatIP(d, device=0)
        idx.set_option("device_zstd", mode)
        added, status = idx.add_c2df(blob, offs, n_threads=4)
        out[mode] = (added, status.copy(), idx.reconstruct_n(0, added), idx.stat("zl_device_frames"),
                     idx.stat("zl_host_rows"), idx.stat("zl_fallback_slabs"))
        idx.close()
    assert out[1][0] == out[0][0] == n - 3
    assert np.array_equal(out[1][1], out[0][1])
    assert np.array_equal(out[1][2], out[0][2])            # bit-identical database rows
    assert out[0][3] == 0                                  # host mode never launches K0
    dev, host = out[1][3], out[1][4]
    assert dev + host == n - 3 and out[1][5] == 0
    assert host == 0, (dev, host)   # every reference-style frame, match sequences included, decodes on the device
    # and the rows are what the reference decodes
    want = c2df_ref.dequantize_clip_u8(
        np.stack([np.frombuffer(__import__("sgic_b200").zstd.decompress(c2df.unpack_c2df(b)[0]["clip_stream"]),
                                dtype=np.uint8) for b in blobs[:8]]))
    assert fp16_ulp_diff(out[1][2][:8], want).max() <= 1


def test_device_zstd_corrupt_frame_falls_back_to_libzstd():
    """A damaged entropy stream inside an in-profile frame: the device refuses it, the slab is redone by
    libzstd (the decoder the reference uses), and the outcome (status codes, rows, order) equals the host-only
    route."""
    from sgic_b200 import faiss_compat as faiss, c2df
    rng = np.random.default_rng(99)
    n, d = 600, 512
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    blobs = _c2df_blobs(vecs, rng, full_size=False)
    damaged = 0
    for i in range(0, n, 37):
        enc, hdr = c2df.unpack_c2df(blobs[i])
        fr = bytearray(enc["clip_stream"])
        fr[len(fr) - 2 - (i % 40)] ^= 0x5A            # inside the Huffman streams
        enc["clip_stream"] = bytes(fr)
        blobs[i] = c2df.pack_c2df(enc, hdr)
        damaged += 1
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(b) for b in blobs], out=offs[1:])
    blob = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    res = {}
    for mode in (1, 0):
        idx = faiss.IndexFlatIP(d, device=0)
        idx.set_option("device_zstd", mode)
        added, status = idx.add_c2df(blob, offs, n_threads=2)
        res[mode] = (added, status.copy(), idx.reconstruct_n(0, added), idx.stat("zl_fallback_slabs"))
        idx.close()
    assert res[1][0] == res[0][0] and np.array_equal(res[1][1], res[0][1]) and np.array_equal(res[1][2], res[0][2])
    # libzstd 1.5.5 regenerates (different) bytes from these frames without complaint; the device decoder checks
    # that every stream is consumed exactly (RFC 8878 4.2.2), refuses, and hands the slab to libzstd — so the
    # reference's behaviour is reproduced either way
    assert res[1][3] >= 1


def test_sgi2_shard_file_round_trip(tmp_path):
    """N3: the v2 on-disk layout holds the rows exactly as they sit in HBM — write, read back through the
    faiss-named `read_index` (format recognised by its magic), same bits, same answers; IxFI keeps working."""
    import os
    from sgic_b200 import faiss_compat as faiss
    rng = np.random.default_rng(77)
    for dtype in ("fp16", "bf16"):
        n, d = 70_001, 512                     # > one 32 MB staging chunk, ragged
        xb = rng.standard_normal((n, d)).astype(np.float32)
        xb /= np.linalg.norm(xb, axis=1, keepdims=True)
        idx = faiss.IndexFlatIP(d, dtype=dtype, device=0)
        idx.add(xb)
        p2 = tmp_path / f"rows-{dtype}.sgi2"
        faiss.write_shard(idx, p2, row_start=1000, total_rows=500_000, shard=3, n_shards=8)
        assert os.path.getsize(p2) == 4096 + n * d * 2
        raw = np.fromfile(p2, dtype=np.uint8)
        assert bytes(raw[:4]) == b"SGI2"
        back = faiss.read_index(str(p2), device=0)
        assert back.ntotal == n and back.d == d and back.dtype == dtype
        assert faiss.shard_info(back) == {"row_start": 1000, "total_rows": 500_000, "shard": 3, "n_shards": 8}
        assert np.array_equal(back.reconstruct_n(0, n), idx.reconstruct_n(0, n))       # bit-identical rows
        q = xb[:7] + 0.01
        D0, I0 = idx.search(q, 10)
        D1, I1 = back.search(q, 10)
        assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
        # payload bytes are the HBM rows: fp16 file == numpy's fp16 rounding of the input
        if dtype == "fp16":
            assert np.array_equal(raw[4096:].view(np.float16).reshape(n, d), xb.astype(np.float16))
        # a truncated file is rejected like a short IxFI file
        (tmp_path / "short.sgi2").write_bytes(raw[:4096 + 1000].tobytes())
        with pytest.raises(RuntimeError):
            faiss.read_index(str(tmp_path / "short.sgi2"), device=0)
        # the interchange format still round-trips
        p1 = tmp_path / f"rows-{dtype}.faiss"
        faiss.write_index(idx, str(p1))
        again = faiss.read_index(str(p1), dtype=dtype, device=0)
        assert again.ntotal == n


def test_resident_service_over_the_real_index(tmp_path):
    """N2 end to end on the GPU: build from .c2df files, keep the index resident, answer concurrent .c2df
    queries through one batched search, emit the reference's stdout document."""
    import threading
    from sgic_b200 import index_build, retrieval
    from sgic_b200.service import SearchService
    rng = np.random.default_rng(3)
    vecs = rng.standard_normal((300, 512)).astype(np.float32)
    vecs /= np.linalg.norm(vecs, axis=1, keepdims=True)
    src = tmp_path / "bitstreams"
    src.mkdir()
    _write_corpus(src, vecs, rng)
    out = tmp_path / "faiss"
    index_build.build_index_from_c2df_dir(src, out)
    svc = SearchService(out, max_batch=64, max_wait_ms=30.0)
    try:
        paths = svc.paths
        n = 40
        res = [None] * n
        bar = threading.Barrier(n)

        def work(i):
            bar.wait()
            res[i] = svc.search_c2df(paths[i], topk=5)

        th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
        [t.start() for t in th]
        [t.join() for t in th]
        index, _, _ = retrieval.load_index(out)
        for i in range(n):
            want = retrieval.do_search(retrieval.encode_c2df_query(paths[i]), index, paths, topk=5)
            assert [p for p, _ in res[i]] == [p for p, _ in want] and res[i][0][0] == paths[i]
            assert np.allclose([s for _, s in res[i]], [s for _, s in want], atol=2e-5)   # K4 batch vs K3 single
        assert svc.stats["batches"] < n / 2 and svc.stats["max_batch_seen"] >= 5, svc.stats
        doc = json.loads(svc.cli_json(res[0]))
        assert doc[0]["path"] == paths[0] and set(doc[0]) == {"path", "score"}
    finally:
        svc.close()


def test_config_c1_10k_vectors_100_queries_one_at_a_time(tmp_path):
    """BASELINE.json configs[0]: 10k x 512 vectors in the reference's on-disk layouts (IO/clip_vecs/*.npy + an IxFI
    index.faiss + ids.txt), 100 text-like queries answered ONE AT A TIME through load_index -> do_search
    (src/search.py:65-88,113-120), k = 10.  The oracle holds what FAISS would hold: the fp32 rows of the file."""
    from sgic_b200 import retrieval
    from oracle.flat_ip import flat_ip_search
    rng = np.random.default_rng(101)
    n, d, nq, k = 10_000, 512, 100, 10
    cone = rng.standard_normal(d).astype(np.float32)
    cone /= np.linalg.norm(cone)
    xb = (cone + rng.standard_normal((n, d)).astype(np.float32) / np.sqrt(d)).astype(np.float32)   # CLIP-like cone
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xb[5000:5010] = xb[10:20]                                                                     # exact duplicates
    ids = [f"../IO/bitstreams/img_{i:05d}.c2df" for i in range(n)]
    c2df_ref.write_ixfi(tmp_path / "index.faiss", xb)
    (tmp_path / "ids.txt").write_text("".join(s + "\n" for s in ids))                             # compress.py:112-114
    index, paths, meta = retrieval.load_index(tmp_path)
    assert index.ntotal == n and index.d == d and paths == ids
    xq = (cone + 1.5 * rng.standard_normal((nq, d)).astype(np.float32) / np.sqrt(d)).astype(np.float32)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    xq[:10] = xb[10:20]                                                                           # queries with ties
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    for i in range(nq):
        res = retrieval.do_search(xq[i:i + 1], index, paths, topk=k)
        assert len(res) == k and all(isinstance(s, float) for _, s in res)
        I[i] = [ids.index(p) for p, _ in res]
        D[i] = [s for _, s in res]
    check_topk(D, I, xb, xq, k, score_tol=1e-3)          # O-ref: fp32 rows, north_star tolerance
    Dref, Iref = flat_ip_search(xb, xq, k)
    for i in range(10):                                  # duplicates: both copies lead, lowest row first
        assert list(I[i, :2]) == [10 + i, 5000 + i] and list(Iref[i, :2]) == [10 + i, 5000 + i]
    assert (I == Iref).mean() > 0.9                       # fp16 storage may swap near-ties (<1e-3 apart) only
