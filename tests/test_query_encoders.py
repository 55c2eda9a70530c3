"""N4: CLIP query encoders in front of the search call (src/search.py:48-105).  No checkpoint can
be downloaded here, so the plumbing is exercised with a small randomly initialised CLIP: unit-norm fp32 embeddings
of the projection width, text and image towers, batches; on a GPU the embeddings go from the encoder straight into
``search_torch`` and match the host route of the reference (encode → numpy → index.search)."""
import numpy as np
import pytest


def tiny_clip(proj=64):
    import torch
    from transformers import CLIPConfig, CLIPModel
    torch.manual_seed(0)
    cfg = CLIPConfig(text_config=dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                                      vocab_size=100, max_position_embeddings=16),
                     vision_config=dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                                        image_size=32, patch_size=16), projection_dim=proj)
    return CLIPModel(cfg)


def test_encoders_give_unit_fp32_vectors_of_the_projection_width():
    import torch
    from sgic_b200.query_encoders import ClipQueryEncoder
    enc = ClipQueryEncoder(tiny_clip(), device=0)
    ids = torch.randint(0, 99, (5, 8))
    zt = enc.encode_text_ids(ids)
    zi = enc.encode_pixels(torch.randn(3, 3, 32, 32))
    assert zt.shape == (5, 64) and zi.shape == (3, 64) and zt.dtype == torch.float32 and enc.dim == 64
    assert torch.allclose(zt.norm(dim=1), torch.ones(5, device=zt.device), atol=1e-5)
    assert torch.allclose(zi.norm(dim=1), torch.ones(3, device=zi.device), atol=1e-5)
    # the same numbers as the model's own normalised embeddings (src/search.py:95-96 does this by hand)
    with torch.no_grad():
        full = enc.model(input_ids=ids.to(enc.device), pixel_values=torch.randn(5, 3, 32, 32).to(enc.device))
    assert torch.allclose(full.text_embeds, zt, atol=1e-5)
    with pytest.raises(RuntimeError, match="tokenizer"):
        enc.encode_text(["a red apple"])


@pytest.mark.gpu
def test_device_resident_query_equals_the_host_route():
    import torch
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.query_encoders import ClipQueryEncoder
    from sgic_b200.retrieval import do_search
    enc = ClipQueryEncoder(tiny_clip(), device=0)
    # "images" of a corpus: their embeddings are the index rows; an image query must find itself
    px = torch.randn(400, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    rows = enc.encode_pixels(px)
    index = faiss.IndexFlatIP(64, device=0)
    index.add(rows.cpu().numpy())
    paths = [f"../IO/bitstreams/img{i:04d}.c2df" for i in range(400)]
    res = enc.search(index, enc.encode_pixels(px[7:9]), paths, topk=5)
    assert [r[0][0] for r in res] == [paths[7], paths[8]] and all(abs(r[0][1] - 1.0) < 2e-3 for r in res)
    # text queries, a batch of 6: device route == reference route (encode, .cpu().numpy(), do_search per query)
    ids = torch.randint(0, 99, (6, 8), generator=torch.Generator().manual_seed(2))
    z = enc.encode_text_ids(ids)
    D, I = enc.search(index, z, topk=10)
    for r in range(6):
        want = do_search(z[r:r + 1].cpu().numpy().astype("float32"), index, paths, topk=10)
        assert [paths[i] for i in I[r].tolist()] == [p for p, _ in want]
    with pytest.raises(ValueError, match="dims"):
        enc.search(faiss.IndexFlatIP(128, device=0), z)


def _image_corpus(root, n=11, size=40):
    from PIL import Image
    rng = np.random.default_rng(5)
    (root / "sub").mkdir(parents=True)
    names = []
    for i in range(n):
        ext = (".png", ".JPG", ".bmp")[i % 3]
        p = (root / "sub" if i % 4 == 0 else root) / f"im{i:03d}{ext}"
        Image.fromarray(rng.integers(0, 256, (size, size + i, 3), dtype=np.uint8)).save(p, format={".png": "PNG", ".JPG": "JPEG", ".bmp": "BMP"}[ext])
        names.append(p)
    (root / "notes.txt").write_text("not an image")
    (root / "broken.png").write_bytes(b"\x89PNG but not really")
    return names


def _tiny_encoder():
    from transformers import CLIPImageProcessor
    from sgic_b200.query_encoders import ClipQueryEncoder
    proc = CLIPImageProcessor(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32})
    return ClipQueryEncoder(tiny_clip(), device=0, image_processor=proc)


def test_list_images_and_batched_embedding_follow_the_reference(tmp_path, capsys):
    """src/build.py:171-205: rglob order, case-insensitive suffixes, unreadable files skipped with a message."""
    from sgic_b200.index_build import encode_images_in_batches, list_images
    names = _image_corpus(tmp_path)
    got = list_images(tmp_path)
    assert set(got) == set(names) | {tmp_path / "broken.png"} and got == [p for p in tmp_path.rglob("*") if p in set(got)]
    assert list_images(tmp_path, exts={"PNG"}) == [p for p in got if p.suffix.lower() == ".png"]
    enc = _tiny_encoder()
    X = encode_images_in_batches(got, enc, batch_size=4)
    assert "[SKIP] Can't read the image" in capsys.readouterr().out
    assert X.shape == (len(names), 64) and X.dtype == np.float32
    assert np.allclose(np.linalg.norm(X, axis=1), 1.0, atol=1e-5)
    one = encode_images_in_batches([p for p in got if p.name != "broken.png"][:3], enc, batch_size=32)
    assert np.allclose(one, X[:3], atol=1e-5)         # the batch size does not change a row
    assert encode_images_in_batches([tmp_path / "broken.png"], enc) is None


@pytest.mark.gpu
def test_build_index_from_image_dir_writes_the_reference_layout(tmp_path):
    """build-images (src/build.py:207-241): four files, both naming schemes; the IxFI bytes are the encoder's fp32
    rows exactly; an image query finds its own file through load_index + do_search."""
    import json
    from PIL import Image
    from oracle import c2df_ref
    from sgic_b200 import faiss_compat as faiss
    from sgic_b200.index_build import build_index_from_image_dir, encode_images_in_batches, list_images
    from sgic_b200.retrieval import do_search, load_index
    img_dir, out = tmp_path / "images", tmp_path / "index"
    img_dir.mkdir()
    names = _image_corpus(img_dir)
    (img_dir / "broken.png").unlink()
    enc = _tiny_encoder()
    build_index_from_image_dir(img_dir, out, "tiny-random-clip", "cuda", batch_size=4, encoder=enc)
    listed = [str(p) for p in list_images(img_dir)]
    assert sorted(p.name for p in out.iterdir()) == ["faiss.index", "ids.txt", "index.faiss", "meta.json", "paths.json"]
    assert json.loads((out / "paths.json").read_text(encoding="utf-8")) == listed
    assert (out / "ids.txt").read_text(encoding="utf-8") == "\n".join(listed)
    assert json.loads((out / "meta.json").read_text()) == {"dim": 64, "model_id": "tiny-random-clip"}
    X = encode_images_in_batches(list_images(img_dir), enc, batch_size=4)
    c2df_ref.write_ixfi(tmp_path / "want.index", X)
    assert (out / "faiss.index").read_bytes() == (tmp_path / "want.index").read_bytes() == (out / "index.faiss").read_bytes()
    index, paths, meta = load_index(out)
    q = enc.encode_image([Image.open(names[3]).convert("RGB")]).cpu().numpy()
    res = do_search(q, index, paths, topk=3)
    assert res[0][0] == str(names[3]) and abs(res[0][1] - 1.0) < 2e-3
    index.close()
    # selection: the first `limit` images in listing order; `desired` with auto_download cannot be served offline
    build_index_from_image_dir(img_dir, tmp_path / "few", None, 0, limit=4, encoder=enc)
    assert json.loads((tmp_path / "few" / "paths.json").read_text()) == listed[:4]
    with pytest.raises(RuntimeError, match="network"):
        build_index_from_image_dir(img_dir, tmp_path / "x", None, 0, desired=500, auto_download=True, encoder=enc)
    with pytest.raises(RuntimeError, match="There is no image"):
        build_index_from_image_dir(tmp_path / "index", tmp_path / "y", None, 0, encoder=enc)


def test_clip_codec_makes_the_retrieval_streams_of_a_c2df(tmp_path):
    """compress.py:57-86 + :262-281: image tensor -> unit vector -> (clip_stream, clip_meta); decoding the file gives
    the vector back within the u8 quantiser's step."""
    import torch
    from sgic_b200 import c2df
    from sgic_b200.index_build import ClipCodec
    from sgic_b200.retrieval import decode_clip_from_c2df
    codec = ClipCodec(_tiny_encoder(), model_name="tiny:random")
    img = torch.rand(3, 48, 40, generator=torch.Generator().manual_seed(3)) * 2 - 1
    z = codec.image_to_unit_vec(img)
    assert z.shape == (64,) and z.dtype == np.float32 and abs(np.linalg.norm(z) - 1.0) < 1e-5
    stream, meta = codec.quantize_u8_and_compress(z)
    assert meta == {"model_id": "tiny:random", "dim": 64, "quant": "u8_symmetric_-1_1", "codec": "zstd", "zstd_level": 19}
    blob = c2df.pack_c2df({"clip_stream": stream, "clip_meta": meta}, {"version": 2, "model_id": "tiny:random"})
    (tmp_path / "img.c2df").write_bytes(blob)
    back, header = decode_clip_from_c2df(tmp_path / "img.c2df")
    assert header["model_id"] == "tiny:random" and float(back @ z) > 0.995


def test_clip_checkpoint_resolution_order(tmp_path, monkeypatch):
    """--clip_dir, then $SGIC_CLIP_DIR, then meta["model_id"] when it is a directory (src/search.py:151-152 takes the
    model from meta["model_id"]); OpenCLIP names and missing directories resolve to nothing."""
    from sgic_b200.retrieval import resolve_clip_dir
    a, b, c = tmp_path / "a", tmp_path / "b", tmp_path / "c"
    for d in (a, b, c):
        d.mkdir()
    monkeypatch.delenv("SGIC_CLIP_DIR", raising=False)
    assert resolve_clip_dir(None, {"model_id": "ViT-B-32:laion2b_s34b_b79k"}) is None
    assert resolve_clip_dir(None, None) is None
    assert resolve_clip_dir(None, {"model_id": str(c)}) == c
    monkeypatch.setenv("SGIC_CLIP_DIR", str(b))
    assert resolve_clip_dir(None, {"model_id": str(c)}) == b
    assert resolve_clip_dir(a, {"model_id": str(c)}) == a
    assert resolve_clip_dir(tmp_path / "missing", {"model_id": None}) == b
