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
