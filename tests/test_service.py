"""Host logic of the resident search service on CPU: micro-batching, the reference's stdout
document (src/search.py:163-166) and NDJSON event stream (webapp.py:243-261).  The index is a stand-in backed
by the oracle; the GPU tier runs the same service over the real index."""
import json
import socket
import threading

import numpy as np
import pytest

from oracle.flat_ip import flat_ip_search
from sgic_b200.service import SearchService, serve


class OracleIndex:
    """IndexFlatIP stand-in (test infrastructure): counts calls and batch sizes."""

    def __init__(self, xb):
        self.xb, self.d, self.ntotal = xb, xb.shape[1], xb.shape[0]
        self.calls = []

    def search(self, q, k):
        self.calls.append((q.shape[0], k))
        return flat_ip_search(self.xb, q, k)


@pytest.fixture()
def svc():
    rng = np.random.default_rng(0)
    xb = rng.standard_normal((500, 64)).astype(np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    idx = OracleIndex(xb)
    s = SearchService(index=idx, paths=[f"IO/bitstreams/img{i:04d}.c2df" for i in range(500)],
                      meta={"dim": 64, "model_id": None}, max_batch=64, max_wait_ms=20.0)
    yield s, idx, xb
    s.close()


def test_single_query_equals_do_search(svc):
    s, idx, xb = svc
    res = s.search_vec(xb[7], topk=5)
    D, I = flat_ip_search(xb, xb[7:8], 5)
    assert [p for p, _ in res] == [f"IO/bitstreams/img{i:04d}.c2df" for i in I[0]]
    assert np.allclose([sc for _, sc in res], D[0])
    assert res[0][0].endswith("img0007.c2df") and abs(res[0][1] - 1.0) < 1e-5
    # topk larger than the index is clamped like src/search.py:114
    assert len(s.search_vec(xb[0], topk=10_000)) == 500
    with pytest.raises(ValueError):
        s.search_vec(np.zeros(63, np.float32))


def test_concurrent_requests_are_batched_into_one_search(svc):
    s, idx, xb = svc
    n = 48
    out = [None] * n
    barrier = threading.Barrier(n)

    def worker(i):
        barrier.wait()
        out[i] = s.search_vec(xb[i], topk=3 + (i % 4))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(n):
        assert out[i][0][0].endswith(f"img{i:04d}.c2df") and len(out[i]) == 3 + (i % 4)
    assert s.stats["requests"] == n
    assert s.stats["batches"] < n / 4 and s.stats["max_batch_seen"] >= 8, s.stats   # far fewer searches than requests
    assert all(k == 6 or nq == 1 or k <= 6 for nq, k in idx.calls)


def test_cli_json_is_the_reference_stdout_document(svc):
    s, idx, xb = svc
    res = s.search_vec(xb[3], topk=2)
    doc = s.cli_json(res)
    assert doc == json.dumps([{"path": p, "score": sc} for p, sc in res], ensure_ascii=False, indent=2)
    assert [set(e) for e in json.loads(doc)] == [{"path", "score"}] * 2


def test_ndjson_event_stream_and_socket_front_end(svc):
    s, idx, xb = svc
    ev = list(s.ndjson_events({"type": "text", "text": "a red apple", "vec": xb[11].tolist(), "topk": 4}))
    assert ev[0] == {"type": "meta", "stage": "start", "query_type": "text", "topk": 4, "query": "a red apple"}
    assert ev[1]["type"] == "meta" and ev[1]["stage"] == "searched" and ev[1]["count"] == 4 and "elapsed_ms" in ev[1]
    assert [e["type"] for e in ev[2:6]] == ["item"] * 4 and set(ev[2]) == {"type", "path", "score", "preview_url"}
    assert ev[2]["path"].endswith("img0011.c2df") and ev[-1]["type"] == "done"
    # a text query without an embedding cannot be served by this path: reported like webapp.py:258-259
    bad = list(s.ndjson_events({"type": "text", "text": "x"}))
    assert bad[-1]["type"] == "error"
    srv = serve(s, port=0)
    try:
        with socket.create_connection(srv.server_address, timeout=10) as c:
            f = c.makefile("rwb")
            f.write((json.dumps({"type": "image", "filename": "q.png", "vec": xb[21].tolist(), "topk": 2}) + "\n").encode())
            f.write(b"not json\n")
            f.flush()
            lines = []
            while True:
                lines.append(json.loads(f.readline()))
                if lines[-1]["type"] == "error":
                    break
        kinds = [l["type"] for l in lines]
        assert kinds == ["meta", "meta", "item", "item", "done", "error"]
        assert lines[0]["filename"] == "q.png" and lines[2]["path"].endswith("img0021.c2df")
    finally:
        srv.shutdown()
        srv.server_close()


def test_untrusted_requests_are_clamped_confined_and_told_little(svc, tmp_path):
    """The TCP front end: topk clamped to max_topk, server-side paths only below an allowed root, no exception
    text in error events; a huge-topk request from a trusted caller runs in a batch of its own."""
    s, idx, xb = svc
    s.max_topk = 8
    ev = list(s.ndjson_events({"type": "image", "vec": xb[5].tolist(), "topk": 10 ** 9}, trusted=False))
    assert ev[0]["topk"] == 8 and ev[1]["count"] == 8 and ev[-1]["type"] == "done"
    # paths: nothing is served from the file system unless a root is configured, and never outside it
    secret = tmp_path / "secret.npy"
    np.save(secret, xb[9])
    ev = list(s.ndjson_events({"type": "image", "vec_path": str(secret), "topk": 2}, trusted=False))
    assert ev[-1] == {"type": "error", "detail": "path not allowed"}
    root = tmp_path / "served"
    root.mkdir()
    np.save(root / "q.npy", xb[9])
    s.allowed_roots = [root.resolve()]
    ev = list(s.ndjson_events({"type": "image", "vec_path": str(root / "q.npy"), "topk": 2}, trusted=False))
    assert ev[-1]["type"] == "done" and ev[2]["path"].endswith("img0009.c2df")
    ev = list(s.ndjson_events({"type": "image", "vec_path": str(root / ".." / "secret.npy"), "topk": 2}, trusted=False))
    assert ev[-1] == {"type": "error", "detail": "path not allowed"}
    ev = list(s.ndjson_events({"type": "c2df", "path": str(root / "missing.c2df")}, trusted=False))
    assert ev[-1] == {"type": "error", "detail": "file not found"}
    assert str(root) not in json.dumps(ev)
    ev = list(s.ndjson_events({"type": "image", "vec": [1.0, 2.0], "topk": "many"}, trusted=False))
    assert ev[-1] == {"type": "error", "detail": "bad request"} and ev[0]["topk"] == 8
    # trusted callers keep the reference's semantics (any topk), but a huge request never shares a batch
    idx.calls.clear()
    out = {}
    barrier = threading.Barrier(9)

    def worker(i, k):
        barrier.wait()
        out[i] = s.search_vec(xb[i], topk=k)

    th = [threading.Thread(target=worker, args=(i, 3)) for i in range(8)] + [threading.Thread(target=worker, args=(8, 400))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert len(out[8]) == 400 and all(len(out[i]) == 3 for i in range(8))
    assert (1, 400) in idx.calls and all(k == 3 for nq, k in idx.calls if (nq, k) != (1, 400))


class _CharTokenizer:
    """Stand-in vocabulary for the randomly initialised CLIP of the tests (no checkpoint can be downloaded)."""

    def __call__(self, texts, padding=None, truncation=None, return_tensors=None):
        import torch
        ids = torch.zeros((len(texts), 8), dtype=torch.long)
        for r, t in enumerate(texts):
            for c, ch in enumerate(t[:8]):
                ids[r, c] = 1 + (ord(ch) % 90)
        return {"input_ids": ids}


def test_text_and_image_queries_are_embedded_by_the_service(tmp_path):
    """query-text / query-image (src/search.py:153-162) through the resident service: the encoder of
    query_encoders in front of the batched search; socket clients may only name images below an allowed root."""
    from PIL import Image
    from transformers import CLIPImageProcessor
    from sgic_b200.query_encoders import ClipQueryEncoder
    from test_query_encoders import tiny_clip
    enc = ClipQueryEncoder(tiny_clip(), device=0, tokenizer=_CharTokenizer(),
                           image_processor=CLIPImageProcessor(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32}))
    texts = [f"{chr(97 + i)} thing {i}" for i in range(20)]
    rng = np.random.default_rng(2)
    (tmp_path / "served").mkdir()
    imgs = []
    for i in range(5):
        p = tmp_path / "served" / f"q{i}.png"
        Image.fromarray(rng.integers(0, 256, (36, 36, 3), dtype=np.uint8)).save(p)
        imgs.append(p)
    outside = tmp_path / "elsewhere.png"
    Image.fromarray(rng.integers(0, 256, (36, 36, 3), dtype=np.uint8)).save(outside)
    rows = np.concatenate([enc.encode_text(texts).cpu().numpy(),
                           enc.encode_image([Image.open(p).convert("RGB") for p in imgs]).cpu().numpy()])
    paths = [f"text{i}" for i in range(20)] + [str(p) for p in imgs]
    s = SearchService(index=OracleIndex(rows), paths=paths, meta={"dim": 64}, max_wait_ms=1.0, encoder=enc,
                      allowed_roots=[tmp_path / "served"])
    try:
        assert s.search_text(texts[7], topk=3)[0][0] == "text7"
        assert s.search_image(imgs[2], topk=1)[0][0] == str(imgs[2])
        ev = list(s.ndjson_events({"type": "text", "text": texts[3], "topk": 2}, trusted=False))
        assert ev[0]["query"] == texts[3] and ev[2]["path"] == "text3" and ev[-1]["type"] == "done"
        ev = list(s.ndjson_events({"type": "image", "path": str(imgs[4]), "topk": 1}, trusted=False))
        assert ev[0]["filename"] == "q4.png" and ev[2]["path"] == str(imgs[4])
        ev = list(s.ndjson_events({"type": "image", "path": str(outside)}, trusted=False))
        assert ev[-1] == {"type": "error", "detail": "path not allowed"}
        ev = list(s.ndjson_events({"type": "audio", "path": str(imgs[0])}, trusted=False))
        assert ev[-1] == {"type": "error", "detail": "bad request"}
        s.encoder = None
        ev = list(s.ndjson_events({"type": "text", "text": "x"}, trusted=False))
        assert ev[-1]["type"] == "error" and "embedding" in ev[-1]["detail"]
    finally:
        s.close()
