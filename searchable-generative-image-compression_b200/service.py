"""Persistent search service — the caller side of the hot path.

The reference answers every web request by spawning ``python ./src/search.py query-… --index_dir D``
(webapp.py:246-248, :274-277, :306-309): a new interpreter, a new ``faiss.read_index`` of the whole fp32 file
(src/search.py:69,76), one search, exit.  Here the index is loaded ONCE and stays resident in HBM; requests
from any number of threads are micro-batched into one ``index.search`` call (so concurrent traffic reaches the
tensor-core regime instead of running batch-1 scans back to back), and the answers keep the reference's two
wire formats:

* :meth:`SearchService.cli_json` — the exact stdout document of ``search.py`` (src/search.py:163-166:
  ``json.dumps([{"path", "score"}, …], ensure_ascii=False, indent=2)``), so ``webapp.py`` can replace its
  ``subprocess.run(cmd, …).stdout`` with an in-process call;
* :meth:`SearchService.ndjson_events` — the NDJSON event stream of webapp.py:243-261 (``meta/start``,
  ``meta/searched``, ``item`` × n, ``done`` | ``error``), and :func:`serve` speaks it over a TCP socket, one
  JSON request per line.

``.c2df`` queries are decoded here exactly as ``encode_c2df_query`` does.  Text and image queries arrive as
embeddings (``vec`` / ``vec_path``) or — when the service was given an ``encoder``
(:class:`~.query_encoders.ClipQueryEncoder`, src/search.py:93-105 on the index's GPU) — as the text itself / the path
of an image, like the reference's ``query-text`` / ``query-image`` commands (src/search.py:153-162).
"""
from __future__ import annotations

import json
import queue
import socketserver
import threading
import time
from pathlib import Path
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["SearchService", "serve"]


class _Pending:
    __slots__ = ("q", "topk", "done", "result", "error")

    def __init__(self, q: np.ndarray, topk: int):
        self.q, self.topk = q, topk
        self.done = threading.Event()
        self.result: Optional[List[Tuple[str, float]]] = None
        self.error: Optional[BaseException] = None


class SearchService:
    """Index resident in HBM + micro-batching front end.

    ``index`` / ``paths`` may be injected (tests, already-loaded shards); otherwise ``index_dir`` is loaded with
    :func:`retrieval.load_index` — either naming scheme, IxFI or SGI2 file."""

    def __init__(self, index_dir=None, *, index=None, paths: Optional[Sequence[str]] = None, meta: Optional[Dict] = None,
                 max_batch: int = 256, max_wait_ms: float = 1.0,
                 preview_url: Optional[Callable[[str], Optional[str]]] = None,
                 max_topk: int = 1024, allowed_roots: Optional[Sequence] = None, encoder=None):
        if index is None:
            from .retrieval import load_index
            index, paths, meta = load_index(index_dir)
        self.index, self.paths, self.meta = index, list(paths), dict(meta or {})
        self.max_batch, self.max_wait = int(max_batch), float(max_wait_ms) / 1e3
        self.preview_url = preview_url or (lambda p: None)
        self.encoder = encoder                      # ClipQueryEncoder or None (then text / image queries need "vec")
        self._enc_lock = threading.Lock()           # one encoder pass at a time (one model, one stream)
        # Untrusted callers (the TCP front end): topk is clamped to `max_topk`, and server-side paths are only
        # opened below one of `allowed_roots` (none configured: no path requests over the socket, inline vectors
        # only).  In-process callers are trusted like the reference's CLI is.
        self.max_topk = max(1, int(max_topk))
        self.allowed_roots = [Path(r).resolve() for r in (allowed_roots or [])]
        self._q: "queue.Queue[Optional[_Pending]]" = queue.Queue()
        self.stats = {"requests": 0, "batches": 0, "max_batch_seen": 0}
        self._worker = threading.Thread(target=self._run, name="sgic-search-batcher", daemon=True)
        self._worker.start()

    # ------------------------------------------------------------------ batching worker
    def _run(self) -> None:
        while True:
            first = self._q.get()
            if first is None:
                return
            batch = [first]
            deadline = time.perf_counter() + self.max_wait
            # a request for more than `max_topk` results runs alone: k > 1024 is answered by repeated full scans,
            # one query at a time, and must not drag the ordinary requests of its micro-batch onto that path
            held: Optional[_Pending] = None
            while len(batch) < self.max_batch and first.topk <= self.max_topk:
                left = deadline - time.perf_counter()
                try:
                    nxt = self._q.get(timeout=max(left, 0.0)) if left > 0 else self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    self._q.put(None)
                    break
                if nxt.topk > self.max_topk:
                    held = nxt
                    break
                batch.append(nxt)
            try:
                ntotal = self.index.ntotal
                kmax = max(1, min(max(p.topk for p in batch), ntotal))      # src/search.py:114
                Q = np.concatenate([p.q for p in batch], axis=0).astype("float32", copy=False)
                sim, ids = self.index.search(Q, kmax)
                for r, p in enumerate(batch):
                    k = max(1, min(p.topk, ntotal))
                    p.result = [(self.paths[i], float(sim[r, j])) for j, i in enumerate(ids[r, :k]) if i != -1]
            except BaseException as e:  # every waiter gets the error, the worker keeps serving
                for p in batch:
                    p.error = e
            self.stats["requests"] += len(batch)
            self.stats["batches"] += 1
            self.stats["max_batch_seen"] = max(self.stats["max_batch_seen"], len(batch))
            for p in batch:
                p.done.set()
            if held is not None:
                self._q.put(held)       # next round, on its own

    def close(self) -> None:
        self._q.put(None)
        self._worker.join(timeout=5)

    # ------------------------------------------------------------------ queries
    def search_vec(self, q, topk: int = 10) -> List[Tuple[str, float]]:
        """``do_search`` (src/search.py:113-120) for one (1, d) / (d,) fp32 query; thread-safe, batched."""
        q = np.asarray(q, dtype="float32").reshape(1, -1)
        if q.shape[1] != self.index.d:
            raise ValueError(f"query has {q.shape[1]} dims, index has {self.index.d}")
        p = _Pending(q, int(topk))
        self._q.put(p)
        p.done.wait()
        if p.error is not None:
            raise p.error
        return p.result

    def search_c2df(self, c2df_path, topk: int = 10) -> List[Tuple[str, float]]:
        from .retrieval import encode_c2df_query
        return self.search_vec(encode_c2df_query(c2df_path), topk)

    def search_vec_file(self, npy_path, topk: int = 10) -> List[Tuple[str, float]]:
        from .retrieval import l2n
        v = np.load(npy_path).astype("float32").reshape(1, -1)
        return self.search_vec(l2n(v).astype("float32"), topk)

    def _need_encoder(self, kind: str):
        if self.encoder is None:
            raise NotImplementedError(f"{kind} query without an embedding: no CLIP encoder was configured "
                                      "(send \"vec\" or \"vec_path\", or start the service with one)")
        return self.encoder

    def search_text(self, text: str, topk: int = 10) -> List[Tuple[str, float]]:
        """``query-text`` (src/search.py:153-156, :93-97): embed on the GPU, then the batched search."""
        enc = self._need_encoder("text")
        with self._enc_lock:
            z = enc.encode_text([text]).cpu().numpy()
        return self.search_vec(z, topk)

    def search_image(self, image_path, topk: int = 10) -> List[Tuple[str, float]]:
        """``query-image`` (src/search.py:157-160, :100-105)."""
        enc = self._need_encoder("image")
        from PIL import Image
        im = Image.open(image_path).convert("RGB")
        with self._enc_lock:
            z = enc.encode_image([im]).cpu().numpy()
        return self.search_vec(z, topk)

    # ------------------------------------------------------------------ the reference's wire formats
    @staticmethod
    def cli_json(results: Iterable[Tuple[str, float]]) -> str:
        """stdout of ``search.py`` (src/search.py:163-166), what webapp.py:249 feeds to ``json.loads``."""
        return json.dumps([{"path": p, "score": s} for p, s in results], ensure_ascii=False, indent=2)

    def _checked_path(self, p, trusted: bool) -> Path:
        """Server-side path of a request.  Untrusted requests may only name files below an allowed root."""
        path = Path(str(p))
        if trusted:
            return path
        real = path.resolve()
        for root in self.allowed_roots:
            if real == root or root in real.parents:
                return real
        raise PermissionError("path outside the served roots")

    @staticmethod
    def _public_error(e: BaseException) -> str:
        """What an untrusted client is told: the kind of failure, never the exception text (which can carry
        server-side paths and library internals)."""
        if isinstance(e, PermissionError):
            return "path not allowed"
        if isinstance(e, FileNotFoundError):
            return "file not found"
        if isinstance(e, NotImplementedError):
            return "query type needs an embedding (send \"vec\")"
        if isinstance(e, (ValueError, AssertionError, KeyError, TypeError)):
            return "bad request"
        return "search failed"

    def ndjson_events(self, request: Dict, trusted: bool = True) -> Iterator[Dict]:
        """Event dictionaries of one streamed search, in the order webapp.py:243-261 yields them.
        ``trusted=False`` (the TCP front end): topk clamped to ``max_topk``, paths checked against
        ``allowed_roots``, error details reduced to a category."""
        t0 = time.perf_counter()
        kind = request.get("type") or request.get("query_type") or "c2df"
        try:
            topk = int(request.get("topk") or 10)
        except (TypeError, ValueError):
            topk = 10
        if not trusted:
            topk = max(1, min(topk, self.max_topk))
        start = {"type": "meta", "stage": "start", "query_type": kind, "topk": topk}
        if kind == "text":
            start["query"] = request.get("text", "")
        else:
            start["filename"] = request.get("filename") or Path(str(request.get("path", ""))).name
        yield start
        try:
            if "vec" in request:
                items = self.search_vec(np.asarray(request["vec"], dtype="float32"), topk)
            elif "vec_path" in request:
                items = self.search_vec_file(self._checked_path(request["vec_path"], trusted), topk)
            elif kind == "c2df":
                items = self.search_c2df(self._checked_path(request["path"], trusted), topk)
            elif kind == "text":
                items = self.search_text(str(request.get("text", "")), topk)
            elif kind == "image":
                self._need_encoder("image")
                items = self.search_image(self._checked_path(request["path"], trusted), topk)
            else:
                raise ValueError(f"unknown query type {kind!r}")
            ms = lambda: int((time.perf_counter() - t0) * 1000)
            yield {"type": "meta", "stage": "searched", "count": len(items), "elapsed_ms": ms()}
            for p, s in items:
                yield {"type": "item", "path": p, "score": float(s), "preview_url": self.preview_url(p)}
            yield {"type": "done", "elapsed_ms": ms()}
        except Exception as e:
            yield {"type": "error", "detail": str(e) if trusted else self._public_error(e)}


class _Handler(socketserver.StreamRequestHandler):
    def handle(self) -> None:
        svc: SearchService = self.server.service  # type: ignore[attr-defined]
        for line in self.rfile:
            line = line.strip()
            if not line:
                continue
            try:
                req = json.loads(line)
            except ValueError as e:
                self.wfile.write((json.dumps({"type": "error", "detail": "bad request"}) + "\n").encode())
                continue
            if not isinstance(req, dict):
                self.wfile.write((json.dumps({"type": "error", "detail": "bad request"}) + "\n").encode())
                continue
            for ev in svc.ndjson_events(req, trusted=False):
                self.wfile.write((json.dumps(ev, ensure_ascii=False) + "\n").encode("utf-8"))   # webapp._yield_ndjson
            self.wfile.flush()


class _Server(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True


def serve(service: SearchService, host: str = "127.0.0.1", port: int = 0) -> _Server:
    """Start the NDJSON socket front end in a background thread; returns the server (``server_address`` has the
    bound port, ``shutdown()`` stops it).  One connection per client thread; all of them share the batcher."""
    srv = _Server((host, port), _Handler)
    srv.service = service  # type: ignore[attr-defined]
    threading.Thread(target=srv.serve_forever, name="sgic-search-server", daemon=True).start()
    return srv


def main(argv=None) -> None:
    import argparse
    ap = argparse.ArgumentParser(description="resident search service (NDJSON over TCP)")
    ap.add_argument("--index_dir", type=Path, required=True)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8765)
    ap.add_argument("--max_batch", type=int, default=256)
    ap.add_argument("--max_wait_ms", type=float, default=1.0)
    ap.add_argument("--max_topk", type=int, default=1024, help="largest topk a socket client may ask for")
    ap.add_argument("--allow_root", type=Path, action="append", default=[],
                    help="directory whose files socket clients may name in \"path\" / \"vec_path\" (repeatable)")
    ap.add_argument("--clip_dir", type=Path, default=None,
                    help="local CLIP checkpoint directory: text / image queries are embedded on the index's GPU")
    a = ap.parse_args(argv)
    svc = SearchService(a.index_dir, max_batch=a.max_batch, max_wait_ms=a.max_wait_ms, max_topk=a.max_topk,
                        allowed_roots=a.allow_root)
    from .retrieval import resolve_clip_dir
    clip_dir = resolve_clip_dir(a.clip_dir, svc.meta)
    if clip_dir is not None:
        from .query_encoders import ClipQueryEncoder
        svc.encoder = ClipQueryEncoder(clip_dir, device=svc.index.device)
    srv = serve(svc, a.host, a.port)
    print(json.dumps({"listening": list(srv.server_address), "ntotal": svc.index.ntotal, "d": svc.index.d}), flush=True)
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        srv.shutdown()
        svc.close()


if __name__ == "__main__":
    main()
