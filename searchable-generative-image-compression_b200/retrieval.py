"""Query side of the path — host-side mirror of the reference's ``src/search.py``.

Same function names, argument meaning, return shapes, error classes and stdout JSON
schema (``[{"path": ..., "score": ...}]``, ``indent=2, ensure_ascii=False`` —
src/search.py:163-166, parsed by webapp.py:249), so a caller of ``search.py`` can call
this instead.  ``index.search`` is the B200 kernel path (``faiss_compat``).

Out of scope here: CLIP text / image encoders — ``open_clip``
and its weights are not available offline.  ``query-text`` / ``query-image`` therefore
accept a pre-computed embedding (``--vec file.npy``); ``query-c2df`` is complete.
"""
from __future__ import annotations

import argparse
import os
import json
import sys
import traceback
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import faiss_compat as faiss
from . import zstd
from .c2df import unpack_c2df


def l2n(x: np.ndarray, axis: int = -1, eps: float = 1e-9) -> np.ndarray:
    """Unit-normalise along ``axis`` (src/search.py:16-18)."""
    norm = np.linalg.norm(x, axis=axis, keepdims=True)
    return x / np.maximum(norm, eps)


def dequantize_clip_u8(q: np.ndarray) -> np.ndarray:
    """u8 → [-1,1] → unit vector, fp32, in the reference's operation order
    ``(q/255)*2 - 1`` (src/search.py:20-22)."""
    z = q.astype(np.float32)
    z = z / 255.0
    z = z * 2.0
    z = z - 1.0
    return l2n(z.astype(np.float32))


def decode_clip_from_c2df(c2df_path) -> Tuple[np.ndarray, Dict]:
    """``.c2df`` → (fp32 unit vector of length dim, header dict) — src/search.py:24-41.

    Raises ``ValueError`` for a file without ``clip_stream`` / ``clip_meta``, for a
    non-positive ``clip_meta.dim`` and for a decoded length that differs from ``dim``;
    ``AssertionError`` for a bad magic (from :func:`unpack_c2df`).
    """
    entries, header = unpack_c2df(c2df_path)
    if "clip_stream" not in entries or "clip_meta" not in entries:
        raise ValueError(f"{c2df_path} No 'clip_stream' or 'clip_meta' was found, this file can't be used to search!")
    meta = entries["clip_meta"] or {}
    dim = int(meta.get("dim", 0))
    if dim <= 0:
        raise ValueError(f"{c2df_path} Invalid clip_meta.dim")
    q = np.frombuffer(zstd.decompress(entries["clip_stream"]), dtype=np.uint8)
    if q.size != dim:
        raise ValueError(f"{c2df_path} Dimension didn't match: q={q.size}, dim={dim}")
    return dequantize_clip_u8(q).astype("float32"), header


def _shard_files(index_dir: Path) -> List[Path]:
    """``shard-00000-of-0000G.sgi2`` … in order, when the directory holds one complete set (additive layout:
    ``IndexFlatIP.save_shards`` / ``ShardedIndexFlatIP.save``)."""
    files = sorted(index_dir.glob("shard-*-of-*.sgi2"))
    if not files:
        return []
    try:
        total = int(files[0].stem.split("-of-")[1])
    except (IndexError, ValueError):
        return []
    want = [index_dir / f"shard-{g:05d}-of-{total:05d}.sgi2" for g in range(total)]
    return want if all(p.exists() for p in want) else []


def _default_devices(n_files: int):
    env = os.environ.get("SGIC_DEVICES")
    if env:
        return [int(t) for t in env.split(",") if t.strip() != ""]
    return list(range(n_files))


def load_index(index_dir, devices=None) -> Tuple[faiss.Index, List[str], Dict]:
    """Open an index directory in either naming scheme — src/search.py:65-88.

    ``faiss.index`` + ``paths.json`` (+ optional ``meta.json``) is preferred; otherwise
    ``index.faiss`` + ``ids.txt`` with ``meta = {"dim": index.d}`` and ``model_id`` sniffed
    from the first listed ``.c2df`` header when it can be read.

    Additive: a directory that holds a set of ``shard-*-of-*.sgi2`` files next to ``paths.json`` / ``ids.txt`` is
    loaded onto ``devices`` (default: ``SGIC_DEVICES`` or GPUs 0..G-1), one file per GPU, behind one index object —
    the id list and ``meta`` are read exactly as for the single-file layouts.
    """
    index_dir = Path(index_dir)
    new_idx, old_idx = index_dir / "faiss.index", index_dir / "index.faiss"
    shards = _shard_files(index_dir)
    if shards and ((index_dir / "paths.json").exists() or (index_dir / "ids.txt").exists()):
        devs = list(devices) if devices is not None else _default_devices(len(shards))
        if len(devs) != len(shards):
            raise RuntimeError(f"{index_dir} holds {len(shards)} shard files, {len(devs)} devices were given")
        index = faiss.read_index_shards(index_dir, devs)
        if (index_dir / "paths.json").exists():
            paths = json.loads((index_dir / "paths.json").read_text(encoding="utf-8"))
            try:
                meta = json.loads((index_dir / "meta.json").read_text(encoding="utf-8"))
            except Exception:
                meta = {}
        else:
            lines = (index_dir / "ids.txt").read_text(encoding="utf-8").splitlines()
            paths = [ln.strip() for ln in lines if ln.strip()]
            meta = {"dim": index.d}
        return index, paths, meta
    if new_idx.exists() and (index_dir / "paths.json").exists():
        index = faiss.read_index(str(new_idx))
        paths = json.loads((index_dir / "paths.json").read_text(encoding="utf-8"))
        try:
            meta = json.loads((index_dir / "meta.json").read_text(encoding="utf-8"))
        except Exception:
            meta = {}
        return index, paths, meta
    if old_idx.exists() and (index_dir / "ids.txt").exists():
        index = faiss.read_index(str(old_idx))
        lines = (index_dir / "ids.txt").read_text(encoding="utf-8").splitlines()
        paths = [ln.strip() for ln in lines if ln.strip()]
        meta = {"dim": index.d}
        try:
            if paths:
                _, header = unpack_c2df(paths[0])
                if isinstance(header, dict) and header.get("model_id"):
                    meta["model_id"] = header["model_id"]
        except Exception:
            pass
        return index, paths, meta
    raise FileNotFoundError(f"Can't find FAISS index in {index_dir}")


def encode_c2df_query(c2df_path) -> np.ndarray:
    """Query embedding (1, d) fp32 from a bitstream — src/search.py:107-109."""
    z, _ = decode_clip_from_c2df(c2df_path)
    return z[None, :].astype("float32")


def do_search(q: np.ndarray, index: faiss.Index, paths: List[str], topk: int = 10) -> List[Tuple[str, float]]:
    """Top-k (path, score) pairs for the first query row — src/search.py:113-120."""
    k = max(1, min(topk, index.ntotal))
    sim, ids = index.search(q, k)
    return [(paths[i], float(sim[0, j])) for j, i in enumerate(ids[0]) if i != -1]


def do_search_batch(q: np.ndarray, index: faiss.Index, paths: List[str], topk: int = 10):
    """Additive: the same for every query row (the reference only consumes row 0)."""
    k = max(1, min(topk, index.ntotal))
    sim, ids = index.search(q, k)
    return [[(paths[i], float(sim[r, j])) for j, i in enumerate(ids[r]) if i != -1] for r in range(ids.shape[0])]


def _load_vec(path: Path, d: int) -> np.ndarray:
    v = np.load(path).astype("float32").reshape(1, -1)
    if v.shape[1] != d:
        raise ValueError(f"{path}: embedding has {v.shape[1]} dims, index has {d}")
    return l2n(v).astype("float32")


def resolve_clip_dir(arg, meta) -> Optional[Path]:
    """Where the CLIP checkpoint of a text / image query comes from: ``--clip_dir``, else ``$SGIC_CLIP_DIR``, else
    ``meta["model_id"]`` when that names a local directory (the reference picks its model from ``meta["model_id"]``,
    src/search.py:151-152 — an OpenCLIP name there, which needs a download).  With the environment variable set, the
    reference's own command line (``query-text --index_dir D --text T --topk K``, webapp.py:246-248) works unchanged."""
    for cand in (arg, os.environ.get("SGIC_CLIP_DIR"), (meta or {}).get("model_id")):
        if cand and Path(str(cand)).is_dir():
            return Path(str(cand))
    return None


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description="query-text / query-image / query-c2df")
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name, flag, kind, helptext in (("query-text", "--text", str, "searching with text"),
                                       ("query-image", "--image", Path, "searching with image"),
                                       ("query-c2df", "--c2df", Path, "searching with .c2df")):
        sp = sub.add_parser(name, help=helptext)
        sp.add_argument("--index_dir", type=Path, required=True)
        sp.add_argument(flag, type=kind, required=True)
        sp.add_argument("--topk", type=int, default=10)
        if name != "query-c2df":
            sp.add_argument("--vec", type=Path, default=None,
                            help="pre-computed CLIP embedding (.npy)")
            sp.add_argument("--clip_dir", type=Path, default=None,
                            help="local CLIP checkpoint directory (Hugging Face layout): the query is embedded on the "
                                 "index's GPU and searched without leaving it (query_encoders.ClipQueryEncoder)")
    args = ap.parse_args(argv)
    try:
        index, paths, _meta = load_index(args.index_dir)
        clip_dir = resolve_clip_dir(getattr(args, "clip_dir", None), _meta) if args.cmd != "query-c2df" else None
        if args.cmd == "query-c2df":
            q = encode_c2df_query(args.c2df)
        elif args.cmd in ("query-text", "query-image") and args.vec is None and clip_dir is not None:
            from .query_encoders import ClipQueryEncoder
            enc = ClipQueryEncoder(clip_dir, device=index.device)
            if args.cmd == "query-text":
                z = enc.encode_text([args.text])
            else:
                from PIL import Image
                z = enc.encode_image([Image.open(args.image).convert("RGB")])
            results = enc.search(index, z, paths, topk=args.topk)[0]
            print(json.dumps([{"path": p, "score": s} for p, s in results], ensure_ascii=False, indent=2))
            return
        elif args.cmd in ("query-text", "query-image"):
            if args.vec is None:
                raise NotImplementedError(
                    f"{args.cmd}: no CLIP weights can be fetched offline; pass a local checkpoint with --clip_dir DIR "
                    "(or $SGIC_CLIP_DIR) or the embedding with --vec file.npy")
            q = _load_vec(args.vec, index.d)
        else:
            raise ValueError(f"Unknown behavior: {args.cmd}")
        results = do_search(q, index, paths, topk=args.topk)
        print(json.dumps([{"path": p, "score": s} for p, s in results], ensure_ascii=False, indent=2))
    except Exception as e:
        print(f"[ERROR] {e}")
        traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
