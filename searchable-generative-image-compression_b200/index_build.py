"""Index construction — host-side mirror of the reference's ``src/build.py``
(``build_index_from_c2df_dir`` :71-103) and of the retrieval pieces of ``src/compress.py``
(``ClipCodec.quantize_u8_and_compress`` :76-86, ``FaissDB`` :89-114).

``build_index_from_c2df_dir`` keeps the reference's observable behaviour — sorted recursive
glob, per-file ``[SKIP]`` on any decode error, both naming schemes written
(``faiss.index`` + ``paths.json`` + ``meta.json`` and ``index.faiss`` + ``ids.txt``) — but the
per-file Python loop is replaced by the batched C++ parser + device-side loader
(``IndexFlatIP.add_c2df``).  ``build_index_from_image_dir`` (:207-241, the ``build-images``
command) embeds an image directory with the CLIP encoder of ``query_encoders`` — a local
checkpoint directory instead of an OpenCLIP download — and writes the same four files; the
``download`` command and ``auto_download`` need the network and are out of scope.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import traceback
from pathlib import Path
from typing import List

import numpy as np

from . import _native
from . import faiss_compat as faiss
from . import zstd
from .c2df import unpack_c2df
from .retrieval import decode_clip_from_c2df, load_index  # build.py carries its own copies (:26-43, :106-126)

__all__ = ["build_index_from_c2df_dir", "build_index_from_image_dir", "list_images", "encode_images_in_batches",
           "quantize_u8_and_compress", "ClipCodec", "FaissDB", "load_index", "from_npy_dir", "pack_npy_dir"]


def quantize_u8_and_compress(z_unit: np.ndarray, model_id: str = "ViT-B-32:laion2b_s34b_b79k"):
    """fp32 unit vector → (zstd-19 bytes of the u8 codes, clip_meta dict) — src/compress.py:76-86.
    ``np.round`` is round-half-to-even, as in the reference."""
    q = np.clip(np.round((z_unit * 0.5 + 0.5) * 255.0), 0, 255).astype(np.uint8)
    meta = {"model_id": model_id, "dim": int(z_unit.shape[0]), "quant": "u8_symmetric_-1_1",
            "codec": "zstd", "zstd_level": 19}
    return zstd.compress(q.tobytes(), 19), meta


class ClipCodec:
    """The retrieval half of src/compress.py:57-86: image → unit CLIP vector → u8 codes + zstd → ``clip_stream`` /
    ``clip_meta`` of a ``.c2df``.  ``encoder`` is a :class:`~.query_encoders.ClipQueryEncoder` (or a local checkpoint
    directory for one); ``model_name`` is what ``clip_meta["model_id"]`` records (the reference writes
    ``"<arch>:<pretrained>"``)."""

    def __init__(self, encoder, model_name: str = "ViT-B-32:laion2b_s34b_b79k", device=0):
        if not hasattr(encoder, "encode_image"):
            from .query_encoders import ClipQueryEncoder
            encoder = ClipQueryEncoder(encoder, device=_device_index(device))
        self.encoder = encoder
        self.device = encoder.device
        self.model_name = model_name

    def image_to_unit_vec(self, img_tensor_CHW) -> np.ndarray:
        """(3, H, W) tensor in [-1, 1] → (D,) fp32 unit vector on the host (compress.py:67-74: through an 8-bit PIL
        image, then the CLIP preprocessing)."""
        from PIL import Image
        x = img_tensor_CHW.detach().clamp(-1, 1).mul(0.5).add(0.5)
        pil = Image.fromarray(x.mul(255).byte().permute(1, 2, 0).cpu().numpy())
        return self.encoder.encode_image([pil]).squeeze(0).cpu().numpy().astype("float32")

    def quantize_u8_and_compress(self, z_unit: np.ndarray):
        return quantize_u8_and_compress(z_unit, self.model_name)


def _skip_reason(path: Path, code: int) -> str:
    """Message of the exception the reference would have printed for this file."""
    if code in (3, 4, 6):
        return f"{path} {_native.C2DF_STATUS[code]}"
    return _native.C2DF_STATUS.get(code, f"status {code}")


def _header_model_id(path):
    """``header.get("model_id")`` of one ``.c2df`` reading only its JSON header (filemaker.py:146-150), not the
    entries: the sniff stays cheap when no file of a large corpus names a model."""
    import struct
    try:
        with open(path, "rb") as f:
            head = f.read(10)
            if len(head) < 10 or head[:4] != b"C2DF":
                return None
            (hlen,) = struct.unpack_from("<I", head, 6)
            header = json.loads(f.read(hlen).decode("utf-8"))
        return header.get("model_id") if isinstance(header, dict) else None
    except Exception:
        return None


def build_index_from_c2df_dir(c2df_dir, index_dir, *, dtype="fp16", device=None, n_threads: int = 0) -> None:
    c2df_dir, index_dir = Path(c2df_dir), Path(index_dir)
    index_dir.mkdir(parents=True, exist_ok=True)
    paths = sorted(c2df_dir.glob("**/*.c2df"))
    if not paths:
        raise RuntimeError(f"Empty folder: {c2df_dir}")

    # the index dimension is that of the first decodable file (X.shape[1] upstream, build.py:92)
    d = None
    for p in paths:
        try:
            z, _ = decode_clip_from_c2df(p)
            d = int(z.shape[0])
            break
        except Exception:
            continue
    if d is None:
        for p in paths:
            print(f"[SKIP] {p.name}: not a usable .c2df")
        raise RuntimeError("No available .c2df")
    # retain_codes: the u8 codes stay on the host (1 byte per element) so that the two IxFI files below hold the
    # fp32 rows of dequantize_clip_u8 exactly as build.py:82-99 writes them, not fp16-rounded values
    index = faiss.IndexFlatIP(d, dtype=dtype, device=device, retain_fp32=False, retain_codes=True)
    statuses = index.add_c2df_paths(paths, n_threads=n_threads)
    if any(int(st) == 7 for st in statuses):
        # a decodable file of another dimension: the reference collects it and then fails as a whole in
        # np.concatenate (build.py:91) — same here, nothing is written
        bad = next(p for p, st in zip(paths, statuses) if int(st) == 7)
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly, "
                         f"but along dimension 1, the first array has size {d} and {bad.name} has another size")
    keep: List[str] = []
    model_id = None
    for p, st in zip(paths, statuses):
        if st == 0:
            keep.append(str(p))
            if model_id is None:      # build.py:85-86: the first kept header that names a model
                model_id = _header_model_id(p)
        else:
            print(f"[SKIP] {p.name}: {_skip_reason(p, int(st))}")
    if not keep:
        raise RuntimeError("No available .c2df")

    faiss.write_index(index, str(index_dir / "faiss.index"))
    (index_dir / "paths.json").write_text(json.dumps(keep, ensure_ascii=False, indent=2), encoding="utf-8")
    (index_dir / "meta.json").write_text(json.dumps({"dim": d, "model_id": model_id}, ensure_ascii=False, indent=2),
                                         encoding="utf-8")
    faiss.write_index(index, str(index_dir / "index.faiss"))
    (index_dir / "ids.txt").write_text("\n".join(keep), encoding="utf-8")
    print(f"[OK] Index process completed!: N={index.ntotal}, dim={d}")
    if model_id:
        print(f"[INFO] Suggested CLIP model: {model_id}")


def _device_index(device) -> int:
    """The reference passes torch device strings ("cuda", "cuda:1"); the index wants a GPU number."""
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    name = str(device)
    if name.startswith("cuda"):
        return int(name.split(":", 1)[1]) if ":" in name else 0
    raise RuntimeError(f"device {device!r}: the index lives in GPU memory (use \"cuda\" or \"cuda:N\")")


def list_images(root, exts=None) -> List[Path]:
    """Image files under ``root`` in ``rglob`` order (NOT sorted — src/build.py:171-180), suffix match is
    case-insensitive and a missing leading dot in ``exts`` is added."""
    if exts is None:
        exts = {".jpg", ".jpeg", ".png", ".webp", ".bmp"}
    exts = {e if e.startswith(".") else "." + e for e in {str(e).lower() for e in exts}}
    return [p for p in Path(root).rglob("*") if p.is_file() and p.suffix.lower() in exts]


def encode_images_in_batches(img_paths, encoder, batch_size: int = 32):
    """src/build.py:182-205: unreadable images are reported with ``[SKIP]`` and left out, the rest go through the
    image tower ``batch_size`` at a time; returns the fp32 unit rows on the host (``None`` if no image could be read).
    ``encoder``: a :class:`~.query_encoders.ClipQueryEncoder` (the embedding itself runs on its device)."""
    from PIL import Image
    feats, cur = [], []

    def flush():
        if cur:
            feats.append(encoder.encode_image(list(cur)).cpu().numpy().astype("float32"))
            cur.clear()

    for pth in img_paths:
        try:
            im = Image.open(pth).convert("RGB")
        except Exception as e:
            print(f"[SKIP] Can't read the image {pth}: {e}")
            continue
        cur.append(im)
        if len(cur) == batch_size:
            flush()
    flush()
    return np.concatenate(feats, 0).astype("float32") if feats else None


def build_index_from_image_dir(image_dir, index_dir, model_id, device=None, batch_size: int = 32, exts=None,
                               limit=None, random_pick: bool = False, seed=None, desired=None,
                               auto_download: bool = False, *, encoder=None, **index_kwargs) -> None:
    """``build-images`` (src/build.py:207-241): embed every image under ``image_dir`` and write ``faiss.index`` +
    ``paths.json`` + ``meta.json`` and ``index.faiss`` + ``ids.txt``.

    ``model_id`` is what ``meta.json`` records; the network comes from ``encoder`` (a ``ClipQueryEncoder``) or, when
    that is not given, from ``model_id`` read as a local checkpoint directory — OpenCLIP's ``arch:pretrained`` names
    need a download.  Selection follows the reference: ``desired`` (if positive) else ``limit`` images, the first ones
    in listing order or ``random.Random(seed).sample`` with ``random_pick``.  As in the reference the path lists name
    every SELECTED image, also one whose pixels could not be read (build.py:236).  The fp32 rows are kept on the host
    as well, so the interchange files hold exactly the encoder's vectors."""
    import random
    image_dir, index_dir = Path(image_dir), Path(index_dir)
    index_dir.mkdir(parents=True, exist_ok=True)
    all_imgs = list_images(image_dir, exts=exts)
    if desired is not None and auto_download and len(all_imgs) < desired:
        raise RuntimeError(f"{len(all_imgs)} images under {image_dir}, {desired} wanted: downloading the shortfall "
                           "(build.py:138-169) needs the network and is not part of this package")
    if not all_imgs:
        raise RuntimeError(f"There is no image in {image_dir}")
    target_n = desired if (desired is not None and desired > 0) else limit
    if target_n is not None and 0 < target_n <= len(all_imgs):
        all_imgs = random.Random(seed).sample(all_imgs, target_n) if random_pick else all_imgs[:target_n]
    print(f"[INFO] Using {len(all_imgs)} images to build the index")
    if encoder is None:
        from .query_encoders import ClipQueryEncoder
        if model_id is None or not Path(str(model_id)).is_dir():
            raise RuntimeError(f"model_id {model_id!r} is not a local CLIP checkpoint directory (OpenCLIP names need a "
                               "download); pass encoder=ClipQueryEncoder(...)")
        encoder = ClipQueryEncoder(model_id, device=_device_index(device))
    X = encode_images_in_batches(all_imgs, encoder, batch_size=batch_size)
    if X is None:
        raise RuntimeError("Failed to build FAISS index!")
    d = X.shape[1]
    if device is not None and "device" not in index_kwargs and "devices" not in index_kwargs:
        index_kwargs["device"] = _device_index(device)
    index = faiss.IndexFlatIP(d, **index_kwargs)
    index.add(X)
    paths = [str(p) for p in all_imgs]
    faiss.write_index(index, str(index_dir / "faiss.index"))
    (index_dir / "paths.json").write_text(json.dumps(paths, ensure_ascii=False, indent=2), encoding="utf-8")
    (index_dir / "meta.json").write_text(json.dumps({"dim": d, "model_id": model_id}, ensure_ascii=False, indent=2),
                                         encoding="utf-8")
    faiss.write_index(index, str(index_dir / "index.faiss"))
    (index_dir / "ids.txt").write_text("\n".join(paths), encoding="utf-8")
    print(f"[OK] Index process completed!: N={index.ntotal}, dim={d}")
    index.close()


class FaissDB:
    """Open-or-create ``index.faiss`` + ``ids.txt`` and append — src/compress.py:89-114.

    ``add`` renormalises with ``v / (||v|| + 1e-12)`` exactly as the reference does, and
    ``persist`` writes one id per line with a trailing newline.
    """

    def __init__(self, index_dir: str, dim: int, **index_kwargs):
        os.makedirs(index_dir, exist_ok=True)
        self.index_path = os.path.join(index_dir, "index.faiss")
        self.ids_path = os.path.join(index_dir, "ids.txt")
        if os.path.exists(self.index_path):
            self.index = faiss.read_index(self.index_path, **index_kwargs)
        else:
            self.index = faiss.IndexFlatIP(dim, **index_kwargs)
        self.ids: List[str] = []
        if os.path.exists(self.ids_path):
            with open(self.ids_path, "r", encoding="utf-8") as f:
                self.ids = [ln.strip() for ln in f if ln.strip()]

    def add(self, vec_unit: np.ndarray, doc_id: str) -> None:
        assert vec_unit.ndim == 1
        v = vec_unit.copy()[None, :]
        v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
        self.index.add(v.astype("float32"))
        self.ids.append(doc_id)

    def add_many(self, vecs: np.ndarray, doc_ids: List[str]) -> None:
        """Additive: the same renormalisation for a whole (n, d) block in one add."""
        v = np.array(vecs, dtype=np.float32, copy=True)
        v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-12
        self.index.add(v.astype("float32"))
        self.ids.extend(doc_ids)

    def persist(self) -> None:
        faiss.write_index(self.index, self.index_path)
        with open(self.ids_path, "w", encoding="utf-8") as f:
            for _id in self.ids:
                f.write(_id + "\n")


PACKED_SUBDIR = "_packed"      # clip_vecs/_packed/{vecs.npy, stems.txt}: not matched by the reference's glob("*.npy")


def pack_npy_dir(clip_vecs_dir) -> Path:
    """On-disk v2 of ``IO/clip_vecs``: ONE ``(N, d)`` fp32 ``.npy`` + the stems, in the order
    ``sorted(glob("*.npy"))`` gives (src/compress.py:296), instead of one 2 KB file per vector (:286) that the
    index build re-opens one by one (:300-305).  Written next to the per-vector files; they stay valid."""
    clip_vecs_dir = Path(clip_vecs_dir)
    files = sorted(clip_vecs_dir.glob("*.npy"))
    if not files:
        raise RuntimeError(f"Empty folder: {clip_vecs_dir}")
    d = int(np.load(files[0]).reshape(-1).shape[0])
    out = clip_vecs_dir / PACKED_SUBDIR
    out.mkdir(exist_ok=True)
    mm = np.lib.format.open_memmap(out / "vecs.npy.tmp", mode="w+", dtype=np.float32, shape=(len(files), d))
    for i, f in enumerate(files):
        mm[i] = np.load(f).astype(np.float32).reshape(-1)
    mm.flush()
    del mm
    (out / "stems.txt").write_text("".join(f.stem + "\n" for f in files), encoding="utf-8")
    os.replace(out / "vecs.npy.tmp", out / "vecs.npy")
    return out / "vecs.npy"


def _packed_vectors(clip_vecs_dir: Path):
    """(memory-mapped (N, d) array, stems) when a packed copy exists and covers the per-vector files; else None."""
    vecs, stems = clip_vecs_dir / PACKED_SUBDIR / "vecs.npy", clip_vecs_dir / PACKED_SUBDIR / "stems.txt"
    if not (vecs.exists() and stems.exists()):
        return None
    names = [ln for ln in stems.read_text(encoding="utf-8").splitlines() if ln]
    arr = np.load(vecs, mmap_mode="r")
    if arr.ndim != 2 or arr.shape[0] != len(names):
        return None
    loose = sorted(p.stem for p in clip_vecs_dir.glob("*.npy"))
    if loose and loose != names:            # per-vector files were added / removed since: the packed copy is stale
        return None
    return arr, names


def from_npy_dir(clip_vecs_dir, bitstream_dir, index_dir, *, require_bitstream: bool = True, **index_kwargs) -> FaissDB:
    """The rank-0 tail of ``compress.test`` (src/compress.py:295-306): every ``clip_vecs/<stem>.npy`` in sorted
    order whose ``<bitstream_dir>/<stem>.c2df`` exists (:304) → ``FaissDB.add`` with that id → ``persist``.
    Vectors are appended block-wise; when ``clip_vecs/_packed`` holds an up-to-date packed copy
    (:func:`pack_npy_dir`) the rows come out of that one file instead of N small ones."""
    clip_vecs_dir = Path(clip_vecs_dir)
    packed = _packed_vectors(clip_vecs_dir)
    if packed is not None:
        arr, stems = packed
        source = ((stems[i], arr[i]) for i in range(len(stems)))
        d = int(arr.shape[1])
    else:
        files = sorted(clip_vecs_dir.glob("*.npy"))
        if not files:
            raise RuntimeError(f"Empty folder: {clip_vecs_dir}")
        d = int(np.load(files[0]).shape[-1])
        source = ((f.stem, np.load(f)) for f in files)
    db = FaissDB(str(index_dir), d, **index_kwargs)
    block, ids = [], []
    for stem, vec in source:
        doc_id = os.path.join(str(bitstream_dir), stem + ".c2df")
        if require_bitstream and not os.path.exists(doc_id):
            continue
        block.append(np.asarray(vec, dtype=np.float32).reshape(-1))
        ids.append(doc_id)
        if len(block) >= 65536:
            db.add_many(np.stack(block), ids)
            block, ids = [], []
    if block:
        db.add_many(np.stack(block), ids)
    db.persist()
    return db


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description="build (from .c2df bitstreams)")
    sub = ap.add_subparsers(dest="cmd", required=True)
    sp = sub.add_parser("build", help="build the index from a directory of .c2df")
    sp.add_argument("--c2df_dir", type=Path, required=True)
    sp.add_argument("--index_dir", type=Path, required=True)
    si = sub.add_parser("build-images", help="build the index from an image directory (CLIP checkpoint: a local directory)")
    si.add_argument("--image_dir", type=Path, required=True)
    si.add_argument("--index_dir", type=Path, required=True)
    si.add_argument("--model_id", type=str, default=None, help="local CLIP checkpoint directory (Hugging Face layout)")
    si.add_argument("--device", type=int, default=0)
    si.add_argument("--batch_size", type=int, default=32)
    si.add_argument("--exts", type=str, nargs="*", default=None)
    si.add_argument("--limit", type=int, default=None)
    si.add_argument("--random_pick", action="store_true")
    si.add_argument("--seed", type=int, default=None)
    si.add_argument("--desired", type=int, default=None)
    args = ap.parse_args(argv)
    try:
        if args.cmd == "build-images":
            build_index_from_image_dir(args.image_dir, args.index_dir, args.model_id, args.device, args.batch_size,
                                       set(args.exts) if args.exts else None, args.limit, args.random_pick, args.seed,
                                       args.desired)
        else:
            build_index_from_c2df_dir(args.c2df_dir, args.index_dir)
    except Exception as e:
        print(f"[ERROR] {e}")
        traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
