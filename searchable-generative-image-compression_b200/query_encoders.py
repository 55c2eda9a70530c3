"""CLIP query encoders on the index's device — the step in front of the hot path.

The reference embeds a text or image query with OpenCLIP on the GPU, brings the vector back to the host and hands
it to the CPU index (src/search.py:48-63 ``load_clip``, :93-97 ``encode_text``, :100-105 ``encode_image``, :162
``do_search``).  Here the embedding never leaves the device: the encoder runs on the index's (home) GPU and stream
and its L2-normalised fp32 output goes straight into ``index.search_torch`` — and a batch of queries is one encoder
pass and one search.

OpenCLIP and its weights are not installable offline, and the default checkpoint of the reference
(``ViT-B-32:laion2b_s34b_b79k``) cannot be downloaded here, so this module takes the model from a LOCAL directory in
the Hugging Face CLIP layout (``transformers.CLIPModel.from_pretrained(dir)`` + its tokenizer / image processor;
``laion/CLIP-ViT-B-32-laion2B-s34B-b79K`` is the same network as the reference's default) or an already constructed
model object.  Without a checkpoint ``retrieval``'s CLI keeps asking for ``--vec``.  Tests run the plumbing with a
small randomly initialised CLIP (no download).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["ClipQueryEncoder"]


def _features(out):
    """``get_text_features`` / ``get_image_features`` return the projected embedding as a tensor (older transformers)
    or as ``pooler_output`` of an output object (newer ones)."""
    return out if hasattr(out, "shape") else out.pooler_output


class ClipQueryEncoder:
    """``load_clip`` + ``encode_text`` + ``encode_image`` of src/search.py on the index's GPU.

    ``model``: a directory with a Hugging Face CLIP checkpoint, or a ``transformers.CLIPModel`` instance.
    ``tokenizer`` / ``image_processor``: taken from the directory when not given (only needed for raw strings /
    PIL images; pre-tokenised ids and pixel tensors need neither).
    """

    def __init__(self, model, device: int = 0, *, tokenizer=None, image_processor=None, dtype=None):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", int(device)) if torch.cuda.is_available() else torch.device("cpu")
        if isinstance(model, (str, bytes)) or hasattr(model, "__fspath__"):
            from transformers import CLIPModel
            path = str(model)
            model_obj = CLIPModel.from_pretrained(path)
            if tokenizer is None:
                try:
                    from transformers import CLIPTokenizerFast
                    tokenizer = CLIPTokenizerFast.from_pretrained(path)
                except Exception:
                    tokenizer = None
            if image_processor is None:
                try:
                    from transformers import CLIPImageProcessor
                    image_processor = CLIPImageProcessor.from_pretrained(path)
                except Exception:
                    image_processor = None
            model = model_obj
        self.model = model.to(self.device).eval()
        if dtype is not None:
            self.model = self.model.to(dtype)
        self.tokenizer, self.image_processor = tokenizer, image_processor

    @property
    def dim(self) -> int:
        return int(self.model.config.projection_dim)

    # ---------------------------------------------------------------- encoders (device tensors in, device tensors out)
    def encode_text_ids(self, input_ids, attention_mask=None):
        """(n, seq) token ids → (n, d) fp32 unit vectors on the device (src/search.py:95-96)."""
        torch = self.torch
        with torch.no_grad():
            z = _features(self.model.get_text_features(input_ids=input_ids.to(self.device),
                                                       attention_mask=None if attention_mask is None
                                                       else attention_mask.to(self.device))).float()
            return (z / z.norm(dim=-1, keepdim=True)).contiguous()

    def encode_pixels(self, pixel_values):
        """(n, 3, H, W) preprocessed images → (n, d) fp32 unit vectors on the device (src/search.py:103-104)."""
        torch = self.torch
        with torch.no_grad():
            z = _features(self.model.get_image_features(pixel_values=pixel_values.to(self.device))).float()
            return (z / z.norm(dim=-1, keepdim=True)).contiguous()

    def encode_text(self, texts: Sequence[str]):
        if self.tokenizer is None:
            raise RuntimeError("no tokenizer: pass token ids to encode_text_ids, or a checkpoint directory with a vocabulary")
        tok = self.tokenizer(list(texts), padding="max_length", truncation=True, return_tensors="pt")
        return self.encode_text_ids(tok["input_ids"], tok.get("attention_mask"))

    def encode_image(self, images):
        if self.image_processor is None:
            raise RuntimeError("no image processor: pass pixel tensors to encode_pixels")
        px = self.image_processor(images=images, return_tensors="pt")["pixel_values"]
        return self.encode_pixels(px)

    # ---------------------------------------------------------------- query → top-k without a host round trip
    def search(self, index, z, paths: Optional[List[str]] = None, topk: int = 10):
        """``do_search`` (src/search.py:113-120) for every row of the device-resident embeddings ``z``: the vectors
        go from the encoder's output buffer into the scan kernels on the same stream.  Returns ``(D, I)`` device
        tensors, or ``[[(path, score), …], …]`` when ``paths`` is given."""
        torch = self.torch
        if z.shape[1] != index.d:
            raise ValueError(f"the encoder gives {z.shape[1]} dims, the index has {index.d}")
        k = max(1, min(int(topk), index.ntotal))
        if z.device.index != index.device:
            z = z.to(torch.device("cuda", index.device))
        D, I = index.search_torch(z.contiguous(), k)
        if paths is None:
            return D, I
        Dh, Ih = D.cpu().numpy(), I.cpu().numpy()
        return [[(paths[i], float(Dh[r, j])) for j, i in enumerate(Ih[r]) if i != -1] for r in range(Ih.shape[0])]

    def search_text(self, index, texts: Sequence[str], paths=None, topk: int = 10):
        return self.search(index, self.encode_text(texts), paths, topk)

    def search_image(self, index, images, paths=None, topk: int = 10):
        return self.search(index, self.encode_image(images), paths, topk)
