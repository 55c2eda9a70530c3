"""Synthetic embeddings (SURVEY.md §8d "Synthetic inputs") generated on the device with
torch — data plumbing for tests and ``bench.py``, not part of the search path."""
from __future__ import annotations

import numpy as np


def fill_index_random(index, n_rows: int, *, seed: int = 1234, chunk_rows: int = 1 << 20, cone: float = 0.0,
                      row0: int = 0):
    """Append ``n_rows`` L2-normalised Gaussian rows (optionally pulled towards one shared
    direction: the "CLIP-cone" distribution) to ``index``.  Row r depends only on
    (seed, row0 + r)'s chunk, so shards of one logical database can be filled independently
    as long as ``row0`` and ``chunk_rows`` line up."""
    import torch
    dev = torch.device("cuda", index.device)
    d = index.d
    want = torch.bfloat16 if index.dtype == "bf16" else torch.float16
    index.reserve(index.ntotal + n_rows)
    g = torch.Generator(device=dev)
    center = None
    if cone:
        gc = torch.Generator(device=dev)
        gc.manual_seed(seed ^ 0x5eed)
        center = torch.randn(d, generator=gc, device=dev)
        center /= center.norm()
    assert row0 % chunk_rows == 0
    done = 0
    while done < n_rows:
        rows = min(chunk_rows, n_rows - done)
        g.manual_seed(seed + (row0 + done) // chunk_rows)
        x = torch.randn((chunk_rows if rows == chunk_rows else rows, d), generator=g, device=dev)
        if center is not None:
            x = x / x.norm(dim=1, keepdim=True) + cone * center
        x = x / x.norm(dim=1, keepdim=True)
        index.add_torch(x.to(want).contiguous())
        done += rows
    torch.cuda.synchronize(dev)


def random_unit_queries(nq: int, d: int, seed: int = 4321) -> np.ndarray:
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q
