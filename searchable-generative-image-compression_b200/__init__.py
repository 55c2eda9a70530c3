"""B200-native exact inner-product k-NN over the CLIP embeddings of ``.c2df`` bitstreams.

Drop-in for the retrieval hot path of lionl1106/Searchable-Generative-Image-Compression
(the FAISS ``IndexFlatIP`` path behind ``src/search.py``).  Import as ``sgic_b200``:

    from sgic_b200 import faiss_compat as faiss      # IndexFlatIP / read_index / write_index
    from sgic_b200.retrieval import load_index, do_search, encode_c2df_query
    from sgic_b200.index_build import build_index_from_c2df_dir, FaissDB

The arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``libsgic.so`` (``include/sgic.h``); there is no CPU or PyTorch fallback.
"""
from . import _native  # noqa: F401  (does not load the library until first use)
from . import c2df, zstd  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # lazy: these import the native library on first touch
    if name in ("faiss_compat", "retrieval", "index_build", "sharded", "synth"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
