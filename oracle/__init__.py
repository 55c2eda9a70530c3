"""CPU oracle for the retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``searchable-generative-image-compression_b200``) never does: it fails loudly
when its CUDA library is missing instead of falling back to anything here.

Parity pinning: the arithmetic of the path lives in ``faiss-cpu``
(unpinned in the reference's requirements.txt:29, not vendored, not installable
here), and the reference ships no tests.  The *formats* and one 1-row index are
pinned by the shipped ``IO/*/apple.*`` fixtures (copied to ``tests/golden/``) and
by vectors generated from the reference's own ``src/filemaker.py`` (importable
here; see ``tests/golden/make_golden.py``).  The flat-IP *search* results are
"parity unpinned" at the FAISS boundary: the oracle restates FAISS's published
algorithm and is cross-checked against fp64 brute force.
"""
