"""CPU oracle for the A-NICE-RAG retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  The only permitted importers
are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker or
the timed CPU baseline, never as something the product path routes through.
The product (``a-nice-rag_b200/``) fails loudly when its CUDA library is
missing; it never falls back to this code.

What is restated here and what pins it
--------------------------------------
* ``oracle.retrieval``  -- numpy restatement of the reference's
  ``src/search_engine.py`` (dense inner-product top-k :57-98 / :100-146, source
  prefix filter :36-55, BM25 top-k :205-243, weighted RRF :21-34).
  PINNED: validated in this container against the reference's own module
  imported verbatim (``oracle.reference_loader``) and against the golden
  vectors that import produced (``tests/golden/*.npz``, generator
  ``oracle/make_golden.py``).
* ``oracle.bm25_okapi`` -- restatement of ``rank_bm25.BM25Okapi`` (PyPI
  ``rank-bm25``; the reference leaves it unpinned in ``requirements.txt:5``, the
  release any modern install resolves to is 0.2.2).  The package source is NOT
  under ``/root/reference`` and is not installable offline, and the reference
  has no test or golden vector at this boundary, so for the BM25 *arithmetic*:
  **parity unpinned** -- the restatement follows the published 0.2.2 algorithm
  and is anchored only on the reference's call sites
  (``src/processing/bm25_search.py:77``, ``src/search_engine.py:219``).
* ``oracle.csr``        -- vectorised numpy CSR form of the same BM25 scoring,
  proven equal to the literal restatement in ``tests/test_oracle.py``; used
  where the literal per-document Python loop is infeasible (>= 1M documents).
"""
