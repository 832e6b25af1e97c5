"""CPU oracle for the AH-RAG retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or the reported CPU baseline.
The product path (``ah_rag_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Parity status
-------------
* graph half (anchor expansion + rerank, expand_parents/children/related, gym step):
  pinned -- ``oracle/ref_search.py`` is checked against the UNMODIFIED reference code
  (``/root/reference/src/ah_rag``) run here through ``oracle/stubs.py``; golden vectors
  produced by that run are committed under ``tests/golden/`` by ``oracle/make_golden.py``.
* similarity half (``Collection.query`` at hierarchical_graph.py:714): PARITY UNPINNED.
  The arithmetic lives in chromadb / chroma-hnswlib, which is neither vendored nor
  version-pinned by the reference (requirements.txt:1) and is not installable here.
  ``oracle/exact.py`` restates the published metric of the ``hnsw:space = cosine``
  collection (distance = 1 - <q, x> on L2-normalised fp32 rows) as an exact brute-force
  scan with a canonical, order-fixed fp32 score and ties broken by lower row id.
"""
