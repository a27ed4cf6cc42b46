"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the *checker* (or as the
timed CPU baseline).  The product package ``pytorch_geometric_temporal_b200`` never
imports it and fails loudly when its CUDA library is missing.

What it restates
----------------
* ``oracle.pyg``        -- the third-party ``torch_geometric`` primitives the reference calls
                           (un-pinned dependency, ``/root/reference/setup.py:7``; source absent from
                           ``/root/reference``).  Restated from PyG's published semantics
                           (SURVEY.md section 8c / Appendix A).
* ``oracle.recurrent``  -- ``nn/recurrent/{dcrnn,gconv_gru,gconv_lstm,temporalgcn,attentiontemporalgcn}.py``
* ``oracle.attention``  -- ``nn/attention/astgcn.py``
* ``oracle.signal``     -- ``signal/{static_graph_temporal_signal,index_dataset,train_test_split}.py``

Pinning status
--------------
* Indexing (snapshot iterator vs index-batching): pinned bit-exactly against the reference's own
  test (``test/index_test.py:93-114``) re-run here on the in-tree chickenpox fixture.
* Module logic (everything in ``oracle.recurrent`` / ``oracle.attention``): pinned against the
  UNMODIFIED reference modules imported in this container on top of ``oracle/stubs`` (a stand-in
  ``torch_geometric`` built from ``oracle.pyg``); vectors committed under ``tests/golden``
  by ``tests/golden/make_goldens.py``.
* PyG primitive floating-point semantics: **parity unpinned** -- the reference ships no numeric
  golden for them (its layer tests assert shapes only, ``test/recurrent_test.py:98-111``) and PyG
  cannot be installed here.  They are validated against closed-form dense-matrix formulas instead
  (``tests/test_oracle_kat.py``).
"""
