"""Stand-in ``torch_geometric`` built from oracle.pyg so that the UNMODIFIED reference modules can be
imported in the build container (PyG is not installable here).  Test infrastructure only; used by
oracle/refload.py, tests/golden/make_goldens.py and the in-container oracle-vs-reference tests."""
