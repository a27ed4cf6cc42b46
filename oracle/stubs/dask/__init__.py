"""Empty stand-in so signal/index_dataset.py:3 (`import dask.array as da`) imports; lazy mode unused."""
