

class Array(object):
    """placeholder type: scipy/array-api-compat probes `dask.array.Array` with issubclass once `dask.array` is in sys.modules."""
