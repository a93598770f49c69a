"""Dense-backed stand-in for pydata `sparse.COO` -- enough for the reference's dispatchers.
The dense numba branch is the oracle of record; this only keeps `import sparse` and the
nnz-ratio test in colosseum/dynamic_programming/infinite_horizon.py:25-33 working."""
import numpy as np


class COO:
    __array_priority__ = 1000
    def __init__(self, coords, data=None, shape=None):
        if data is None:
            self._d = np.asarray(coords)
        else:
            d = np.zeros(shape, dtype=np.asarray(data).dtype if len(data) else np.float32)
            d[tuple(np.asarray(c) for c in coords)] = data
            self._d = d

    @property
    def nnz(self):
        return int(np.count_nonzero(self._d))

    @property
    def shape(self):
        return self._d.shape

    @property
    def size(self):
        return self._d.size

    @property
    def ndim(self):
        return self._d.ndim

    def todense(self):
        return self._d

    def sum(self, *a, **k):
        return COO(np.asarray(self._d.sum(*a, **k)))

    def __matmul__(self, o):
        return self._d @ (o._d if isinstance(o, COO) else o)

    def __getitem__(self, i):
        return COO(self._d[i])

    def reshape(self, s):
        return COO(self._d.reshape(s))

    def __mul__(self, o):
        return COO(self._d * (o._d if isinstance(o, COO) else o))

    __rmul__ = __mul__

    def squeeze(self):
        return self._d.squeeze()
