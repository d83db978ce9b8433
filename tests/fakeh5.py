"""A few lines of the h5py group / dataset interface, in memory (h5py is not installed in the build image)."""
import numpy as np


class Dataset(object):
    def __init__(self, shape=None, maxshape=None, dtype=float, data=None):
        self.data = np.array(data) if data is not None else np.zeros(shape, dtype=dtype)
        self.maxshape = maxshape

    shape = property(lambda self: self.data.shape)

    def resize(self, size, axis=0):
        assert self.maxshape is not None and self.maxshape[axis] is None
        new = list(self.data.shape)
        new[axis] = size
        grown = np.zeros(new, dtype=self.data.dtype)
        keep = tuple(slice(0, min(a, b)) for a, b in zip(self.data.shape, new))
        grown[keep] = self.data[keep]
        self.data = grown

    def __setitem__(self, key, value):
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]


class Group(dict):
    def __init__(self):
        dict.__init__(self)
        self.attrs = {}

    def create_group(self, name):
        self[name] = Group()
        return self[name]

    def create_dataset(self, name, shape=None, maxshape=None, dtype=float, data=None):
        self[name] = Dataset(shape, maxshape, dtype, data)
        return self[name]
