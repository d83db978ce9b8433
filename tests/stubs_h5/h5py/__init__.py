"""In-memory stand-in for the few h5py calls the trajectory writers make (h5py is not installed in this image).
``File`` keeps everything in memory and, when closed, saves the ``trajectory`` datasets to ``<filename>.npz`` so a
test can look at what a script wrote."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fakeh5 import Group  # noqa: E402


class File(Group):
    def __init__(self, filename, mode="r"):
        Group.__init__(self)
        self.filename = filename

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        out = {}
        for grp in ("system", "trajectory"):
            for key, ds in self.get(grp, {}).items():
                if ds.data.dtype.kind in "fiub":
                    out["%s/%s" % (grp, key)] = ds.data
        np.savez(self.filename + ".npz", **out)
