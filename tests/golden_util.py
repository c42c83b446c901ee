"""Loader for tests/golden/dense_small.npz (written by tests/golden/make_golden.py from the
unmodified reference)."""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_small():
    z = np.load(os.path.join(HERE, "golden", "dense_small.npz"))
    nb = int(z["nblocks"])
    blocks = []
    for i in range(nb):
        blocks.append(dict(gid=i, particles=z[f"b{i}_particles"], tets=z[f"b{i}_tets"], num_orig=int(z[f"b{i}_num_orig"]),
                           bounds_min=z[f"b{i}_bounds"][0], bounds_max=z[f"b{i}_bounds"][1], vert_to_tet=z[f"b{i}_v2t"]))
    return z, blocks, tuple(int(x) for x in z["gsize"])


def load_c1():
    return np.load(os.path.join(HERE, "golden", "config1.npz"))
