import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    g["meta"] = json.loads(str(g["meta"]))
    shape = tuple(g["meta"]["shape"])
    n = int(np.prod(shape))
    g["binaries"] = np.unpackbits(g["binaries_packed"])[:n].reshape(shape)
    g["csv"] = bytes(g["csv"]).decode()
    return g
