"""Load the committed fixtures of tests/golden/ (made by tests/golden/make_golden.py)."""
import os

import numpy as np

from . import cases as K

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(c: K.Case):
    z = np.load(os.path.join(GOLDEN_DIR, c.name + ".npz"))
    params = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    return params, (z["u"], z["gout"]), ref
