"""transform.anchors -- reference lib/transform/anchors.py:38-49."""
import numpy as np

from mnc_b200 import ops


def generate_anchors(base_size=16, ratios=[0.5, 1, 2], scales=2 ** np.arange(3, 6)):
    if base_size != 16 or list(ratios) != [0.5, 1, 2] or list(scales) != [8, 16, 32]:
        raise NotImplementedError("only the MNC default anchor set is built into the library")
    return ops.generate_anchors().astype(np.float64)
