"""Host-side mirror of the reference's Python API for the inference path.

Put this directory on sys.path (``mnc_b200.lib.install()``) and the reference's import lines work
unchanged:  ``import caffe``, ``from mnc_config import cfg``, ``from nms.nms_wrapper import nms``,
``from nms.mv import mv``, ``from utils.cython_bbox import bbox_overlaps``,
``from transform.mask_transform import gpu_mask_voting``, ``from pylayer.proposal_layer import
ProposalLayer`` ...  (module names from lib/setup.py:114-160 and the reference's lib/ tree).
Every function keeps the reference's signature, argument meaning, dtypes and return shapes; the
work happens in libmnc_b200.so on the GPU.
"""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
