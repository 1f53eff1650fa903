"""cfg constants consumed by the inference path -- reference lib/mnc_config.py (values at :16-28,
:112-152).  Only the keys the path reads are present; the YAML merge machinery is out of scope."""
import numpy as np


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


cfg = _AttrDict()
cfg.USE_GPU_NMS = True                       # :16
cfg.GPU_ID = 0                               # :17
cfg.PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])  # :20
cfg.BINARIZE_THRESH = 0.4                    # :26
cfg.MASK_SIZE = 21                           # :28
cfg.TRAIN = _AttrDict(MAX_SIZE=1000, SCALES=(600,))
cfg.TEST = _AttrDict()
cfg.TEST.SCALES = (600,)                     # :115
cfg.TEST.MAX_SIZE = 1000                     # :118
cfg.TEST.NMS = 0.3                           # :122
cfg.TEST.RPN_NMS_THRESH = 0.7                # :126
cfg.TEST.RPN_PRE_NMS_TOP_N = 6000            # :128
cfg.TEST.RPN_POST_NMS_TOP_N = 300            # :130
cfg.TEST.RPN_MIN_SIZE = 16                   # :132
cfg.TEST.MASK_MERGE_IOU_THRESH = 0.5         # :136
cfg.TEST.MASK_MERGE_NMS_THRESH = 0.3         # :137
cfg.TEST.USE_MASK_MERGE = True               # :151
cfg.TEST.USE_GPU_MASK_MERGE = True           # :152
cfg.TEST.CFM_INPUT_MASK_SIZE = 14            # :138
cfg.TEST.MAX_ROIS_GPU = [2000]               # :144
cfg.TEST.GROUP_SCALE = 1                     # :145
cfg.TEST.USE_TOP_K_MCG = 0                   # :148
