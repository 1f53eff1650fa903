"""nms.gpu_nms -- reference lib/nms/gpu_nms.pyx:16-31 over mnc_nms_host (the `_nms` drop-in)."""
import ctypes

import numpy as np

from mnc_b200._lib import lib, check


def _order_desc(scores):
    # `scores.argsort()[::-1]` (gpu_nms.pyx:26) with ties resolved (score desc, index asc)
    return np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))


def gpu_nms(dets, thresh, device_id=0):
    if dets.dtype != np.float32 or dets.ndim != 2:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' 2-D")  # cython's check
    boxes_num, boxes_dim = dets.shape
    keep = np.zeros(boxes_num, dtype=np.int32)
    num_out = ctypes.c_int(0)
    scores = dets[:, 4]
    order = _order_desc(scores)
    sorted_dets = np.ascontiguousarray(dets[order, :])
    check(lib.mnc_nms_host(keep.ctypes.data_as(ctypes.c_void_p), ctypes.byref(num_out),
                           sorted_dets.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(boxes_num),
                           ctypes.c_int(boxes_dim), ctypes.c_float(thresh),
                           ctypes.c_int(device_id)), "mnc_nms_host")
    keep = keep[:num_out.value]
    return list(order[keep])
