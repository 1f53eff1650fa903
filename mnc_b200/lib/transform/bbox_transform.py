"""transform.bbox_transform -- the three functions on the inference path
(reference lib/transform/bbox_transform.py:64-130).  The callers use them on the host on a few
hundred boxes (tools/demo.py:94-95); the device versions live in csrc/proposal.cu."""
import numpy as np


def bbox_transform_inv(boxes, deltas):
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    boxes = boxes.astype(deltas.dtype, copy=False)
    widths = boxes[:, 2] - boxes[:, 0] + 1.0
    heights = boxes[:, 3] - boxes[:, 1] + 1.0
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    pred_ctr_x = deltas[:, 0::4] * widths[:, np.newaxis] + ctr_x[:, np.newaxis]
    pred_ctr_y = deltas[:, 1::4] * heights[:, np.newaxis] + ctr_y[:, np.newaxis]
    pred_w = np.exp(deltas[:, 2::4]) * widths[:, np.newaxis]
    pred_h = np.exp(deltas[:, 3::4]) * heights[:, np.newaxis]
    pred_boxes = np.zeros(deltas.shape, dtype=deltas.dtype)
    pred_boxes[:, 0::4] = pred_ctr_x - 0.5 * pred_w
    pred_boxes[:, 1::4] = pred_ctr_y - 0.5 * pred_h
    pred_boxes[:, 2::4] = pred_ctr_x + 0.5 * pred_w
    pred_boxes[:, 3::4] = pred_ctr_y + 0.5 * pred_h
    return pred_boxes


def clip_boxes(boxes, im_shape):
    x1, y1, x2, y2 = boxes[:, 0::4], boxes[:, 1::4], boxes[:, 2::4], boxes[:, 3::4]
    keep = np.where((x1 >= 0) & (x2 <= im_shape[1] - 1) & (y1 >= 0) & (y2 <= im_shape[0] - 1))[0]
    clipped = np.zeros(boxes.shape, dtype=boxes.dtype)
    clipped[:, 0::4] = np.maximum(np.minimum(x1, im_shape[1] - 1), 0)
    clipped[:, 1::4] = np.maximum(np.minimum(y1, im_shape[0] - 1), 0)
    clipped[:, 2::4] = np.maximum(np.minimum(x2, im_shape[1] - 1), 0)
    clipped[:, 3::4] = np.maximum(np.minimum(y2, im_shape[0] - 1), 0)
    return clipped, keep


def filter_small_boxes(boxes, min_size):
    ws = boxes[:, 2] - boxes[:, 0] + 1
    hs = boxes[:, 3] - boxes[:, 1] + 1
    return np.where((ws >= min_size) & (hs >= min_size))[0]
