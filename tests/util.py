"""Shared generators and comparison helpers for the parity tests (SURVEY.md section 8d inputs)."""
import numpy as np


def random_boxes(n, seed, width=1000, height=600, smin=16, smax=512, integer=False):
    """centres uniform, sizes log-uniform [smin, smax], clipped to the image."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, width, n)
    cy = rng.uniform(0, height, n)
    w = np.exp(rng.uniform(np.log(smin), np.log(smax), n))
    h = np.exp(rng.uniform(np.log(smin), np.log(smax), n))
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, width - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, height - 1)
    if integer:
        b = np.round(b)
    return b.astype(np.float32)


def tie_free_scores(n, seed, lo=0.001, hi=0.999):
    rng = np.random.default_rng(seed)
    return rng.permutation(np.linspace(lo, hi, n)).astype(np.float32)


def iou_matrix64(b):
    """float64 IoU (+1 convention) of every pair, for margin checks."""
    b = b.astype(np.float64)
    area = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    iw = np.minimum(b[:, None, 2], b[None, :, 2]) - np.maximum(b[:, None, 0], b[None, :, 0]) + 1
    ih = np.minimum(b[:, None, 3], b[None, :, 3]) - np.maximum(b[:, None, 1], b[None, :, 1]) + 1
    inter = np.clip(iw, 0, None) * np.clip(ih, 0, None)
    return inter / (area[:, None] + area[None, :] - inter)


def nudge_off_threshold(boxes, thresh, margin=1e-5, max_rounds=20, seed=0):
    """Perturb boxes until no pair's IoU lies within `margin` of `thresh`, so that fp32 rounding /
    FMA-contraction differences between implementations cannot flip a comparison."""
    rng = np.random.default_rng(seed)
    b = boxes.copy()
    for _ in range(max_rounds):
        iou = iou_matrix64(b)
        np.fill_diagonal(iou, 0)
        bad = np.where(np.abs(iou - thresh) < margin)
        if bad[0].size == 0:
            return b
        idx = np.unique(bad[0])
        b[idx, 2] += rng.uniform(0.25, 0.75, idx.size).astype(np.float32)
    raise AssertionError("could not separate IoUs from threshold")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() / denom
