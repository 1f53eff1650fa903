"""utils.blob -- reference lib/utils/blob.py (input preparation).  `prep_im_for_blob` /
`im_list_to_blob` (:17-50) stay host numpy + cv2 exactly as the reference has them (they feed
`caffe.Net.forward`, whose inputs are host arrays); the batched engine does the same arithmetic on
the device (mnc_prep_images).  `prep_im_for_blob_cfm` (:53-85) builds the CFM image pyramid with
that device kernel; `pred_rois_for_blob` (:88-106) assigns proposals to pyramid levels."""
import numpy as np


def im_list_to_blob(ims):
    max_shape = np.array([im.shape for im in ims]).max(axis=0)
    num_images = len(ims)
    blob = np.zeros((num_images, max_shape[0], max_shape[1], 3), dtype=np.float32)
    for i in range(num_images):
        im = ims[i]
        blob[i, 0:im.shape[0], 0:im.shape[1], :] = im
    return blob.transpose((0, 3, 1, 2))


def prep_im_for_blob(im, pixel_means, target_size, max_size):
    im = im.astype(np.float32, copy=False)
    im -= pixel_means
    im_shape = im.shape
    im_size_min = np.min(im_shape[0:2])
    im_size_max = np.max(im_shape[0:2])
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    if im_scale != 1.0:
        import cv2
        im = cv2.resize(im, None, None, fx=im_scale, fy=im_scale, interpolation=cv2.INTER_LINEAR)
    return im, im_scale


def prep_im_for_blob_cfm(im, input_scales):
    """uint8 BGR image -> (blob (S,3,Hmax,Wmax) fp32 with one mean-subtracted, resized copy per
    scale, zero padded; scale factors).  Resizing runs on the device (cv2 INTER_LINEAR semantics)."""
    import torch
    from mnc_config import cfg
    from mnc_b200 import ops
    im = np.ascontiguousarray(im)
    if im.dtype != np.uint8:
        raise TypeError("prep_im_for_blob_cfm: uint8 BGR image expected (as cv2.imread returns)")
    size_min, size_max = np.min(im.shape[0:2]), np.max(im.shape[0:2])
    dev = torch.device("cuda", cfg.GPU_ID)
    d_im = torch.from_numpy(im).to(dev)[None]
    levels, scales = [], []
    for target_size in input_scales:
        im_scale = float(target_size) / float(size_min)
        if np.round(im_scale * size_max) > cfg.TEST.MAX_SIZE:
            im_scale = float(cfg.TEST.MAX_SIZE) / float(size_max)
        levels.append(ops.prep_images(d_im, im_scale))
        scales.append(im_scale)
    hmax = max(x.shape[2] for x in levels)
    wmax = max(x.shape[3] for x in levels)
    blob = torch.zeros((len(levels), 3, hmax, wmax), dtype=torch.float32, device=dev)
    for i, x in enumerate(levels):
        blob[i, :, :x.shape[2], :x.shape[3]] = x[0]
    return blob.cpu().numpy(), np.array(scales)


def pred_rois_for_blob(im_rois, im_scales):
    """(n,4) boxes -> (n,5) [level, box * scale[level]]; the level is the scale that brings the
    box area closest to 224 x 224."""
    im_rois = np.asarray(im_rois).astype(np.float64, copy=False)
    im_scales = np.asarray(im_scales, dtype=np.float64)
    if len(im_scales) > 1:
        areas = (im_rois[:, 2] - im_rois[:, 0] + 1) * (im_rois[:, 3] - im_rois[:, 1] + 1)
        levels = np.abs(areas[:, None] * im_scales[None, :] ** 2 - 224 * 224).argmin(axis=1)[:, None]
    else:
        levels = np.zeros((im_rois.shape[0], 1), dtype=np.int64)
    return np.hstack((levels.astype(np.float64), im_rois * im_scales[levels]))
