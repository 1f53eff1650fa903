"""Result rendering -- mirror of the reference's lib/utils/vis_seg.py (`_convert_pred_to_image`
:101-131, `_get_voc_color_map` :133-148, `_prepare_dict` :64-98) and of `get_vis_dict`
(tools/demo.py:103-120), on the device kernels of mnc_b200/csrc/render.cu."""
import numpy as np
import torch

from mnc_b200 import ops
from mnc_config import cfg


def _get_voc_color_map(n=256):
    """(n,3) RGB table of the PASCAL VOC palette: bit j (from the top) of each channel is taken
    from bits 3j (R), 3j+1 (G), 3j+2 (B) of the class id."""
    ids = np.arange(n, dtype=np.int64)
    table = np.zeros((n, 3))
    for j in range(8):
        for ch in range(3):
            table[:, ch] += ((ids >> (3 * j + ch)) & 1) << (7 - j)
    return table


def get_vis_dict(result_box, result_mask, img_name, cls_names, vis_thresh=0.5):
    """Per-class voting results -> flat lists of the instances to draw (score >= vis_thresh)."""
    boxes, masks, classes = [], [], []
    for cls_ind in range(len(cls_names)):
        dets = np.asarray(result_box[cls_ind])
        if dets.size == 0:
            continue
        for keep in np.where(dets[:, -1] >= vis_thresh)[0]:
            boxes.append(dets[keep])
            masks.append(result_mask[cls_ind][keep][0])
            classes.append(cls_ind + 1)
    return {"image_name": img_name, "cls_name": classes, "boxes": boxes, "masks": masks}


def _prepare_dict(img_names, cls_names, det_pkl, seg_pkl, vis_thresh=0.5):
    """Same as the reference's `_prepare_dict`, on already-loaded `res_boxes` / `res_masks`
    structures ([class][image] arrays; class 0 = background is skipped)."""
    out = []
    for img_ind, image_name in enumerate(img_names):
        boxes, masks, classes = [], [], []
        for cls_ind, cls_name in enumerate(cls_names):
            if cls_name == "__background__" or len(det_pkl[cls_ind][img_ind]) == 0:
                continue
            dets = det_pkl[cls_ind][img_ind]
            for keep in np.where(dets[:, -1] >= vis_thresh)[0]:
                boxes.append(dets[keep])
                masks.append(seg_pkl[cls_ind][img_ind][keep][0])
                classes.append(cls_ind)
        out.append({"image_name": image_name, "cls_name": classes, "boxes": boxes, "masks": masks})
    return out


def _convert_pred_to_image(img_width, img_height, pred_dict, device=None, want_bgr=False):
    """-> (inst_img, cls_img) integer label images (H, W); with want_bgr also the colour image."""
    n = len(pred_dict["boxes"])
    dev = torch.device(device or "cuda:%d" % cfg.GPU_ID)
    M = cfg.MASK_SIZE
    boxes = torch.zeros((1, max(n, 1), 4), dtype=torch.float32)
    masks = torch.zeros((1, max(n, 1), M, M), dtype=torch.float32)
    cls = torch.zeros((1, max(n, 1)), dtype=torch.int32)
    if n:
        boxes[0] = torch.from_numpy(np.stack([np.asarray(b, dtype=np.float32)[:4] for b in pred_dict["boxes"]]))
        masks[0] = torch.from_numpy(np.stack([np.asarray(m, dtype=np.float32).reshape(M, M)
                                              for m in pred_dict["masks"]]))
        cls[0] = torch.tensor([int(c) for c in pred_dict["cls_name"]], dtype=torch.int32)
    counts = torch.tensor([n], dtype=torch.int32)
    res = ops.paste_instances(boxes.to(dev), masks.to(dev), cls.to(dev), counts.to(dev),
                              int(img_height), int(img_width), thresh=float(cfg.BINARIZE_THRESH),
                              want_bgr=want_bgr)
    out = tuple(r[0].cpu().numpy() for r in res)
    return (out[0].astype(int), out[1].astype(int)) + out[2:]
