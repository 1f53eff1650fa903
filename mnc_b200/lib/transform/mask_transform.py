"""transform.mask_transform.gpu_mask_voting -- reference lib/transform/mask_transform.py:213-286.

Same signature and return value; the whole function (20 per-class NMS, global threshold, float64
IoU candidate lists, mask render / aggregate / resize) runs on the device in one pass
(mnc_b200.ops.mask_voting) instead of 20 `_nms` round trips + Python loops + `_mv`."""
import numpy as np
import torch

from mnc_config import cfg
from mnc_b200 import ops


def gpu_mask_voting(masks, boxes, scores, num_classes, max_per_image, im_width, im_height):
    dev = torch.device("cuda", cfg.GPU_ID)
    with torch.cuda.device(dev):
        b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float32)).to(dev)[None]
        m = torch.from_numpy(np.ascontiguousarray(masks, dtype=np.float32)).to(dev)[None]
        s = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32)).to(dev)[None]
        hw = torch.tensor([[int(im_height), int(im_width)]], dtype=torch.int32, device=dev)
        nb = b.shape[1]
        cap = max(128, max_per_image + 28)
        while True:
            r = ops.mask_voting(b, m, s, hw, max_per_image=max_per_image,
                                nms_thresh=cfg.TEST.MASK_MERGE_NMS_THRESH,
                                iou_thresh=cfg.TEST.MASK_MERGE_IOU_THRESH, max_results=cap)
            if int(r["overflow"].item()) == 0:
                break
            cap *= 2  # more score ties at the threshold than slots: retry with room
        k = int(r["n_res"][0].item())
        class_bar = r["class_bar"][0].cpu().numpy()
        result_mask = r["result_mask"][0, :k].cpu().numpy()
        result_box = r["result_box"][0, :k].cpu().numpy()
        cand_scores = r["res_score"][0, :k].cpu().numpy()
    result_box = np.hstack((result_box, cand_scores[:, np.newaxis]))
    list_result_box, list_result_mask = [], []
    for i in range(num_classes - 1):
        cls_start = class_bar[i - 1] if i > 0 else 0
        cls_end = class_bar[i]
        list_result_box.append(result_box[cls_start:cls_end, :])
        list_result_mask.append(result_mask[cls_start:cls_end, :, :, :])
    return list_result_mask, list_result_box


def mask_overlap(box1, box2, mask1, mask2):
    """Region IoU of two binary masks living in different integer boxes
    (reference lib/transform/mask_transform.py:16-46)."""
    from utils.voc_eval import _region_iou
    return _region_iou(box1, np.asarray(mask1), np.asarray(mask1).sum(),
                       box2, np.asarray(mask2), np.asarray(mask2).sum())
