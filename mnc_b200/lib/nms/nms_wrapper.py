"""nms.nms_wrapper -- reference lib/nms/nms_wrapper.py:13-21."""
from mnc_config import cfg
from nms.gpu_nms import gpu_nms


def nms(dets, thresh):
    """Dispatch to the GPU NMS (cfg.USE_GPU_NMS, lib/mnc_config.py:16).  The reference's CPU
    variant uses a different comparison (IoU >= thresh, lib/nms/cpu_nms.pyx:65), is not on the
    path, and is deliberately not provided: there is no CPU fallback."""
    if dets.shape[0] == 0:
        return []
    if cfg.USE_GPU_NMS:
        return gpu_nms(dets, thresh, device_id=cfg.GPU_ID)
    raise NotImplementedError("cpu_nms is out of scope (different semantics, no CPU fallback)")


def apply_nms_mask_single(box, mask, thresh):
    """NMS on one image's (n,5) detections, masks following the kept rows
    (reference lib/nms/nms_wrapper.py:65-71)."""
    if len(box) == 0:
        return box, mask
    keep = nms(box, thresh)
    if len(keep) == 0:
        return box, mask
    return box[keep, :].copy(), mask[keep, :].copy()


def apply_nms(all_boxes, thresh):
    """Per class and image NMS over `all_boxes[cls][image]` (n,5) arrays
    (reference lib/nms/nms_wrapper.py:24-41)."""
    num_classes, num_images = len(all_boxes), len(all_boxes[0])
    out = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    for c in range(num_classes):
        for i in range(num_images):
            dets = all_boxes[c][i]
            if len(dets) == 0:
                continue
            keep = nms(dets, thresh)
            if len(keep) == 0:
                continue
            out[c][i] = dets[keep, :].copy()
    return out
