"""AP^r evaluation boundary -- mirror of the reference's lib/utils/voc_eval.py for the
segmentation task: `voc_ap` (:19-55) and `voc_eval_sds` (:195-283).

The per-prediction `cv2.resize(mask, box size) >= cfg.BINARIZE_THRESH` (:249-251) runs for all
predictions in one device launch (csrc/render.cu, mnc_binarize_masks); the greedy
true/false-positive assignment stays on the host, as in the reference.  The ground-truth cache
(`parse_inst` :306-348, `check_voc_sds_cache` :351-391) is built from the SBD `inst/` and `cls/`
.mat files with scipy.io, as upstream does, when the `<cls>_mask_gt.pkl` files do not exist yet.
"""
import os
import pickle

import numpy as np
import torch

from mnc_b200 import ops
from mnc_config import cfg


def voc_ap(rec, prec, use_07_metric=False):
    rec, prec = np.asarray(rec, dtype=np.float64), np.asarray(prec, dtype=np.float64)
    if use_07_metric:      # 11-point interpolation
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            sel = rec >= t
            ap += (prec[sel].max() if sel.any() else 0.0) / 11.0
        return ap
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]          # precision envelope
    step = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[step + 1] - mrec[step]) * mpre[step + 1]))


def _region_iou(box_a, mask_a, area_a, box_b, mask_b, area_b):
    """mask_overlap (lib/transform/mask_transform.py:16-46) with the mask sums precomputed."""
    x1, y1 = max(box_a[0], box_b[0]), max(box_a[1], box_b[1])
    x2, y2 = min(box_a[2], box_b[2]), min(box_a[3], box_b[3])
    if x1 > x2 or y1 > y2:
        return 0
    w, h = x2 - x1 + 1, y2 - y1 + 1
    sa = mask_a[y1 - box_a[1]: y1 - box_a[1] + h, x1 - box_a[0]: x1 - box_a[0] + w]
    sb = mask_b[y1 - box_b[1]: y1 - box_b[1] + h, x1 - box_b[0]: x1 - box_b[0] + w]
    assert sa.shape == sb.shape
    inter = np.logical_and(sa, sb).sum()
    union = area_a + area_b - inter
    if union < 1.0:
        return 0
    return float(inter) / float(union)


def eval_sds_arrays(boxes_pkl, masks_pkl, image_names, gt_pkl, ov_thresh=0.5, device=None):
    """Steps 3-7 of `voc_eval_sds` on loaded structures.  boxes_pkl[i]: (n_i,5) [x1,y1,x2,y2,score];
    masks_pkl[i]: (n_i,1,M,M); gt_pkl: {image name: [{'mask_bound': box, 'mask': bool (h,w)}, ...]}
    (not modified).  -> AP (11-point, as the reference calls voc_ap(rec, prec, True))."""
    M = cfg.MASK_SIZE
    dev = torch.device(device or "cuda:%d" % cfg.GPU_ID)
    counts = [len(b) for b in boxes_pkl]
    total = int(sum(counts))
    num_pos = sum(len(v) for v in gt_pkl.values())
    if total == 0:
        return voc_ap(np.zeros(0), np.zeros(0), True)
    boxes = np.concatenate([np.asarray(b, dtype=np.float64).reshape(-1, 5) for b, c in zip(boxes_pkl, counts) if c])
    masks = np.concatenate([np.asarray(m, dtype=np.float32).reshape(-1, M, M) for m, c in zip(masks_pkl, counts) if c])
    owner = np.repeat(np.arange(len(image_names)), counts)
    order = np.argsort(-boxes[:, -1])
    boxes, masks, owner = boxes[order], masks[order], owner[order]
    rboxes = np.round(boxes[:, :4]).astype(np.int64)
    packed, offsets = ops.binarize_masks(torch.from_numpy(rboxes.astype(np.int32)).to(dev),
                                         torch.from_numpy(masks).to(dev),
                                         thresh=float(cfg.BINARIZE_THRESH))
    packed = packed.cpu().numpy().astype(bool)

    gts = {}
    for name, lst in gt_pkl.items():
        gts[name] = [(np.round(g["mask_bound"]).astype(np.int64), np.asarray(g["mask"]),
                      np.asarray(g["mask"]).sum()) for g in lst]
    taken = {name: np.zeros(len(lst), dtype=bool) for name, lst in gt_pkl.items()}
    tp = np.zeros(total)
    fp = np.zeros(total)
    for i in range(total):
        name = image_names[owner[i]]
        if name not in gts:
            fp[i] = 1
            continue
        pb = rboxes[i]
        pm = packed[offsets[i]:offsets[i + 1]].reshape(pb[3] - pb[1] + 1, pb[2] - pb[0] + 1)
        parea = pm.sum()
        best, best_ind = -1000, -1
        for j, (gb, gm, garea) in enumerate(gts[name]):
            ov = _region_iou(gb, gm, garea, pb, pm, parea)
            if ov > best:
                best, best_ind = ov, j
        if best >= ov_thresh and not taken[name][best_ind]:
            tp[i] = 1
            taken[name][best_ind] = True
        else:
            fp[i] = 1
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(num_pos)
    prec = tp / np.maximum(fp + tp, np.finfo(np.float64).eps)
    return voc_ap(rec, prec, True)


def parse_inst(image_name, devkit_path):
    """Instances of one SBD image: [{'mask': bool (h,w) inside the tight bound, 'mask_cls': class id,
    'mask_bound': [x1,y1,x2,y2]}] from `inst/<name>.mat` (GTinst.Segmentation = instance ids) and
    `cls/<name>.mat` (GTcls.Segmentation = class ids)."""
    import scipy.io as sio
    inst = sio.loadmat(os.path.join(devkit_path, "inst", image_name + ".mat"))["GTinst"]["Segmentation"][0][0]
    clsm = sio.loadmat(os.path.join(devkit_path, "cls", image_name + ".mat"))["GTcls"]["Segmentation"][0][0]
    record = []
    for inst_id in np.unique(inst):
        if inst_id == 0:                      # background
            continue
        where = inst == inst_id
        rows, cols = np.where(where)
        x1, y1, x2, y2 = cols.min(), rows.min(), cols.max(), rows.max()
        mask = where[y1:y2 + 1, x1:x2 + 1]
        classes = np.unique(clsm[y1:y2 + 1, x1:x2 + 1][mask])
        assert classes.shape[0] == 1, "instance %d of %s spans several classes" % (inst_id, image_name)
        record.append({"mask": mask, "mask_cls": classes[0],
                       "mask_bound": np.array([x1, y1, x2, y2], dtype=np.float64)})
    return record


def check_voc_sds_cache(cache_dir, devkit_path, image_names, class_names):
    """Write `<cls>_mask_gt.pkl` ({image name: [instance dicts]}) for every class unless all of
    them already exist."""
    os.makedirs(cache_dir, exist_ok=True)
    fg = [(i, n) for i, n in enumerate(class_names) if n != "__background__"]
    if all(os.path.isfile(os.path.join(cache_dir, n + "_mask_gt.pkl")) for _, n in fg):
        return
    per_class = [{} for _ in class_names]
    for image_name in image_names:
        for rec in parse_inst(image_name, devkit_path):
            rec["already_detect"] = False
            per_class[int(rec["mask_cls"])].setdefault(image_name, []).append(rec)
    for i, n in fg:
        with open(os.path.join(cache_dir, n + "_mask_gt.pkl"), "wb") as f:
            pickle.dump(per_class[i], f)


def voc_eval_sds(det_file, seg_file, devkit_path, image_list, cls_name, cache_dir, class_names,
                 ov_thresh=0.5):
    """File-level entry with the reference's signature: `det_file` / `seg_file` are the per-class
    pickles ([image] -> arrays) that `PascalVOCSeg._write_voc_seg_results_file` writes."""
    with open(image_list) as f:
        image_names = [x.strip() for x in f.readlines()]
    check_voc_sds_cache(cache_dir, devkit_path, image_names, class_names)
    gt_cache = os.path.join(cache_dir, cls_name + "_mask_gt.pkl")
    with open(gt_cache, "rb") as f:
        gt_pkl = pickle.load(f, encoding="latin1")
    with open(det_file, "rb") as f:
        boxes_pkl = pickle.load(f, encoding="latin1")
    with open(seg_file, "rb") as f:
        masks_pkl = pickle.load(f, encoding="latin1")
    return eval_sds_arrays(boxes_pkl, masks_pkl, image_names, gt_pkl, ov_thresh)


def reformat_result(all_boxes, all_masks, num_classes, num_images):
    """`PascalVOCSeg._reformat_result` (lib/datasets/pascal_voc_seg.py:179-193): masks to
    (n, M, M) and binarised at cfg.BINARIZE_THRESH before they are written out."""
    M = cfg.MASK_SIZE
    out = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    for c in range(1, num_classes):
        for i in range(num_images):
            m = all_masks[c][i]
            if len(m) == 0:
                continue
            out[c][i] = np.asarray(m).reshape(len(m), M, M) >= cfg.BINARIZE_THRESH
    return all_boxes, out


def write_voc_seg_results_file(all_boxes, all_masks, classes, output_dir):
    """Per-class result pickles `<cls>_det.pkl` / `<cls>_seg.pkl`, the files `voc_eval_sds` reads
    (`PascalVOCSeg._write_voc_seg_results_file`, pascal_voc_seg.py:160-177).  -> list of paths."""
    num_images = len(all_boxes[1]) if len(all_boxes) > 1 else 0
    boxes, masks = reformat_result(all_boxes, all_masks, len(classes), num_images)
    os.makedirs(output_dir, exist_ok=True)
    written = []
    for c, cls in enumerate(classes):
        if cls == "__background__":
            continue
        for suffix, payload in (("_det.pkl", boxes[c]), ("_seg.pkl", masks[c])):
            path = os.path.join(output_dir, cls + suffix)
            with open(path, "wb") as f:
                pickle.dump(payload, f, pickle.HIGHEST_PROTOCOL)
            written.append(path)
    return written
