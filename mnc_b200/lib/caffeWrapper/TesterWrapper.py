"""Dataset-level result collection -- mirror of the reference's lib/caffeWrapper/TesterWrapper.py
for the segmentation task (`get_result` :46-67, `get_segmentation_result` :149-214,
`_segmentation_forward` :239-260): run every image of an imdb through the 5-stage net + mask
voting, keep [class][image] arrays, write `res_boxes.pkl` / `res_masks.pkl`, hand them to
`imdb.evaluate_segmentation`.

The reference walks the images one by one (batch 1).  Here images are bucketed by size and sent
through the batched engine (`Detector.im_detect_images` + device mask voting), `max_batch` at a
time; results are stored by image index, so the output structure is the reference's.

`imdb` is duck-typed as in the reference: `image_index`, `num_classes`, `image_path_at(i)`,
`evaluate_segmentation(all_boxes, all_masks, output_dir)`; optional `image_at(i)` returning a
BGR uint8 array bypasses cv2.imread, `output_dir` overrides the default output location.
"""
import heapq
import os
import pickle

import numpy as np
import torch

from mnc_b200 import ops
from mnc_b200.api import Detector
from mnc_config import cfg
from nms.nms_wrapper import apply_nms_mask_single


class TesterWrapper(object):
    def __init__(self, test_prototxt, imdb, test_model, task_name, max_batch=8, device=None):
        from caffe.net import load_weights
        self.device = torch.device(device or "cuda:%d" % cfg.GPU_ID)
        self.detector = Detector(load_weights(test_prototxt, test_model), device=self.device,
                                 max_batch=max_batch)
        self.name = (os.path.splitext(os.path.basename(test_model))[0]
                     if isinstance(test_model, str) else "mnc_5stage")
        self.imdb = imdb
        self.output_dir = getattr(imdb, "output_dir", None) or os.path.join("output", self.name)
        self.task_name = task_name
        self.num_images = len(self.imdb.image_index)
        self.num_classes = self.imdb.num_classes
        self.max_per_set = 40 * self.num_images     # :39
        self.max_per_image = 100                    # :41
        os.makedirs(self.output_dir, exist_ok=True)

    # ------------------------------------------------------------------ :46-67
    def get_result(self):
        det_file = os.path.join(self.output_dir, "res_boxes.pkl")
        seg_file = os.path.join(self.output_dir, "res_masks.pkl")
        if self.task_name != "seg":
            raise NotImplementedError("task '%s': only the MNC 5-stage 'seg' task is on this path"
                                      % self.task_name)
        if os.path.isfile(det_file) and os.path.isfile(seg_file):
            with open(det_file, "rb") as f:
                seg_box = pickle.load(f)
            with open(seg_file, "rb") as f:
                seg_mask = pickle.load(f)
        else:
            seg_box, seg_mask = self.get_segmentation_result()
            with open(det_file, "wb") as f:
                pickle.dump(seg_box, f, pickle.HIGHEST_PROTOCOL)
            with open(seg_file, "wb") as f:
                pickle.dump(seg_mask, f, pickle.HIGHEST_PROTOCOL)
        return self.imdb.evaluate_segmentation(seg_box, seg_mask, self.output_dir)

    def _read(self, i):
        if hasattr(self.imdb, "image_at"):
            return np.ascontiguousarray(self.imdb.image_at(i))
        import cv2
        return cv2.imread(self.imdb.image_path_at(i))

    def _buckets(self):
        """image indices grouped by image size, each group cut into engine batches."""
        by_shape = {}
        for i in range(self.num_images):
            by_shape.setdefault(self._read(i).shape, []).append(i)
        mb = self.detector.max_batch
        for idx in by_shape.values():
            for s in range(0, len(idx), mb):
                yield idx[s:s + mb]

    # ------------------------------------------------------------------ :149-214
    def get_segmentation_result(self):
        nc, ni = self.num_classes, self.num_images
        thresh = -np.inf * np.ones(nc)
        top_scores = [[] for _ in range(nc)]
        all_boxes = [[[] for _ in range(ni)] for _ in range(nc)]
        all_masks = [[[] for _ in range(ni)] for _ in range(nc)]
        det = self.detector
        for batch in self._buckets():
            ims = np.stack([self._read(i) for i in batch])
            H, W = ims.shape[1:3]
            if cfg.TEST.USE_MASK_MERGE:
                if not cfg.TEST.USE_GPU_MASK_MERGE:
                    raise NotImplementedError("cpu_mask_voting is off the GPU path (out of scope)")
                per_image = self._vote_batch(ims)
                for i, (result_mask, result_box) in zip(batch, per_image):
                    for j in range(1, nc):
                        all_boxes[j][i] = result_box[j - 1]
                        all_masks[j][i] = result_mask[j - 1]
            else:
                boxes, masks, scores, valid, _ = det.im_detect_images(ims)
                for b, i in enumerate(batch):
                    ok = valid[b].astype(bool)
                    bx, mk, sc = boxes[b][ok], masks[b][ok], scores[b][ok]
                    for j in range(1, nc):
                        inds = np.where(sc[:, j] > thresh[j])[0]
                        top = np.argsort(-sc[inds, j])[:self.max_per_image]
                        inds = inds[top]
                        for val in sc[inds, j]:
                            heapq.heappush(top_scores[j], val)
                        if len(top_scores[j]) > self.max_per_set:
                            while len(top_scores[j]) > self.max_per_set:
                                heapq.heappop(top_scores[j])
                            thresh[j] = top_scores[j][0]
                        dets = np.hstack((bx[inds], sc[inds, j][:, None])).astype(np.float32, copy=False)
                        all_boxes[j][i], all_masks[j][i] = apply_nms_mask_single(
                            dets, mk[inds].astype(np.float32, copy=False), cfg.TEST.NMS)
        for j in range(1, nc):
            for i in range(ni):
                if len(all_boxes[j][i]) == 0:
                    continue
                inds = np.where(all_boxes[j][i][:, -1] > thresh[j])[0]
                all_boxes[j][i] = all_boxes[j][i][inds, :]
                all_masks[j][i] = all_masks[j][i][inds]
        return all_boxes, all_masks

    def _vote_batch(self, ims):
        """forward + gpu_mask_voting for one batch of equally sized images, everything resident
        on the device until the voted results come back.  -> per image (list_mask, list_box) in
        the format `gpu_mask_voting` returns (mask_transform.py:270-286)."""
        det = self.detector
        B, H, W = ims.shape[:3]
        dev = self.device
        scale = ops.im_scale_for((H, W))
        out_h, out_w = int(np.rint(H * scale)), int(np.rint(W * scale))
        det._fit_input(out_h, out_w)
        with torch.cuda.device(dev):
            d_u8 = torch.from_numpy(ims).to(dev)
            ops.prep_images(d_u8, scale, out=det._d_in[:B])
            info = torch.tensor([[out_h, out_w, scale]] * B, dtype=torch.float32, device=dev)
            hw = torch.tensor([[H, W]] * B, dtype=torch.float32, device=dev)
            sc = torch.full((B,), scale, dtype=torch.float32, device=dev)
            boxes, masks, scores, valid, _ = det.engine.detect(det._d_in[:B], info, hw, sc)
            vote = det.mask_voting(boxes, masks, scores, valid, [[H, W]] * B,
                                   max_per_image=self.max_per_image)
            return unpack_voting(vote, self.num_classes)


def unpack_voting(vote, num_classes):
    """Device voting results -> per image (list_result_mask, list_result_box), each a list over
    the num_classes-1 foreground classes of (k,1,M,M) fp32 masks / (k,5) fp32 [box, score]."""
    n_res = vote["n_res"].cpu().numpy()
    rcls = vote["res_class"].cpu().numpy()
    rscore = vote["res_score"].cpu().numpy()
    rmask = vote["result_mask"].cpu().numpy()
    rbox = vote["result_box"].cpu().numpy()
    out = []
    M = rmask.shape[-1]
    for b in range(len(n_res)):
        k = int(n_res[b])
        list_mask, list_box = [], []
        for c in range(1, num_classes):
            sel = np.where(rcls[b, :k] == c)[0]
            list_mask.append(rmask[b, sel].reshape(-1, 1, M, M).astype(np.float32))
            list_box.append(np.hstack((rbox[b, sel].astype(np.float32),
                                       rscore[b, sel, None].astype(np.float32))))
        out.append((list_mask, list_box))
    return out
