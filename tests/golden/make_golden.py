"""Generates tests/golden/*.npz from the oracle (run in the build container; vectors are committed).

The reference's Python is Python 2 and cannot be imported here, and Caffe cannot be built, so these
fixtures are produced by the oracle restatement itself: they pin the oracle (and, on the GPU box,
the CUDA path) against regressions and carry the edge cases listed in tests/test_oracle_golden.py.
The only reference-held known answers (anchors table, Caffe pooling vector) are asserted directly
in that test, not stored here.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import util  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(2016)
    # ---- ProposalLayer on a 6x8 map (image 96x128): includes boxes clipped at every border,
    # boxes failing the min-size filter, and NMS suppression
    H, W = 6, 8
    logits = rng.normal(0, 1, size=(1, 2, 9 * H, W))
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    prob = (e / e.sum(axis=1, keepdims=True)).reshape(1, 18, H, W).astype(np.float32)
    deltas = rng.normal(0, 0.6, size=(1, 36, H, W)).astype(np.float32)
    deltas[0, 2::4] -= 1.5   # shrink some widths below min_size
    im_info = np.array([[96, 128, 1.0]], dtype=np.float32)
    rois, inter = O.proposal_layer_forward(prob, deltas, im_info, return_intermediate=True)
    np.savez_compressed(os.path.join(OUT, "proposal_6x8.npz"), prob=prob, deltas=deltas,
                        im_info=im_info, rois=rois, all_proposals=inter["all_proposals"],
                        keep_filter=inter["keep_filter"], order=inter["order"],
                        nms_keep=inter["nms_keep"], roi_anchor_index=inter["roi_anchor_index"])
    # ---- ROIWarping / MaskResize / MaskPooling
    feat = np.maximum(rng.normal(size=(2, 6, 9, 13)), 0).astype(np.float32)
    r = np.array([[0, 0, 0, 207, 143], [1, 16, 16, 111, 95], [0, 50, 50, 50, 50],
                  [0, -100, -80, 40, 30], [1, 300, 300, 400, 400], [0, 190, 130, 400, 300],
                  [0, 8, 8, 23.9, 24.1], [1, 100, 90, 60, 40]], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "roi_ops.npz"), feat=feat, rois=r,
                        warp28=O.roi_warp(feat, r, 28, 28), warp14=O.roi_warp(feat, r, 14, 14),
                        warp7x5=O.roi_warp(feat, r, 7, 5))
    masks = rng.uniform(size=(5, 1, 21, 21)).astype(np.float32)
    m14 = O.mask_resize(masks, 14, 14)
    f14 = rng.normal(size=(5, 4, 14, 14)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "mask_ops.npz"), masks=masks, resize14=m14,
                        resize30x17=O.mask_resize(masks, 30, 17), feat=f14, pooled=O.mask_pool(f14, m14))
    # ---- NMS (strict >) with integer boxes incl. exact-threshold pairs
    boxes = util.random_boxes(300, seed=5, width=200, height=150, smin=8, smax=120, integer=True)
    boxes[1] = boxes[0]                       # IoU == 1
    boxes[3] = [10, 10, 29, 29]
    boxes[4] = [10, 10, 29, 49]               # IoU with [3] == 0.5 exactly
    scores = util.tie_free_scores(300, seed=6)
    dets = np.hstack([boxes, scores[:, None]]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), dets=dets,
                        keep07=np.asarray(O.gpu_nms(dets, 0.7), dtype=np.int64),
                        keep05=np.asarray(O.gpu_nms(dets, 0.5), dtype=np.int64),
                        keep03=np.asarray(O.gpu_nms(dets, 0.3), dtype=np.int64))
    # ---- mask voting on a 90x120 image, 80 boxes
    nb, Hh, Ww = 80, 90, 120
    vb = util.random_boxes(nb, seed=7, width=Ww, height=Hh, smin=10, smax=70)
    vm = (1.0 / (1.0 + np.exp(-rng.normal(0, 2, size=(nb, 1, 21, 21))))).astype(np.float32)
    lg = rng.normal(0, 1, size=(nb, 21))
    vs = (np.exp(lg) / np.exp(lg).sum(1, keepdims=True)).astype(np.float32)
    inds, start, wts, cs, bar = O.mask_voting_candidates(vb, vs, 21, 100)
    rm, rb = O.mv(vb, vm, inds, start, wts, Hh, Ww)
    np.savez_compressed(os.path.join(OUT, "voting.npz"), boxes=vb, masks=vm, scores=vs, inds=inds,
                        start=start, weights=wts, cand_scores=cs, class_bar=np.asarray(bar),
                        result_mask=rm, result_box=rb, hw=np.array([Hh, Ww]))
    # ---- StageBridge
    n = 40
    sr = np.hstack([np.zeros((n, 1), np.float32), util.random_boxes(n, 8, width=320, height=224)])
    sd = rng.normal(0, 0.3, size=(n, 84)).astype(np.float32)
    sl = rng.normal(0, 1, size=(n, 21))
    sp = (np.exp(sl) / np.exp(sl).sum(1, keepdims=True)).astype(np.float32)
    si = np.array([[224, 320, 1.0]], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "stage_bridge.npz"), rois=sr.astype(np.float32), deltas=sd,
                        prob=sp, im_info=si, rois_ext=O.stage_bridge_forward(sr.astype(np.float32), sd, sp, si))
    # ---- input preparation: cv2.resize(INTER_LINEAR) of the mean-subtracted float image
    # (blob.py:36-50); up-scaling 1.6x, the cap branch (max side), and down-scaling
    cases = {}
    for name, (h, w, tgt, mx) in {"up": (30, 50, 48, 100), "cap": (20, 70, 48, 100), "down": (90, 120, 48, 100)}.items():
        im = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        out, sc = O.prep_im_for_blob(im, tgt, mx)
        cases["im_" + name] = im
        cases["out_" + name] = out
        cases["scale_" + name] = np.array([sc, tgt, mx], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "prep.npz"), **cases)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
