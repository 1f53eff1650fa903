"""CPU tests of the oracle itself: the reference's own known answers (anchors table, Caffe's
max-pool vector, conv-vs-naive at Caffe's 1e-4), independent brute-force restatements of the small
algorithms, and the committed golden fixtures (tests/golden/, made by make_golden.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from tests import util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_anchors_known_answer():
    """lib/transform/anchors.py:15-35 lists Shaoqing's MATLAB (1-based) anchors; the Python
    function returns the same windows 0-based, i.e. the table minus one."""
    matlab = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                       [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                       [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]], dtype=np.float64)
    assert np.array_equal(O.generate_anchors(), matlab - 1)


def test_shifted_anchor_order():
    a = O.shifted_anchors(3, 5, 16)
    base = O.generate_anchors()
    assert a.shape == (3 * 5 * 9, 4)
    # index (y*W + x)*9 + a  (proposal_layer.py:96-100)
    y, x, k = 2, 3, 4
    assert np.array_equal(a[(y * 5 + x) * 9 + k], base[k] + np.array([x * 16, y * 16, x * 16, y * 16]))


def test_caffe_maxpool_known_answer_and_ceil_mode_sizes():
    """caffe-mnc/src/caffe/test/test_pooling_layer.cpp:60-118 (2x2 kernel, stride 1)."""
    x = torch.tensor([[1, 2, 5, 2, 3], [9, 4, 1, 4, 8], [1, 2, 5, 2, 3]], dtype=torch.float32)
    x = x.view(1, 1, 3, 5).repeat(2, 2, 1, 1)
    y = F.max_pool2d(x, 2, 1, ceil_mode=True)
    assert y.shape == (2, 2, 2, 4)
    assert torch.equal(y[1, 1], torch.tensor([[9., 5, 5, 8], [9, 5, 5, 8]]))
    # pooling_layer.cpp:90-93 ceil-mode sizes at 600x1000
    h, w = 600, 1000
    for _ in range(4):
        t = F.max_pool2d(torch.zeros(1, 1, h, w), 2, 2, ceil_mode=True)
        h, w = t.shape[2:]
    assert (h, w) == (38, 63)


def test_conv_matches_naive_reference_1e4():
    """test_convolution_layer.cpp:231-263: layer vs naive loop at 1e-4."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1, 3, 6, 7)).astype(np.float32)
    w = rng.normal(size=(4, 3, 3, 3)).astype(np.float32)
    b = rng.normal(size=4).astype(np.float32)
    got = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1).numpy()
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    want = np.zeros((1, 4, 6, 7))
    for o in range(4):
        for i in range(6):
            for j in range(7):
                want[0, o, i, j] = (xp[0, :, i:i + 3, j:j + 3] * w[o]).sum() + b[o]
    assert np.abs(got - want).max() < 1e-4


def test_softmax_sums_to_one():
    """test_softmax_layer.cpp:43-72."""
    w = {"rpn_conv_3x3": (torch.randn(8, 8, 3, 3) * 0.1, torch.zeros(8)),
         "rpn_cls_score": (torch.randn(18, 8, 1, 1), torch.zeros(18)),
         "rpn_bbox_pred": (torch.randn(36, 8, 1, 1), torch.zeros(36))}
    prob, _ = O.rpn_forward(w, torch.randn(1, 8, 5, 6))
    p = prob.numpy()
    assert np.allclose(p[0, :9] + p[0, 9:], 1.0, atol=1e-6)  # channels [bg a | fg a]


def test_nms_against_bruteforce_python():
    boxes = util.random_boxes(200, seed=1, width=300, height=200, smin=8, smax=150, integer=True)
    scores = util.tie_free_scores(200, seed=2)
    dets = np.hstack([boxes, scores[:, None]]).astype(np.float32)
    for thr in (0.3, 0.5, 0.7):
        order = np.argsort(-scores, kind="stable")
        iou = util.iou_matrix64(boxes)
        keep, dead = [], np.zeros(200, bool)
        for i in order:
            if dead[i]:
                continue
            keep.append(i)
            dead |= iou[i] > thr          # strict '>' (nms_kernel.cu:71), not cpu_nms's '>='
        assert [int(k) for k in O.gpu_nms(dets, thr)] == [int(k) for k in keep]


def test_nms_strictness_at_exact_threshold():
    """IoU == thresh must NOT suppress on the GPU path (nms_kernel.cu:71)."""
    dets = np.array([[10, 10, 29, 29, 0.9], [10, 10, 29, 49, 0.8]], dtype=np.float32)  # IoU = 0.5
    assert O.gpu_nms(dets, 0.5) == [0, 1]
    assert O.gpu_nms(dets, 0.49) == [0]
    assert O.nms(np.zeros((0, 5), np.float32), 0.5) == []


def test_roi_warp_semantics():
    feat = np.arange(2 * 1 * 4 * 5, dtype=np.float32).reshape(2, 1, 4, 5)
    # one bilinear sample per cell at start + p*(extent/P), extent = end - start (no +1)
    rois = np.array([[0, 16, 16, 48, 48]], dtype=np.float32)  # s=(1,1), e=(3,3) -> bin 1.0 at P=2
    out = O.roi_warp(feat, rois, 2, 2)
    assert np.array_equal(out[0, 0], feat[0, 0, 1:3, 1:3])
    # degenerate RoI: every cell samples the same point
    out = O.roi_warp(feat, np.array([[1, 32, 16, 32, 16]], dtype=np.float32), 3, 3)
    assert np.all(out == feat[1, 0, 1, 2])
    # sample entirely outside -> 0; CUDA round() is half-away-from-zero: 24/16 = 1.5 -> 2
    out = O.roi_warp(feat, np.array([[0, 24, 24, 24, 24], [0, 900, 900, 950, 950]], np.float32), 1, 1)
    assert out[0, 0, 0, 0] == feat[0, 0, 2, 2] and out[1, 0, 0, 0] == 0


def test_mask_resize_is_1p5_stride_bilinear():
    m = np.arange(21 * 21, dtype=np.float32).reshape(1, 1, 21, 21)
    out = O.mask_resize(m, 14, 14)
    assert out[0, 0, 0, 0] == m[0, 0, 0, 0]
    assert out[0, 0, 2, 4] == m[0, 0, 3, 6]                      # src = dst * 1.5, integer hit
    assert np.isclose(out[0, 0, 1, 1], m[0, 0, 1:3, 1:3].mean())  # (1.5, 1.5) -> 4-tap average


def test_mv_against_naive_render():
    """orc_mv (no render buffer) == the reference's literal algorithm (render all, aggregate,
    reduce, resize) written with numpy loops on a tiny case."""
    g = np.load(os.path.join(GOLD, "voting.npz"))
    boxes, masks = g["boxes"][:12], g["masks"][:12]
    H, W, M = 30, 40, 21
    boxes = boxes * np.array([W / 120.0, H / 90.0, W / 120.0, H / 90.0], dtype=np.float32)
    inds = np.array([0, 3, 5, 1, 2, 7, 11], dtype=np.int32)
    start = np.array([3, 3, 7], dtype=np.int32)   # second result has an empty candidate list
    wts = np.array([0.2, 0.3, 0.5, 0.25, 0.25, 0.25, 0.25], dtype=np.float32)
    rm, rb, agg = O.mv(boxes, masks, inds, start, wts, H, W, return_agg=True)

    def render(n):
        out = np.zeros((H, W), np.float32)
        x1, y1, x2, y2 = boxes[n]
        for h in range(H):
            for w in range(W):
                if w < x1 or w > x2 or h < y1 or h > y2:
                    continue
                rw = np.float32(M) / np.float32(x2 - x1 + np.float32(1))
                rh = np.float32(M) / np.float32(y2 - y1 + np.float32(1))
                ix = (np.float32(w) - x1) * rw
                iy = (np.float32(h) - y1) * rh
                sx, sy = int(np.floor(ix)), int(np.floor(iy))
                mk = masks[n, 0]
                if sx == M - 1 or sy == M - 1:
                    out[h, w] = mk[sy, sx]
                else:
                    fx, fy = ix - sx, iy - sy
                    out[h, w] = ((1 - fx) * (1 - fy) * mk[sy, sx] + fx * (1 - fy) * mk[sy, sx + 1] +
                                 (1 - fx) * fy * mk[sy + 1, sx] + fx * fy * mk[sy + 1, sx + 1])
        return out
    rend = {int(n): render(int(n)) for n in set(inds.tolist())}
    for k in range(3):
        s = 0 if k == 0 else start[k - 1]
        a = np.zeros((H, W), np.float32)
        for i in range(s, start[k]):
            a += rend[int(inds[i])] * wts[i]
        assert np.allclose(a, agg[k], atol=1e-6)
        ys, xs = np.where(a > np.float32(0.4))
        want = [xs.min(), ys.min(), xs.max(), ys.max()] if len(xs) else [W // 2, H // 2, W // 2, H // 2]
        assert not (np.abs(a - 0.4) < 1e-6).any()   # seeded input: no pixel inside float noise of 0.4
        assert list(rb[k]) == [int(v) for v in want]
    assert list(rb[1]) == [W // 2, H // 2, W // 2, H // 2]   # empty list -> defaults (mv_kernel.cu:148,172)


def test_bbox_overlaps_float64():
    b = np.array([[0, 0, 9, 9], [5, 5, 14, 14], [20, 20, 30, 30]], dtype=np.float64)
    ov = O.bbox_overlaps(b, b[:1])
    assert ov[0, 0] == 1.0 and ov[2, 0] == 0.0
    assert np.isclose(ov[1, 0], 25.0 / (100 + 100 - 25))


def test_mask_voting_weight_normalisation_and_end_offsets():
    g = np.load(os.path.join(GOLD, "voting.npz"))
    inds, start, wts, cs, bar = O.mask_voting_candidates(g["boxes"], g["scores"], 21, 100)
    assert np.array_equal(inds, g["inds"]) and np.array_equal(start, g["start"])
    assert np.array_equal(wts, g["weights"]) and np.array_equal(cs, g["cand_scores"])
    assert start[-1] == len(inds) and np.all(np.diff(start) > 0)  # END offsets; self always included
    s = 0
    for e in start:
        assert abs(float(wts[s:e].astype(np.float64).sum()) - 1.0) < 1e-5
        s = e
    assert len(start) >= 100 and bar[-1] == len(start)
    thresh = np.sort(cs)[::-1][99]
    assert cs.min() >= thresh


@pytest.mark.parametrize("name", ["proposal_6x8", "roi_ops", "mask_ops", "nms", "voting", "stage_bridge"])
def test_golden_fixtures(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    if name == "proposal_6x8":
        rois, inter = O.proposal_layer_forward(g["prob"], g["deltas"], g["im_info"], return_intermediate=True)
        assert np.array_equal(rois, g["rois"])
        assert np.array_equal(inter["roi_anchor_index"], g["roi_anchor_index"])
        assert 0 < len(g["keep_filter"]) < 6 * 8 * 9          # the min-size filter bites
        assert len(g["nms_keep"]) < len(g["order"])           # NMS bites
        assert rois[:, 1:].min() >= 0 and rois[:, 3].max() <= 127 and rois[:, 4].max() <= 95
    elif name == "roi_ops":
        for key, (ph, pw) in {"warp28": (28, 28), "warp14": (14, 14), "warp7x5": (7, 5)}.items():
            assert np.array_equal(O.roi_warp(g["feat"], g["rois"], ph, pw), g[key])
        assert np.abs(g["warp14"][4]).max() == 0              # RoI outside the map
    elif name == "mask_ops":
        assert np.array_equal(O.mask_resize(g["masks"], 14, 14), g["resize14"])
        assert np.array_equal(O.mask_resize(g["masks"], 30, 17), g["resize30x17"])
        assert np.array_equal(O.mask_pool(g["feat"], g["resize14"]), g["pooled"])
    elif name == "nms":
        for thr, key in ((0.7, "keep07"), (0.5, "keep05"), (0.3, "keep03")):
            assert np.array_equal(np.asarray(O.gpu_nms(g["dets"], thr)), g[key])
    elif name == "voting":
        rm, rb = O.mv(g["boxes"], g["masks"], g["inds"], g["start"], g["weights"], int(g["hw"][0]), int(g["hw"][1]))
        assert np.array_equal(rb, g["result_box"]) and np.array_equal(rm, g["result_mask"])
    else:
        out = O.stage_bridge_forward(g["rois"], g["deltas"], g["prob"], g["im_info"])
        assert np.array_equal(out, g["rois_ext"])


# ------------------------------------------------------------------ SURVEY.md section 8f rows 3 and 4
def test_roi_pool_against_bruteforce_numpy():
    """ROIPooling restatement (roi_pooling_layer.cu:17-77) vs an independent numpy version:
    round-half-away start/end, +1 sizes forced to >= 1, floor/ceil bin edges, clipping, empty -> 0,
    first maximum wins (argmax)."""
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    feat = rng.normal(size=(2, 6, 20, 30)).astype(np.float32)
    feat[0, 0, 3:6, 4:9] = 7.0     # a plateau: the FIRST maximum must be reported
    rois = np.array([[0, 10, 20, 200, 150], [1, 0, 0, 479, 319], [0, 300, 100, 310, 105],
                     [1, 470, 300, 600, 400], [0, 50, 50, 40, 40], [0, 56, 40, 150, 100],
                     [0, 900, 900, 950, 950]], np.float32)

    def rnd(v):  # CUDA round(): half away from zero
        return int(np.floor(v + 0.5)) if v >= 0 else int(np.ceil(v - 0.5))

    for P in (7, 14, 3):
        out, arg = O.roi_pool(feat, rois, P, P, return_argmax=True)
        for i, roi in enumerate(rois):
            b = int(roi[0])
            sw, sh, ew, eh = (rnd(np.float32(v) * np.float32(0.0625)) for v in roi[1:])
            rw, rh = max(ew - sw + 1, 1), max(eh - sh + 1, 1)
            bh, bw = np.float32(rh) / np.float32(P), np.float32(rw) / np.float32(P)
            for ph in range(P):
                for pw in range(P):
                    h0 = min(max(int(np.floor(np.float32(ph) * bh)) + sh, 0), 20)
                    h1 = min(max(int(np.ceil(np.float32(ph + 1) * bh)) + sh, 0), 20)
                    w0 = min(max(int(np.floor(np.float32(pw) * bw)) + sw, 0), 30)
                    w1 = min(max(int(np.ceil(np.float32(pw + 1) * bw)) + sw, 0), 30)
                    if h1 <= h0 or w1 <= w0:
                        assert np.all(out[i, :, ph, pw] == 0) and np.all(arg[i, :, ph, pw] == -1)
                        continue
                    win = feat[b, :, h0:h1, w0:w1].reshape(6, -1)
                    assert np.array_equal(out[i, :, ph, pw], win.max(axis=1))
                    k = win.argmax(axis=1)                       # numpy argmax = first maximum
                    want_arg = (h0 + k // (w1 - w0)) * 30 + (w0 + k % (w1 - w0))
                    assert np.array_equal(arg[i, :, ph, pw], want_arg)


def test_convert_pred_to_image_known_answers():
    """`_convert_pred_to_image` restatement (vis_seg.py:101-131) on hand-checkable cases: painting
    order, the 150 outline including its numpy `[a-1:a+1]` slices (empty when a == 0), rounding
    half-to-even and clipping of the box."""
    from oracle import oracle as O
    ones = np.ones((21, 21), np.float32)
    pred = {"cls_name": [3, 7], "masks": [ones, ones],
            "boxes": [np.array([10, 10, 29, 24, 0.9], np.float32), np.array([20, 5, 40, 15, 0.8], np.float32)]}
    inst, cls = O.convert_pred_to_image(64, 48, pred)
    assert inst.dtype.kind == "i" and inst.shape == (48, 64)
    assert np.all(inst[5:16, 20:41] == 2)                    # the later instance overwrites
    assert np.all(inst[16:25, 10:30] == 1) and inst[4, 20] == 0
    assert np.all(cls[7:14, 22:39] == 7)                     # interior of instance 2
    assert np.all(cls[5:16, 19:21] == 150) and np.all(cls[4:6, 20:41] == 150)   # its outline
    assert np.all(cls[17:23, 12:28] == 3)                    # instance 1 interior below instance 2
    # box touching the top-left corner: the [-1:1] slices are empty, only x2 / y2 sides are drawn
    pred0 = {"cls_name": [5], "masks": [ones], "boxes": [np.array([0, 0, 9.5, 6.5, 1.0], np.float32)]}
    inst0, cls0 = O.convert_pred_to_image(32, 24, pred0)     # round-half-even: 9.5 -> 10, 6.5 -> 6
    assert np.all(inst0[0:7, 0:11] == 1) and inst0[7, 0] == 0 and inst0[0, 11] == 0
    assert cls0[0, 0] == 5 and cls0[3, 0] == 5               # no outline at x = 0 / y = 0
    assert np.all(cls0[0:7, 9:11] == 150) and np.all(cls0[5:7, 0:11] == 150)
    # a mask that is 0 everywhere paints nothing but its outline
    predz = {"cls_name": [9], "masks": [np.zeros((21, 21), np.float32)],
             "boxes": [np.array([4, 4, 12, 12, 1.0], np.float32)]}
    instz, clsz = O.convert_pred_to_image(20, 20, predz)
    assert instz.sum() == 0 and set(np.unique(clsz)) == {0, 150}
    # binarisation is >= 0.4 on the resized mask
    m = np.full((21, 21), 0.4, np.float32)
    predt = {"cls_name": [2], "masks": [m], "boxes": [np.array([2, 2, 9, 9, 1.0], np.float32)]}
    assert O.convert_pred_to_image(16, 16, predt)[0][5, 5] == 1


def test_voc_ap_and_eval_sds_known_answers():
    from oracle import oracle as O
    # perfect ranking: AP = 1 under both metrics
    rec = np.array([0.25, 0.5, 0.75, 1.0])
    prec = np.ones(4)
    assert O.voc_ap(rec, prec, True) == pytest.approx(1.0) and O.voc_ap(rec, prec, False) == pytest.approx(1.0)
    # never more than half the ground truth found, precision 0.5 throughout
    rec, prec = np.array([0.25, 0.25, 0.5, 0.5]), np.array([1.0, 0.5, 2 / 3, 0.5])
    assert O.voc_ap(rec, prec, True) == pytest.approx((3 * 1.0 + 3 * (2 / 3)) / 11)   # t = 0, .1, .2 | .3, .4, .5
    assert O.voc_ap(rec, prec, False) == pytest.approx(0.25 * 1.0 + 0.25 * (2 / 3))
    # one image, one ground-truth square; a matching prediction, a duplicate and a miss
    gt = {"a": [{"mask_bound": np.array([10, 10, 29, 29]), "mask": np.ones((20, 20), bool)}]}
    ones = np.ones((1, 21, 21), np.float32)
    boxes = [np.array([[10, 10, 29, 29, 0.9], [11, 11, 30, 30, 0.8], [40, 40, 50, 50, 0.7]], np.float32)]
    masks = [np.stack([ones, ones, ones])]
    ap = O.eval_sds(boxes, masks, ["a"], gt, ov_thresh=0.5)
    # tp, fp (duplicate), fp: recall reaches 1 at precision 1 -> 11-point AP = 1
    assert ap == pytest.approx(1.0)
    ap2 = O.eval_sds([boxes[0][::-1].copy()], [masks[0]], ["a"], gt, ov_thresh=0.5)   # same scores order kept
    assert ap2 == pytest.approx(1.0)
    worse = [np.array([[40, 40, 50, 50, 0.95], [10, 10, 29, 29, 0.9]], np.float32)]
    assert O.eval_sds(worse, [np.stack([ones, ones])], ["a"], gt) == pytest.approx(0.5)
    assert O.eval_sds([np.zeros((0, 5), np.float32)], [np.zeros((0, 1, 21, 21), np.float32)], ["a"], gt) == 0.0
    # an image missing from the ground-truth cache only produces false positives
    assert O.eval_sds(boxes, masks, ["zzz"], gt) == 0.0


def test_cfm_blob_helpers():
    """prep_im_for_blob_cfm / pred_rois_for_blob restatements (blob.py:53-106): scale rule with the
    MAX_SIZE cap, zero padding to the largest level, level = scale closest to a 224x224 box."""
    from oracle import oracle as O
    im = O.synthetic_image(0, 100, 400)
    blob, scales = O.prep_im_for_blob_cfm(im, (200, 300), max_size=1000)
    assert np.allclose(scales, [2.0, 2.5])                 # 300/100 = 3 would make 1200 > 1000 -> 1000/400
    assert blob.shape == (2, 3, 250, 1000) and np.all(blob[0, :, 200:, :] == 0) and np.all(blob[0, :, :, 800:] == 0)
    rois = O.pred_rois_for_blob(np.array([[0, 0, 111, 111], [0, 0, 89, 89], [0, 0, 99, 99]], float), scales)
    # areas x scale^2 vs 224^2 = 50176: 112^2 -> 50176 | 78400;  90^2 -> 32400 | 50625;  100^2 -> 40000 | 62500
    assert rois[:, 0].tolist() == [0.0, 1.0, 0.0]
    assert np.allclose(rois[0, 1:], [0, 0, 222, 222]) and np.allclose(rois[1, 1:], [0, 0, 222.5, 222.5])
    one = O.pred_rois_for_blob(np.array([[1, 2, 3, 4]], float), np.array([1.5]))
    assert one.tolist() == [[0.0, 1.5, 3.0, 4.5, 6.0]]
