"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's (daijifeng001/MNC) 5-stage inference path, used only as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under mnc_b200/ imports this module.

 * the reference's Python layers (host numpy in the reference too) are transliterated to
   Python 3 / numpy, expression by expression, each function citing file:line;
 * the reference's CUDA kernels are restated in plain C (oracle/mnc_oracle.c, loaded via ctypes);
 * conv / inner-product / pooling / softmax / sigmoid follow Caffe's CPU semantics through
   torch CPU fp32 (im2col + sgemm class; caffe-mnc/src/caffe/util/im2col.cpp:19-55,
   util/math_functions.cpp:19, layers/base_conv_layer.cpp:257-279, inner_product_layer.cpp:31-34).

Pinning status: pinned to the reference's own code throughout.
 * The Python restatements (anchors, bbox transforms, ProposalLayer, StageBridgeLayer, MaskLayer,
   gpu_mask_voting host logic, prep_im_for_blob, im_detect tail, bbox_overlaps, the cfg constants)
   are held bit for bit to fixtures the REFERENCE's own Python produced, run in the build container
   from /root/reference (scripts/make_ref_fixtures.py -> tests/golden/ref_*.npz;
   tests/test_ref_fixtures.py).  Two lines evaluate differently under numpy 1.x (what the reference
   ran on) and numpy 2 (this image); both evaluations are recorded and reproduced (`numpy2=`).
 * The C restatements of the CUDA kernels are held to the reference's sources compiled unmodified
   into oracle/_ref (tests/test_ref_pin.py): bit-exact against the -fmad=false build.
 * Known answers the reference itself holds (anchors table lib/transform/anchors.py:15-35, Caffe
   max-pooling test_pooling_layer.cpp:60-118, conv-vs-naive) are in tests/test_oracle_golden.py.
 * conv / inner-product arithmetic lives in cuDNN / cuBLAS / a CBLAS the reference does not pin:
   tolerance parity (1e-3 relative, north_star) against torch CPU fp32 is the only meaningful bar.
 * The SURVEY.md section 8f additions (rendering, the AP^r evaluator, the CFM blob helpers) are
   cross-checked by brute-force and hand-computed known answers in tests/test_oracle_golden.py;
   ROIPooling is pinned like the other layers.

Tie rule: the reference sorts with `scores.argsort()[::-1]` (proposal_layer.py:139,
gpu_nms.pyx:26), numpy's unstable introsort, so the order of equal scores is unspecified there.
This oracle (and the CUDA path) use (score descending, index ascending).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c():
    """Compile oracle/mnc_oracle.c -> oracle/liboracle.so (gcc, no FMA contraction)."""
    import subprocess
    src = os.path.join(_HERE, "mnc_oracle.c")
    out = os.path.join(_HERE, "liboracle.so")
    if (not os.path.exists(out)) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared",
                               "-fPIC", "-o", out, src, "-lm"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- config constants
class CFG:
    """lib/mnc_config.py values consumed by the path (line numbers in SURVEY.md Appendix A)."""
    PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])  # :20
    BINARIZE_THRESH = 0.4                                       # :26
    MASK_SIZE = 21                                              # :28
    TEST_SCALES = (600,)                                        # :115
    TEST_MAX_SIZE = 1000                                        # :118
    TEST_NMS = 0.3                                              # :122
    RPN_NMS_THRESH = 0.7                                        # :126
    RPN_PRE_NMS_TOP_N = 6000                                    # :128
    RPN_POST_NMS_TOP_N = 300                                    # :130
    RPN_MIN_SIZE = 16                                           # :132
    MASK_MERGE_IOU_THRESH = 0.5                                 # :136
    MASK_MERGE_NMS_THRESH = 0.3                                 # :137
    FEAT_STRIDE = 16                                            # test.prototxt:473


# --------------------------------------------------------------------------- anchors
def _whctrs(anchor):  # lib/transform/anchors.py:52-61
    w = anchor[2] - anchor[0] + 1
    h = anchor[3] - anchor[1] + 1
    x_ctr = anchor[0] + 0.5 * (w - 1)
    y_ctr = anchor[1] + 0.5 * (h - 1)
    return w, h, x_ctr, y_ctr


def _mkanchors(ws, hs, x_ctr, y_ctr):  # anchors.py:64-76
    ws = ws[:, np.newaxis]
    hs = hs[:, np.newaxis]
    return np.hstack((x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1),
                      x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)))


def _ratio_enum(anchor, ratios):  # anchors.py:79-90
    w, h, x_ctr, y_ctr = _whctrs(anchor)
    size = w * h
    size_ratios = size / ratios
    ws = np.round(np.sqrt(size_ratios))
    hs = np.round(ws * ratios)
    return _mkanchors(ws, hs, x_ctr, y_ctr)


def _scale_enum(anchor, scales):  # anchors.py:93-102
    w, h, x_ctr, y_ctr = _whctrs(anchor)
    ws = w * scales
    hs = h * scales
    return _mkanchors(ws, hs, x_ctr, y_ctr)


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    """lib/transform/anchors.py:38-49."""
    ratios = np.array(ratios, dtype=np.float64)
    base_anchor = np.array([1, 1, base_size, base_size]) - 1
    ratio_anchors = _ratio_enum(base_anchor, ratios)
    return np.vstack([_scale_enum(ratio_anchors[i, :], scales)
                      for i in range(ratio_anchors.shape[0])])


# --------------------------------------------------------------------------- bbox transforms
def bbox_transform_inv(boxes, deltas):
    """lib/transform/bbox_transform.py:64-99 (all arithmetic in deltas.dtype = fp32, un-fused)."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    boxes = boxes.astype(deltas.dtype, copy=False)
    widths = boxes[:, 2] - boxes[:, 0] + 1.0
    heights = boxes[:, 3] - boxes[:, 1] + 1.0
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    dx = deltas[:, 0::4]
    dy = deltas[:, 1::4]
    dw = deltas[:, 2::4]
    dh = deltas[:, 3::4]
    pred_ctr_x = dx * widths[:, np.newaxis] + ctr_x[:, np.newaxis]
    pred_ctr_y = dy * heights[:, np.newaxis] + ctr_y[:, np.newaxis]
    pred_w = np.exp(dw) * widths[:, np.newaxis]
    pred_h = np.exp(dh) * heights[:, np.newaxis]
    pred_boxes = np.zeros(deltas.shape, dtype=deltas.dtype)
    pred_boxes[:, 0::4] = pred_ctr_x - 0.5 * pred_w
    pred_boxes[:, 1::4] = pred_ctr_y - 0.5 * pred_h
    pred_boxes[:, 2::4] = pred_ctr_x + 0.5 * pred_w
    pred_boxes[:, 3::4] = pred_ctr_y + 0.5 * pred_h
    return pred_boxes


def clip_boxes(boxes, im_shape):
    """lib/transform/bbox_transform.py:102-120.  im_shape = (height, width[, ...])."""
    x1 = boxes[:, 0::4]
    y1 = boxes[:, 1::4]
    x2 = boxes[:, 2::4]
    y2 = boxes[:, 3::4]
    keep = np.where((x1 >= 0) & (x2 <= im_shape[1] - 1) & (y1 >= 0) & (y2 <= im_shape[0] - 1))[0]
    clipped = np.zeros(boxes.shape, dtype=boxes.dtype)
    clipped[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], im_shape[1] - 1), 0)
    clipped[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], im_shape[0] - 1), 0)
    clipped[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], im_shape[1] - 1), 0)
    clipped[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], im_shape[0] - 1), 0)
    return clipped, keep


def filter_small_boxes(boxes, min_size):
    """lib/transform/bbox_transform.py:123-130."""
    ws = boxes[:, 2] - boxes[:, 0] + 1
    hs = boxes[:, 3] - boxes[:, 1] + 1
    return np.where((ws >= min_size) & (hs >= min_size))[0]


def order_desc(scores):
    """`scores.argsort()[::-1]` with the documented tie rule (score desc, index asc)."""
    scores = np.asarray(scores).ravel()
    return np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))


# --------------------------------------------------------------------------- NMS
def nms_sorted(boxes_sorted, thresh):
    """`_nms` (lib/nms/nms_kernel.cu:91-144) on already score-sorted boxes -> kept positions."""
    b = np.ascontiguousarray(boxes_sorted, dtype=np.float32)
    n, dim = b.shape
    keep = np.zeros(max(n, 1), dtype=np.int32)
    num = ctypes.c_int(0)
    _lib().orc_nms(_p(b), ctypes.c_int(n), ctypes.c_int(dim), ctypes.c_float(thresh), _p(keep),
                   ctypes.byref(num))
    return keep[:num.value].copy()


def gpu_nms(dets, thresh):
    """lib/nms/gpu_nms.pyx:16-31: sort by score desc, `_nms`, map back to original indices."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    order = order_desc(dets[:, 4])
    keep = nms_sorted(dets[order, :], thresh)
    return list(order[keep])


def nms(dets, thresh):
    """lib/nms/nms_wrapper.py:13-21 with cfg.USE_GPU_NMS = True."""
    if dets.shape[0] == 0:
        return []
    return gpu_nms(dets, thresh)


def bbox_overlaps(boxes, query_boxes):
    """lib/utils/bbox.pyx:15-55 (float64)."""
    b = np.ascontiguousarray(boxes, dtype=np.float64)
    q = np.ascontiguousarray(query_boxes, dtype=np.float64)
    out = np.zeros((b.shape[0], q.shape[0]), dtype=np.float64)
    _lib().orc_bbox_overlaps(_p(b), ctypes.c_int(b.shape[0]), _p(q), ctypes.c_int(q.shape[0]),
                             _p(out))
    return out


# --------------------------------------------------------------------------- Python layers
def shifted_anchors(height, width, feat_stride=16):
    """proposal_layer.py:84-100: (K*A, 4) float64, index (y*W + x)*A + a."""
    anchors0 = generate_anchors()
    A = anchors0.shape[0]
    shift_x = np.arange(0, width) * feat_stride
    shift_y = np.arange(0, height) * feat_stride
    shift_x, shift_y = np.meshgrid(shift_x, shift_y)
    shifts = np.vstack((shift_x.ravel(), shift_y.ravel(), shift_x.ravel(), shift_y.ravel())).transpose()
    K = shifts.shape[0]
    anchors = anchors0.reshape((1, A, 4)) + shifts.reshape((1, K, 4)).transpose((1, 0, 2))
    return anchors.reshape((K * A, 4))


def proposal_layer_forward(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, return_intermediate=False):
    """ProposalLayer.forward, TEST phase -- lib/pylayer/proposal_layer.py:52-175.
    rpn_cls_prob_reshape (1, 2A, H, W) fp32, rpn_bbox_pred (1, 4A, H, W) fp32, im_info (1, 3)."""
    assert rpn_cls_prob_reshape.shape[0] == 1
    A = 9
    pre_nms_topN = CFG.RPN_PRE_NMS_TOP_N
    post_nms_topN = CFG.RPN_POST_NMS_TOP_N
    nms_thresh = CFG.RPN_NMS_THRESH
    min_size = CFG.RPN_MIN_SIZE
    scores = rpn_cls_prob_reshape[:, A:, :, :]
    bbox_deltas = rpn_bbox_pred
    im_info = np.asarray(im_info, dtype=np.float32).reshape(-1, 3)[0, :]
    height, width = scores.shape[-2:]
    anchors = shifted_anchors(height, width, CFG.FEAT_STRIDE)
    bbox_deltas = bbox_deltas.transpose((0, 2, 3, 1)).reshape((-1, 4))
    scores = scores.transpose((0, 2, 3, 1)).reshape((-1, 1))
    proposals = bbox_transform_inv(anchors, bbox_deltas)
    proposals, _ = clip_boxes(proposals, im_info[:2])
    all_proposals, all_scores = proposals, scores
    keep_filter = filter_small_boxes(proposals, min_size * im_info[2])
    proposals = proposals[keep_filter, :]
    scores = scores[keep_filter]
    order = order_desc(scores)
    if pre_nms_topN > 0:
        order = order[:pre_nms_topN]
    proposals = proposals[order, :]
    scores = scores[order]
    keep = nms(np.hstack((proposals, scores)), nms_thresh)
    if post_nms_topN > 0:
        keep = keep[:post_nms_topN]
    keep = np.asarray(keep, dtype=np.int64)
    final = proposals[keep, :]
    batch_inds = np.zeros((final.shape[0], 1), dtype=np.float32)
    rois = np.hstack((batch_inds, final.astype(np.float32, copy=False)))
    if return_intermediate:
        return rois, {
            "all_proposals": all_proposals.astype(np.float32), "all_scores": all_scores.ravel(),
            "keep_filter": keep_filter, "order": order, "sorted_proposals": proposals,
            "sorted_scores": scores.ravel(), "nms_keep": keep,
            # index into the (h, w, a) anchor enumeration of every output RoI
            "roi_anchor_index": keep_filter[order[keep]],
        }
    return rois


def mask_layer_forward(mask_output):
    """MaskLayer.forward_test -- lib/pylayer/mask_layer.py:95-102."""
    n = mask_output.shape[0]
    return mask_output.reshape((n, 1, CFG.MASK_SIZE, CFG.MASK_SIZE))


def stage_bridge_forward(rois, bbox_pred, seg_cls_prob, im_info):
    """StageBridgeLayer.forward_test -- lib/pylayer/stage_bridge_layer.py:237-255."""
    im_info = np.asarray(im_info, dtype=np.float32).reshape(-1, 3)
    all_rois = bbox_transform_inv(rois[:, 1:5], bbox_pred)
    score_max = seg_cls_prob.argmax(axis=1)
    rois_out = np.zeros((rois.shape[0], 5))
    rois_out[:, 0] = 0
    for i in range(len(score_max)):
        rois_out[i, 1:5] = all_rois[i, 4 * score_max[i]:4 * (score_max[i] + 1)]
    rois_out[:, 1:5], _ = clip_boxes(rois_out[:, 1:5], im_info[0, :2])
    return rois_out.astype(np.float32)


# --------------------------------------------------------------------------- Caffe MNC layers (C)
def roi_warp(feat, rois, pooled_h, pooled_w, spatial_scale=0.0625):
    """ROIWarpingLayer forward -- roi_warping_layer.cu:67-107.  feat (B,C,H,W), rois (R,5)."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.zeros((R, C, pooled_h, pooled_w), dtype=np.float32)
    _lib().orc_roi_warp(_p(feat), ctypes.c_int(C), ctypes.c_int(H), ctypes.c_int(W), _p(rois),
                        ctypes.c_int(R), ctypes.c_int(pooled_h), ctypes.c_int(pooled_w),
                        ctypes.c_float(spatial_scale), _p(out))
    return out


def roi_pool(feat, rois, pooled_h, pooled_w, spatial_scale=0.0625, return_argmax=False):
    """ROIPoolingLayer forward -- roi_pooling_layer.cu:17-77.  feat (B,C,H,W), rois (R,5)."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    B, C, H, W = feat.shape
    R = rois.shape[0]
    out = np.zeros((R, C, pooled_h, pooled_w), dtype=np.float32)
    arg = np.zeros((R, C, pooled_h, pooled_w), dtype=np.int32)
    _lib().orc_roi_pool(_p(feat), ctypes.c_int(C), ctypes.c_int(H), ctypes.c_int(W), _p(rois),
                        ctypes.c_int(R), ctypes.c_int(pooled_h), ctypes.c_int(pooled_w),
                        ctypes.c_float(spatial_scale), _p(out), _p(arg))
    return (out, arg) if return_argmax else out


def mask_resize(masks, out_h, out_w):
    """MaskResizeLayer forward -- mask_resize_layer.cu:57-73."""
    masks = np.ascontiguousarray(masks, dtype=np.float32)
    N, C, ih, iw = masks.shape
    out = np.zeros((N, C, out_h, out_w), dtype=np.float32)
    _lib().orc_mask_resize(_p(masks), ctypes.c_int(N), ctypes.c_int(C), ctypes.c_int(ih),
                           ctypes.c_int(iw), ctypes.c_int(out_h), ctypes.c_int(out_w), _p(out))
    return out


def mask_pool(feat, mask):
    """MaskPoolingLayer forward -- mask_pooling_layer.cu:13-26."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    N, C, H, W = feat.shape
    assert mask.shape == (N, 1, H, W)
    out = np.zeros_like(feat)
    _lib().orc_mask_pool(_p(feat), _p(mask), ctypes.c_int(N), ctypes.c_int(C), ctypes.c_int(H),
                         ctypes.c_int(W), _p(out))
    return out


def mv(all_boxes, all_masks, candidate_inds, candidate_start, candidate_weights, image_height,
       image_width, return_agg=False):
    """nms.mv.mv -- lib/nms/gpu_mv.pyx:13-31 -> `_mv` (lib/nms/mv_kernel.cu:242-348)."""
    all_boxes = np.ascontiguousarray(all_boxes, dtype=np.float32)
    all_masks = np.ascontiguousarray(all_masks, dtype=np.float32)
    candidate_inds = np.ascontiguousarray(candidate_inds, dtype=np.int32)
    candidate_start = np.ascontiguousarray(candidate_start, dtype=np.int32)
    candidate_weights = np.ascontiguousarray(candidate_weights, dtype=np.float32)
    nb, box_dim = all_boxes.shape
    M = all_masks.shape[3]
    k = candidate_start.shape[0]
    result_mask = np.zeros((k, 1, all_masks.shape[2], M), dtype=np.float32)
    result_box = np.zeros((k, box_dim), dtype=np.int32)
    agg = np.zeros((k, image_height, image_width), dtype=np.float32) if return_agg else None
    _lib().orc_mv(_p(all_boxes), _p(all_masks), ctypes.c_int(nb), _p(candidate_inds),
                  _p(candidate_start), _p(candidate_weights), ctypes.c_int(candidate_inds.shape[0]),
                  ctypes.c_int(image_height), ctypes.c_int(image_width), ctypes.c_int(box_dim),
                  ctypes.c_int(M), ctypes.c_int(k), _p(result_mask), _p(result_box),
                  _p(agg) if return_agg else ctypes.c_void_p(0))
    if return_agg:
        return result_mask, result_box, agg
    return result_mask, result_box


def mask_voting_candidates(boxes, scores, num_classes, max_per_image, numpy2=False):
    """Host half of gpu_mask_voting -- lib/transform/mask_transform.py:213-274.
    Returns (candidate_inds i32, candidate_start i32 END offsets, candidate_weights f32,
    candidate_scores f32, class_bar list).

    Weight normalisation (`cur_weights / sum(cur_weights)`, :266-267): the reference ran under
    numpy 1.x where builtin sum() over fp32 scalars starting from int 0 accumulates in float64,
    and `f32_array / f64_scalar` is evaluated in fp32 with the scalar cast to fp32; restated so
    (numpy2=False, the semantics the CUDA path implements).  numpy2=True is the same line as
    numpy >= 2 evaluates it (NEP 50: the int 0 is weak, the sum stays float32).  Both are pinned
    to the reference's own function run under each rule (tests/golden/ref_voting.npz)."""
    sup_boxes, sup_scores, tobesort = [], [], []
    for i in range(num_classes):
        if i == 0:
            sup_boxes.append([])
            sup_scores.append([])
            continue
        dets = np.hstack((boxes.astype(np.float32), scores[:, i:i + 1])).astype(np.float32)
        inds = nms(dets, CFG.MASK_MERGE_NMS_THRESH)
        inds = np.asarray(inds, dtype=np.int64)
        ind_boxes = boxes[inds]
        ind_scores = scores[inds, i]
        num_keep = min(len(ind_scores), max_per_image)
        sup_boxes.append(ind_boxes[0:num_keep, :])
        sup_scores.append(ind_scores[0:num_keep])
        tobesort.extend(ind_scores[0:num_keep])
    sorted_scores = np.sort(np.asarray(tobesort, dtype=np.float32))[::-1]
    num_keep = min(len(sorted_scores), max_per_image)
    thresh = sorted_scores[num_keep - 1]
    candidate_inds, candidate_weights, candidate_start, candidate_scores, class_bar = [], [], [], [], []
    for c in range(num_classes):
        if c == 0:
            continue
        cls_box = sup_boxes[c]
        cls_score = sup_scores[c]
        keep = np.where(cls_score >= thresh)[0]
        new_sup_boxes = cls_box[keep]
        for i in range(len(new_sup_boxes)):
            cur_ov = bbox_overlaps(boxes.astype(np.float64),
                                   new_sup_boxes[i, np.newaxis].astype(np.float64))
            cur_inds = np.where(cur_ov >= CFG.MASK_MERGE_IOU_THRESH)[0]
            candidate_inds.extend(cur_inds)
            cur_weights = scores[cur_inds, c].astype(np.float32)
            if numpy2:
                total = np.float32(0.0)
                for v in cur_weights:  # builtin sum(): sequential, float32 accumulator
                    total = np.float32(total + v)
            else:
                total = np.float64(0.0)
                for v in cur_weights:  # builtin sum(): sequential, float64 accumulator
                    total = total + np.float64(v)
            cur_weights = cur_weights / np.float32(total)
            candidate_weights.extend(cur_weights)
            candidate_start.append(len(candidate_inds))
        candidate_scores.extend(cls_score[keep])
        class_bar.append(len(candidate_scores))
    return (np.array(candidate_inds, dtype=np.int32), np.array(candidate_start, dtype=np.int32),
            np.array(candidate_weights, dtype=np.float32),
            np.array(candidate_scores, dtype=np.float32), class_bar)


def gpu_mask_voting(masks, boxes, scores, num_classes, max_per_image, im_width, im_height,
                    numpy2=False):
    """lib/transform/mask_transform.py:213-286."""
    inds, start, weights, cscores, class_bar = mask_voting_candidates(
        boxes, scores, num_classes, max_per_image, numpy2=numpy2)
    result_mask, result_box = mv(boxes.astype(np.float32), masks, inds, start, weights,
                                 im_height, im_width)
    result_box = np.hstack((result_box, cscores[:, np.newaxis]))
    list_result_box, list_result_mask = [], []
    for i in range(num_classes - 1):
        cls_start = class_bar[i - 1] if i > 0 else 0
        cls_end = class_bar[i]
        list_result_box.append(result_box[cls_start:cls_end, :])
        list_result_mask.append(result_mask[cls_start:cls_end, :, :, :])
    return list_result_mask, list_result_box


# --------------------------------------------------------------------------- network
TRUNK_NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
               "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"]
POOL_AFTER = {"conv1_2", "conv2_2", "conv3_3", "conv4_3"}
# Weights are data handed to the oracle by its caller: {caffe layer name: (weight, bias)} fp32
# torch CPU tensors in Caffe layouts (conv (Cout,Cin,kh,kw); InnerProduct (N,K), K = (c,h,w)).


def synthetic_image(i=0, height=600, width=1000):
    """uint8 BGR uniform 0..255, rng seed 1234+i."""
    rng = np.random.default_rng(1234 + i)
    return rng.integers(0, 256, size=(height, width, 3), dtype=np.uint8)


def prep_blob(im):
    """prep_im_for_blob + im_list_to_blob at scale 1.0 (lib/utils/blob.py:17-50): mean-subtract,
    HWC -> (1,3,H,W) fp32.  (Synthetic 600x1000 inputs need no resize: 600/600 = 1.0 and
    round(1.0*1000) <= 1000, blob.py:43-46.)"""
    x = im.astype(np.float32, copy=True)
    x -= CFG.PIXEL_MEANS
    blob = x[np.newaxis].transpose(0, 3, 1, 2).astype(np.float32)
    im_info = np.array([[blob.shape[2], blob.shape[3], 1.0]], dtype=np.float32)
    return np.ascontiguousarray(blob), im_info


def prep_im_for_blob(im, target_size=600, max_size=1000):
    """lib/utils/blob.py:36-50 verbatim (cv2 is the reference's own dependency): returns the
    mean-subtracted, resized float32 HWC image and the scale."""
    import cv2
    im = im.astype(np.float32, copy=True)
    im -= CFG.PIXEL_MEANS
    im_shape = im.shape
    im_size_min = np.min(im_shape[0:2])
    im_size_max = np.max(im_shape[0:2])
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    im = cv2.resize(im, None, None, fx=im_scale, fy=im_scale, interpolation=cv2.INTER_LINEAR)
    return im, im_scale


def _t(x):
    import torch
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def trunk_forward(w, data):
    """conv1_1 .. conv5_3 (+ReLU, 4x ceil-mode 2x2 max pool) -- test.prototxt:19-387."""
    import torch
    import torch.nn.functional as F
    x = _t(data)
    for name in TRUNK_NAMES:
        x = F.relu(F.conv2d(x, w[name][0], w[name][1], padding=1))
        if name in POOL_AFTER:
            x = F.max_pool2d(x, 2, 2, ceil_mode=True)  # pooling_layer.cpp:90-93
    return x


def rpn_forward(w, conv5_3):
    """rpn_conv_3x3+ReLU, rpn_cls_score, rpn_bbox_pred, Reshape/Softmax/Reshape --
    test.prototxt:391-462; softmax with max-subtraction (softmax_layer.cu:86-120)."""
    import torch
    import torch.nn.functional as F
    x = F.relu(F.conv2d(conv5_3, w["rpn_conv_3x3"][0], w["rpn_conv_3x3"][1], padding=1))
    cls = F.conv2d(x, w["rpn_cls_score"][0], w["rpn_cls_score"][1])
    bbox = F.conv2d(x, w["rpn_bbox_pred"][0], w["rpn_bbox_pred"][1])
    n, c, h, wd = cls.shape
    prob = torch.softmax(cls.reshape(n, 2, -1, wd), dim=1).reshape(n, c, h, wd)
    return prob, bbox


def head_forward(w, roi_feat14):
    """One cascade stage after RoI warping (+28->14 pool in stage 1): test.prototxt:509-785.
    roi_feat14: (R, C, 14, 14) fp32 (`roi_interpolate_conv5`).  Returns dict of blobs."""
    import torch
    import torch.nn.functional as F
    x14 = _t(roi_feat14)
    R = x14.shape[0]
    out = {}
    h = F.relu(F.linear(x14.reshape(R, -1), *w["fc6_maskest"]))
    mask_pred = F.linear(h, *w["mask_pred"])
    mask_output = torch.sigmoid(mask_pred)
    out["mask_pred"] = mask_pred.numpy()
    mask_proposal = mask_layer_forward(mask_output.numpy())
    out["mask_proposal"] = mask_proposal
    mask14 = mask_resize(mask_proposal, 14, 14)
    out["mask_proposal_resize"] = mask14
    box7 = F.max_pool2d(x14, 2, 2)
    fc6 = F.relu(F.linear(box7.reshape(R, -1), *w["fc6"]))
    fc7 = F.relu(F.linear(fc6, *w["fc7"]))
    masked = _t(mask_pool(x14.numpy(), mask14))
    m7 = F.max_pool2d(masked, 2, 2)
    out["roi_interpolate_conv5_mask"] = m7.numpy()
    fc6m = F.relu(F.linear(m7.reshape(R, -1), *w["fc6_mask"]))
    fc7m = F.relu(F.linear(fc6m, *w["fc7_mask"]))
    join = torch.cat([fc7m, fc7], dim=1)  # Concat order: mask first (test.prototxt:700-705)
    out["join_box_mask"] = join.numpy()
    out["cls_prob"] = torch.softmax(F.linear(join, *w["cls_score"]), dim=1).numpy()
    out["seg_cls_score"] = F.linear(join, *w["seg_cls_score"]).numpy()
    out["seg_cls_prob"] = torch.softmax(_t(out["seg_cls_score"]), dim=1).numpy()
    out["bbox_pred"] = F.linear(join, *w["bbox_pred"]).numpy()
    return out


def net_forward(w, data, im_info):
    """Whole 5-stage forward for ONE image (the reference is batch-1, proposal_layer.py:65).
    Returns the blob dict the callers read (tools/demo.py:84-90) plus intermediates."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        blobs = {}
        conv5_3 = trunk_forward(w, data)
        blobs["conv5_3"] = conv5_3.numpy()
        prob, bbox = rpn_forward(w, conv5_3)
        blobs["rpn_cls_prob_reshape"] = prob.numpy()
        blobs["rpn_bbox_pred"] = bbox.numpy()
        rois = proposal_layer_forward(blobs["rpn_cls_prob_reshape"], blobs["rpn_bbox_pred"], im_info)
        blobs["rois"] = rois
        premax = roi_warp(blobs["conv5_3"], rois, 28, 28)
        feat14 = F.max_pool2d(_t(premax), 2, 2).numpy()
        blobs["roi_interpolate_conv5"] = feat14
        s1 = head_forward(w, feat14)
        for k, v in s1.items():
            blobs[k] = v
        rois_ext = stage_bridge_forward(rois, s1["bbox_pred"], s1["seg_cls_prob"], im_info)
        blobs["rois_ext"] = rois_ext
        feat14e = roi_warp(blobs["conv5_3"], rois_ext, 14, 14)
        blobs["roi_interpolate_conv5_ext"] = feat14e
        s2 = head_forward(w, feat14e)
        for k, v in s2.items():
            blobs[k + "_ext"] = v
    return blobs


def frcnn_net_forward(w, data, im_info):
    """Faster R-CNN test net for one image -- models/VGG16/faster_rcnn_end2end/test.prototxt:
    trunk, RPN, proposal (:463-476), ROIWarping 7x7 (:479-490), fc6/fc7 (Dropout is the identity
    in TEST phase, dropout_layer.cpp:41-44), cls_score+Softmax, bbox_pred (:558-616)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        blobs = {}
        conv5_3 = trunk_forward(w, data)
        blobs["conv5_3"] = conv5_3.numpy()
        prob, bbox = rpn_forward(w, conv5_3)
        blobs["rpn_cls_prob_reshape"] = prob.numpy()
        blobs["rpn_bbox_pred"] = bbox.numpy()
        rois = proposal_layer_forward(blobs["rpn_cls_prob_reshape"], blobs["rpn_bbox_pred"], im_info)
        blobs["rois"] = rois
        pool5 = roi_warp(blobs["conv5_3"], rois, 7, 7)
        blobs["pool5"] = pool5
        R = rois.shape[0]
        fc6 = F.relu(F.linear(_t(pool5).reshape(R, -1), *w["fc6"]))
        fc7 = F.relu(F.linear(fc6, *w["fc7"]))
        blobs["fc7"] = fc7.numpy()
        blobs["cls_prob"] = torch.softmax(F.linear(fc7, *w["cls_score"]), dim=1).numpy()
        blobs["bbox_pred"] = F.linear(fc7, *w["bbox_pred"]).numpy()
    return blobs


def detection_tail(blobs, im_shape, im_scale=1.0):
    """TesterWrapper._detection_forward :226-237."""
    boxes = blobs["rois"][:, 1:5] / im_scale
    pred_boxes = bbox_transform_inv(boxes, blobs["bbox_pred"])
    pred_boxes, _ = clip_boxes(pred_boxes, im_shape)
    return blobs["cls_prob"], pred_boxes


def cfm_net_forward(w, data, rois, masks):
    """CFM test net -- models/VGG16/cfm/test.prototxt: data (S,3,H,W) image pyramid, rois (R,5)
    [level, x1,y1,x2,y2], masks (R,1,14,14).  ROIPooling 7x7 -> fc6/fc7 (:397-443); ROIPooling
    14x14 -> MaskPooling with the input masks -> 2x2 max pool -> fc6_mask/fc7_mask (:447-512);
    fc6_maskest/mask_pred/Sigmoid on the un-masked 14x14 feature (:517-549); Concat, cls_score,
    seg_cls_score, bbox_pred (:553-620)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        blobs = {}
        conv5_3 = trunk_forward(w, data).numpy()
        blobs["conv5_3"] = conv5_3
        R = rois.shape[0]
        box7 = roi_pool(conv5_3, rois, 7, 7)
        blobs["roi_pooling_conv5"] = box7
        fc6 = F.relu(F.linear(_t(box7).reshape(R, -1), *w["fc6"]))
        fc7 = F.relu(F.linear(fc6, *w["fc7"]))
        feat14 = roi_pool(conv5_3, rois, 14, 14)
        blobs["roi_pooling_conv5_mask"] = feat14
        masked = mask_pool(feat14, masks)
        m7 = F.max_pool2d(_t(masked), 2, 2)
        blobs["roi_mask_conv5_pool"] = m7.numpy()
        fc6m = F.relu(F.linear(m7.reshape(R, -1), *w["fc6_mask"]))
        fc7m = F.relu(F.linear(fc6m, *w["fc7_mask"]))
        h = F.relu(F.linear(_t(feat14).reshape(R, -1), *w["fc6_maskest"]))
        blobs["mask_pred"] = F.linear(h, *w["mask_pred"]).numpy()
        blobs["mask_prob"] = torch.sigmoid(_t(blobs["mask_pred"])).numpy()
        join = torch.cat([fc7m, fc7], dim=1)
        blobs["cls_prob"] = torch.softmax(F.linear(join, *w["cls_score"]), dim=1).numpy()
        blobs["seg_cls_score"] = F.linear(join, *w["seg_cls_score"]).numpy()
        blobs["seg_cls_prob"] = torch.softmax(_t(blobs["seg_cls_score"]), dim=1).numpy()
        blobs["bbox_pred"] = F.linear(join, *w["bbox_pred"]).numpy()
    return blobs


def prep_im_for_blob_cfm(im, input_scales, max_size=1000):
    """lib/utils/blob.py:53-85: one resized copy per scale, zero-padded into one blob."""
    import cv2
    im_orig = im.astype(np.float32, copy=True)
    im_orig -= CFG.PIXEL_MEANS
    size_min, size_max = np.min(im_orig.shape[0:2]), np.max(im_orig.shape[0:2])
    ims, scales = [], []
    for target_size in input_scales:
        im_scale = float(target_size) / float(size_min)
        if np.round(im_scale * size_max) > max_size:
            im_scale = float(max_size) / float(size_max)
        ims.append(cv2.resize(im_orig, None, None, fx=im_scale, fy=im_scale, interpolation=cv2.INTER_LINEAR))
        scales.append(im_scale)
    max_shape = np.array([i.shape for i in ims]).max(axis=0)
    blob = np.zeros((len(ims), max_shape[0], max_shape[1], 3), dtype=np.float32)
    for i, x in enumerate(ims):
        blob[i, 0:x.shape[0], 0:x.shape[1], :] = x
    return blob.transpose((0, 3, 1, 2)), np.array(scales)


def pred_rois_for_blob(im_rois, im_scales):
    """lib/utils/blob.py:88-106: level = scale whose scaled area is closest to 224^2."""
    im_rois = im_rois.astype(np.float64, copy=False)
    if len(im_scales) > 1:
        widths = im_rois[:, 2] - im_rois[:, 0] + 1
        heights = im_rois[:, 3] - im_rois[:, 1] + 1
        areas = widths * heights
        scaled_areas = areas[:, np.newaxis] * (im_scales[np.newaxis, :] ** 2)
        levels = np.abs(scaled_areas - 224 * 224).argmin(axis=1)[:, np.newaxis]
    else:
        levels = np.zeros((im_rois.shape[0], 1), dtype=np.int64)
    im_rois = im_rois * im_scales[levels]
    return np.hstack((levels.astype(np.float64), im_rois))


def im_detect_tail(blobs, im_shape, im_scale=1.0, numpy2=False):
    """tools/demo.py:84-100 (== TesterWrapper.py:244-260).

    `rois[:, 1:5] / im_scales[0]` divides a float32 array by a 0-d float64 array: under the numpy
    1.x the reference ran on that is a float32 division by float32(scale) (value-based casting)
    and boxes stay float32 -- numpy2=False, what the CUDA path implements; under numpy >= 2 it is a
    float64 division and boxes come out float64 -- numpy2=True, pinned bit-exact to the reference's
    own im_detect run here (tests/golden/ref_prep_tail.npz).  The two agree to 1 float32 ulp."""
    if numpy2:
        div = np.float64(im_scale)
        rois1 = blobs["rois"][:, 1:5].astype(np.float64) / div
        rois2 = blobs["rois_ext"][:, 1:5].astype(np.float64) / div
    else:
        div = np.float32(im_scale)
        rois1 = blobs["rois"][:, 1:5] / div
        rois2 = blobs["rois_ext"][:, 1:5] / div
    rois1, _ = clip_boxes(rois1, im_shape)
    rois2, _ = clip_boxes(rois2, im_shape)
    masks = np.concatenate((blobs["mask_proposal"], blobs["mask_proposal_ext"]), axis=0)
    boxes = np.concatenate((rois1, rois2), axis=0)
    if not numpy2:
        boxes = boxes.astype(np.float32)
    scores = np.concatenate((blobs["seg_cls_prob"], blobs["seg_cls_prob_ext"]), axis=0)
    return boxes, masks, scores


def im_detect(w, im):
    """prepare_mnc_args + net.forward + tail for a uint8/fp32 BGR image at scale 1.0."""
    data, im_info = prep_blob(im)
    blobs = net_forward(w, data, im_info)
    return im_detect_tail(blobs, im.shape), blobs


# ------------------------------------------------------------------- result rendering / AP^r
def voc_color_map(n=256):  # vis_seg.py:133-148
    cmap = np.zeros((n, 3))
    for i in range(n):
        r = g = b = 0
        cid = i
        for j in range(8):
            bits = np.unpackbits(np.array([cid], dtype=np.uint8))
            r |= int(bits[-1]) << (7 - j)
            g |= int(bits[-2]) << (7 - j)
            b |= int(bits[-3]) << (7 - j)
            cid >>= 3
        cmap[i] = (r, g, b)
    return cmap


def convert_pred_to_image(img_width, img_height, pred_dict, thresh=0.4):  # vis_seg.py:101-131
    import cv2
    inst_img = np.zeros((img_height, img_width))
    cls_img = np.zeros((img_height, img_width))
    for i in range(len(pred_dict["boxes"])):
        box = np.round(pred_dict["boxes"][i]).astype(int)
        box[0] = min(max(box[0], 0), img_width - 1)
        box[1] = min(max(box[1], 0), img_height - 1)
        box[2] = min(max(box[2], 0), img_width - 1)
        box[3] = min(max(box[3], 0), img_height - 1)
        mask = cv2.resize(pred_dict["masks"][i].astype(np.float32),
                          (int(box[2] - box[0] + 1), int(box[3] - box[1] + 1)))
        mask = mask >= thresh
        ys, xs = slice(box[1], box[3] + 1), slice(box[0], box[2] + 1)
        inst_img[ys, xs] = np.where(mask, i + 1, inst_img[ys, xs])
        cls_img[ys, xs] = np.where(mask, pred_dict["cls_name"][i], cls_img[ys, xs])
        cls_img[box[1]:box[3] + 1, box[0] - 1:box[0] + 1] = 150
        cls_img[box[1]:box[3] + 1, box[2] - 1:box[2] + 1] = 150
        cls_img[box[1] - 1:box[1] + 1, box[0]:box[2] + 1] = 150
        cls_img[box[3] - 1:box[3] + 1, box[0]:box[2] + 1] = 150
    return inst_img.astype(int), cls_img.astype(int)


def voc_ap(rec, prec, use_07_metric=False):  # voc_eval.py:19-55
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap += p / 11.0
        return ap
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def mask_overlap(box1, box2, mask1, mask2):  # mask_transform.py:16-46
    x1, y1 = max(box1[0], box2[0]), max(box1[1], box2[1])
    x2, y2 = min(box1[2], box2[2]), min(box1[3], box2[3])
    if x1 > x2 or y1 > y2:
        return 0
    w, h = x2 - x1 + 1, y2 - y1 + 1
    a = mask1[y1 - box1[1]: y1 - box1[1] + h, x1 - box1[0]: x1 - box1[0] + w]
    b = mask2[y1 - box2[1]: y1 - box2[1] + h, x1 - box2[0]: x1 - box2[0] + w]
    inter = np.logical_and(b, a).sum()
    union = mask1.sum() + mask2.sum() - inter
    if union < 1.0:
        return 0
    return float(inter) / float(union)


def eval_sds(boxes_pkl, masks_pkl, image_names, gt_pkl, ov_thresh=0.5, thresh=0.4, mask_size=21):
    """voc_eval.py:195-283 steps 3-7 on in-memory structures: boxes_pkl[i] (n_i,5),
    masks_pkl[i] (n_i,1,M,M) per image; gt_pkl {image: [{'mask_bound', 'mask'}]}."""
    import cv2
    import copy
    gt_pkl = copy.deepcopy(gt_pkl)
    for lst in gt_pkl.values():
        for g in lst:
            g["already_detect"] = 0
    box_num = sum(len(b) for b in boxes_pkl)
    new_boxes = np.zeros((box_num, 5))
    new_masks = np.zeros((box_num, mask_size, mask_size))
    new_image = []
    cnt = 0
    for image_ind in range(len(image_names)):
        for box_ind in range(len(boxes_pkl[image_ind])):
            new_boxes[cnt] = boxes_pkl[image_ind][box_ind]
            new_masks[cnt] = masks_pkl[image_ind][box_ind]
            new_image.append(image_names[image_ind])
            cnt += 1
    keep_inds = np.argsort(-new_boxes[:, -1])
    new_boxes = new_boxes[keep_inds, :]
    new_masks = new_masks[keep_inds, :, :]
    num_pred = new_boxes.shape[0]
    fp = np.zeros((num_pred, 1))
    tp = np.zeros((num_pred, 1))
    for i in range(num_pred):
        pred_box = np.round(new_boxes[i, :4]).astype(int)
        pred_mask = cv2.resize(new_masks[i].astype(np.float32),
                               (int(pred_box[2] - pred_box[0] + 1), int(pred_box[3] - pred_box[1] + 1)))
        pred_mask = pred_mask >= thresh
        image_index = new_image[keep_inds[i]]
        if image_index not in gt_pkl:
            fp[i] = 1
            continue
        cur_overlap, cur_ind = -1000, -1
        for ind2, gt in enumerate(gt_pkl[image_index]):
            ov = mask_overlap(np.round(gt["mask_bound"]).astype(int), pred_box, gt["mask"], pred_mask)
            if ov > cur_overlap:
                cur_overlap, cur_ind = ov, ind2
        if cur_overlap >= ov_thresh:
            if gt_pkl[image_index][cur_ind]["already_detect"]:
                fp[i] = 1
            else:
                tp[i] = 1
                gt_pkl[image_index][cur_ind]["already_detect"] = 1
        else:
            fp[i] = 1
    num_pos = sum(len(v) for v in gt_pkl.values())
    fp = np.cumsum(fp)
    tp = np.cumsum(tp)
    rec = tp / float(num_pos)
    prec = tp / np.maximum(fp + tp, np.finfo(np.float64).eps)
    return voc_ap(rec, prec, True)
