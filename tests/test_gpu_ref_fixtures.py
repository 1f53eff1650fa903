"""CUDA path == REFERENCE, on the fixtures the reference's own Python produced
(tests/golden/ref_*.npz, scripts/make_ref_fixtures.py; the oracle is held to the same files on the
CPU in tests/test_ref_fixtures.py).  Integer / index results bit-exact; coordinates that pass
through expf on the device and numpy's exp in the reference within 1e-5 relative (fixtures are
generated with margins so no decision sits inside that noise -- these tests have no escape hatch)."""
import numpy as np
import pytest
import torch

from tests import util
from tests.test_ref_fixtures import load, voting_case, tail_case

pytestmark = pytest.mark.gpu


def _install():
    import mnc_b200.lib as L
    L.install()


def _blobs(*arrays):
    import caffe
    out = []
    for a in arrays:
        b = caffe.Blob()
        if a is not None:
            b.data = a
        out.append(b)
    return out


def test_proposal_layer_vs_reference_forward():
    """ProposalLayer drop-in driven through the caffe.Layer protocol on the inputs the reference's
    ProposalLayer.forward (proposal_layer.py:52-175) was run on: same number of RoIs, same RoIs in
    the same order (a different anchor would be tens of pixels away), coordinates 1e-5."""
    _install()
    import caffe
    from pylayer.proposal_layer import ProposalLayer
    f = load("ref_proposal.npz")
    for tag in f["cases"]:
        layer = ProposalLayer(param_str="'feat_stride': 16", phase=caffe.TEST)
        bottom = _blobs(f["prob_" + tag], f["deltas_" + tag], f["im_info_" + tag])
        top = _blobs(None)
        layer.setup(bottom, top)
        layer.reshape(bottom, top)
        layer.forward(bottom, top)
        got, want = top[0].data, f["rois_" + tag]
        assert got.dtype == np.float32 and got.shape == want.shape, (tag, got.shape, want.shape)
        assert np.all(got[:, 0] == 0)
        assert np.allclose(got, want, rtol=1e-5, atol=2e-3), (tag, np.abs(got - want).max())


def test_proposal_indices_vs_reference_forward():
    """Device pipeline internals against the index lists the reference layer keeps
    (_ind_after_filter, _ind_after_sort, _proposal_index): bit-exact."""
    from mnc_b200 import ops
    f = load("ref_proposal.npz")
    for tag in f["cases"]:
        prob, deltas, im_info = f["prob_" + tag], f["deltas_" + tag], f["im_info_" + tag]
        H, W = prob.shape[2:]
        r, cnt, dev = ops.proposals_from_rpn(torch.from_numpy(prob).cuda(), torch.from_numpy(deltas).cuda(),
                                             torch.from_numpy(im_info).cuda(), 1, H, W, "nchw", False,
                                             batch_index_mode=False, return_intermediate=True)
        valid = dev["valid"][0].cpu().numpy()
        keep_filter = np.where(valid != 0)[0]
        assert np.array_equal(keep_filter, f["ind_after_filter_" + tag]), tag
        n_sorted = int(dev["n_valid"][0].item())
        order = dev["order"][0, :min(n_sorted, 6000)].cpu().numpy()      # indices into all anchors
        want_order = f["ind_after_filter_" + tag][f["ind_after_sort_" + tag]]
        assert np.array_equal(order, want_order), tag
        num = int(dev["num"][0].item())
        keep = dev["keep"][0, :num].cpu().numpy()
        assert np.array_equal(keep, f["proposal_index_" + tag]), tag
        assert int(cnt[0].item()) == f["rois_" + tag].shape[0]


def test_stage_bridge_and_mask_layer_vs_reference_forward():
    _install()
    import caffe
    from pylayer.stage_bridge_layer import StageBridgeLayer
    from pylayer.mask_layer import MaskLayer
    f = load("ref_stage_bridge.npz")
    for tag in ("a", "b"):
        bottom = _blobs(f["rois_" + tag], f["bbox_pred_" + tag], f["prob_" + tag], f["im_info_" + tag])
        top = _blobs(None)
        layer = StageBridgeLayer(phase=caffe.TEST)
        layer.setup(bottom, top)
        layer.forward(bottom, top)
        want = f["rois_ext_" + tag]
        assert top[0].data.shape == want.shape and top[0].data.dtype == np.float32
        assert np.allclose(top[0].data, want, rtol=1e-5, atol=2e-3), np.abs(top[0].data - want).max()
    bottom, top = _blobs(f["mask_output"]), _blobs(None)
    layer = MaskLayer(phase=caffe.TEST)
    layer.setup(bottom, top)
    layer.forward(bottom, top)
    assert np.array_equal(top[0].data, f["mask_proposal"])


def test_bbox_transform_and_overlaps_vs_reference():
    _install()
    from transform.bbox_transform import bbox_transform_inv, clip_boxes, filter_small_boxes
    from transform.anchors import generate_anchors
    from utils.cython_bbox import bbox_overlaps
    f = load("ref_bbox.npz")
    assert np.array_equal(np.asarray(generate_anchors(), dtype=np.float64), f["anchors"])
    assert np.array_equal(bbox_overlaps(f["ov_boxes"], f["ov_boxes"][::3].copy()), f["ov"])
    for tag in ("a", "b"):
        pred = bbox_transform_inv(f["boxes_" + tag], f["deltas_" + tag])
        assert pred.shape == f["pred_" + tag].shape
        assert np.allclose(pred, f["pred_" + tag], rtol=1e-5, atol=2e-3)
        clipped, keep = clip_boxes(f["pred_" + tag], np.array([600, 1000], np.float32))
        assert np.array_equal(clipped, f["clipped_" + tag]) and np.array_equal(keep, f["clip_keep_" + tag])
        assert np.array_equal(filter_small_boxes(f["clipped_" + tag][:, :4], 16 * 1.6), f["small_keep_" + tag])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_gpu_mask_voting_vs_reference(tag):
    """Device voting pipeline against the reference's gpu_mask_voting (mask_transform.py:213-286,
    numpy-1.x evaluation of the weight sum): per-class NMS keeps, candidate lists and weights
    bit-exact; result boxes int-exact; masks 1e-4 (FMA contraction in the render)."""
    _install()
    from transform.mask_transform import gpu_mask_voting
    from mnc_b200 import ops
    f = load("ref_voting.npz")
    boxes, masks, scores, H, W = voting_case(f, tag)
    nb = boxes.shape[0]
    r = ops.mask_voting(torch.from_numpy(boxes).cuda()[None], torch.from_numpy(masks).cuda()[None],
                        torch.from_numpy(scores).cuda()[None], torch.tensor([[H, W]], dtype=torch.int32).cuda())
    assert int(r["overflow"].item()) == 0
    order = r["order"].cpu().numpy().reshape(20, nb)
    keep = r["keep"].cpu().numpy().reshape(20, -1)
    num = r["num_keep"].cpu().numpy().reshape(20)
    for c in range(1, 21):
        want = f["nms_keep_%s_c%d" % (tag, c)][:100]
        got = order[c - 1][keep[c - 1, :min(num[c - 1], 100)]]
        assert np.array_equal(got, want), c
    sfx = "_%s_np1" % tag
    k = int(r["n_res"][0].item())
    start = f["cand_start" + sfx]
    assert k == len(start)
    beg = r["cand_begin"][0, :k].cpu().numpy()
    end = r["cand_end"][0, :k].cpu().numpy()
    ci = r["cand_inds"][0].cpu().numpy().ravel()
    cw = r["cand_weights"][0].cpu().numpy().ravel()
    assert np.array_equal(np.cumsum(end - beg), start)
    assert np.array_equal(np.concatenate([ci[b:e] for b, e in zip(beg, end)]), f["cand_inds" + sfx])
    assert np.array_equal(np.concatenate([cw[b:e] for b, e in zip(beg, end)]), f["cand_weights" + sfx])
    lm, lb = gpu_mask_voting(masks, boxes, scores, 21, 100, W, H)
    assert np.array_equal(np.array([len(b) for b in lb]), f["class_counts" + sfx])
    got_box, want_box = np.vstack(lb), f["result_box" + sfx]
    assert got_box.shape == want_box.shape
    assert np.array_equal(got_box[:, 4], want_box[:, 4])            # scores
    assert np.array_equal(got_box[:, :4], want_box[:, :4]), np.abs(got_box[:, :4] - want_box[:, :4]).max()
    assert util.rel_err(np.concatenate(lm, 0), f["result_mask" + sfx]) < 1e-4


def test_prep_and_im_detect_tail_vs_reference():
    """a13: input preparation and the im_detect tail against the reference's own
    prepare_mnc_args / im_detect (tools/demo.py:54-100) on non-600x1000 images with non-unit scale."""
    from mnc_b200 import ops, weights as Wt
    from mnc_b200.engine import MNCEngine
    from oracle import oracle as O
    f = load("ref_prep_tail.npz")
    eng = MNCEngine(Wt.make_weights(Wt.TINY_ARCH), device="cuda:0")
    for tag in f["cases"]:
        im, blobs = tail_case(f, tag)
        scale = ops.im_scale_for(im.shape)
        assert scale == float(f["scale_" + tag])
        data = ops.prep_images(torch.from_numpy(im[None]).cuda(), scale).cpu().numpy()
        assert np.array_equal(np.array(data.shape), f["data_shape_" + tag])
        assert np.abs(data[0, :, ::37, ::41] - f["data_probe_" + tag]).max() < 1e-4
        assert abs(data.astype(np.float64).sum() - float(f["data_sum_" + tag])) < 1e-6 * np.abs(data).sum()
        n = blobs["rois"].shape[0]
        o = {k: torch.from_numpy(v).cuda() for k, v in blobs.items()}
        o["roi_counts"] = torch.tensor([n - 3], dtype=torch.int32).cuda()
        hw = torch.tensor([[im.shape[0], im.shape[1]]], dtype=torch.float32).cuda()
        sc = torch.tensor([scale], dtype=torch.float32).cuda()
        boxes, masks, scores, valid = eng.detect_tail(o, 1, hw, sc, n=n)
        boxes = boxes[0].cpu().numpy()
        want = f["out_boxes_" + tag]                     # float64: the reference under numpy 2
        ob, om, osc = O.im_detect_tail(blobs, im.shape, scale, numpy2=False)
        assert np.array_equal(boxes, ob), tag             # numpy-1.x evaluation: bit-exact
        assert np.all(np.abs(boxes.astype(np.float64) - want) <= np.spacing(np.maximum(np.abs(boxes), 1e-30)))
        assert np.array_equal(masks[0].cpu().numpy(), f["out_masks_" + tag])       # stage 1 then stage 2
        assert np.array_equal(scores[0].cpu().numpy(), f["out_scores_" + tag])
        v = valid[0].cpu().numpy()
        assert v.shape == (2 * n,) and v[:n - 3].all() and not v[n - 3:n].any() and v[n:2 * n - 3].all()
