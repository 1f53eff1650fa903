"""End-to-end parity of the fused engine against the oracle.

From pixels, a 1-ulp score difference can legally reorder near-ties, so "bit-exact RoI indices"
is asserted per stage on the engine's own inputs to that stage (teacher forcing), and values are
compared at the 1e-3 relative fp32 tolerance of north_star on everything the oracle and the engine
compute from identical inputs."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-3  # BASELINE.json north_star: "within 1e-3 relative fp32"


def _run(arch_name, H, W, B):
    from oracle import oracle as O
    from mnc_b200 import weights as Wt
    from mnc_b200.engine import MNCEngine
    arch = getattr(Wt, arch_name)
    w = Wt.make_weights(arch)
    ims = [O.synthetic_image(i, H, W) for i in range(B)]
    blobs = [O.prep_blob(im) for im in ims]
    data = np.concatenate([b[0] for b in blobs])
    im_info = np.concatenate([b[1] for b in blobs])
    eng = MNCEngine(w)
    out = eng.forward(torch.from_numpy(data).cuda(), torch.from_numpy(im_info).cuda(), keep_intermediate=True)
    torch.cuda.synchronize()
    return w, ims, data, im_info, eng, out


def _check_stagewise(w, data, im_info, eng, out, img, n_per=300):
    """Everything below feeds the ORACLE with the ENGINE's inputs to that stage."""
    import torch.nn.functional as F
    from oracle import oracle as O
    from mnc_b200 import dense
    sl = slice(img * n_per, (img + 1) * n_per)
    n = int(out["roi_counts"][img].item())
    # ---- trunk: conv5_3 vs torch CPU fp32 from pixels
    c5 = dense.merge(out["_conv5_3"])[img].permute(2, 0, 1).cpu().numpy()
    with torch.no_grad():
        want_c5 = O.trunk_forward(w, data[img:img + 1])[0].numpy()
    assert util.rel_err(c5, want_c5) < TOL
    # ---- RPN head on the engine's conv5_3
    with torch.no_grad():
        prob, bbox = O.rpn_forward(w, torch.from_numpy(c5[None]))
    rpn = out["_rpn_out"][img].cpu().numpy()                       # (H5, W5, 64) logits | deltas
    assert util.rel_err(rpn[..., 18:54].transpose(2, 0, 1), bbox[0].numpy()) < TOL
    # ---- proposals: exact indices given the engine's decoded boxes / scores
    p = out["_proposal"]
    props, scores, valid = (p[k][img].cpu().numpy() for k in ("proposals", "scores", "valid"))
    want_scores = prob[0, 9:].numpy().transpose(1, 2, 0).ravel()
    assert np.abs(scores - want_scores).max() < TOL
    from tests.test_gpu_proposal import _oracle_from_device_decode
    forced_boxes, _ = _oracle_from_device_decode(props, scores, valid)
    rois = out["rois"][sl].cpu().numpy()
    assert n == forced_boxes.shape[0]
    assert np.array_equal(rois[:n, 1:], forced_boxes)
    assert np.all(rois[:n, 0] == img)
    # ---- stage 1 RoI features on the engine's rois + conv5_3
    rois0 = rois[:n].copy()
    rois0[:, 0] = 0
    want14 = F.max_pool2d(torch.from_numpy(O.roi_warp(c5[None], rois0, 28, 28)), 2, 2).numpy()
    got14 = dense.merge(out["_feat14"])[sl][:n].permute(0, 3, 1, 2).cpu().numpy()
    assert util.rel_err(got14, want14) < 1e-4
    # ---- stage 1 head on the engine's features
    with torch.no_grad():
        h = O.head_forward(w, got14)
    logit_scale = max(1.0, np.abs(h["mask_pred"]).max())
    assert np.abs(out["_mask_logits"][sl][:n, :441].cpu().numpy() - h["mask_pred"]).max() < TOL * logit_scale
    assert np.abs(out["mask_proposal"][sl][:n].cpu().numpy() - h["mask_proposal"]).max() < TOL
    assert np.abs(out["seg_cls_prob"][sl][:n].cpu().numpy() - h["seg_cls_prob"]).max() < TOL
    assert np.abs(out["cls_prob"][sl][:n].cpu().numpy() - h["cls_prob"]).max() < TOL
    bb = out["bbox_pred"][sl][:n].cpu().numpy()
    assert np.abs(bb - h["bbox_pred"]).max() < TOL * max(1.0, np.abs(h["bbox_pred"]).max())
    # ---- stage bridge on the engine's own bbox_pred / seg_cls_prob: argmax exact, coords 1e-3
    want_ext = O.stage_bridge_forward(rois0, bb, out["seg_cls_prob"][sl][:n].cpu().numpy(), im_info[img:img + 1])
    got_ext = out["rois_ext"][sl][:n].cpu().numpy()
    assert np.abs(got_ext[:, 1:] - want_ext[:, 1:]).max() <= TOL * max(1.0, np.abs(want_ext).max())
    # ---- stage 2 features + head
    ext0 = got_ext.copy()
    ext0[:, 0] = 0
    want14e = O.roi_warp(c5[None], ext0, 14, 14)
    got14e = dense.merge(out["_feat14_ext"])[sl][:n].permute(0, 3, 1, 2).cpu().numpy()
    assert util.rel_err(got14e, want14e) < 1e-4
    with torch.no_grad():
        h2 = O.head_forward(w, got14e)
    assert np.abs(out["mask_proposal_ext"][sl][:n].cpu().numpy() - h2["mask_proposal"]).max() < TOL
    assert np.abs(out["seg_cls_prob_ext"][sl][:n].cpu().numpy() - h2["seg_cls_prob"]).max() < TOL
    return n


def test_tiny_arch_batch3_stagewise():
    w, ims, data, im_info, eng, out = _run("TINY_ARCH", 224, 320, 3)
    for img in range(3):
        n = _check_stagewise(w, data, im_info, eng, out, img)
        assert n > 50


def _match_rois(got, want, tol=0.05):
    """Pair RoIs of two lists by coordinates (different anchors are many pixels apart): returns
    index pairs (i, j) with max |got[i] - want[j]| < tol, each row used once."""
    d = np.abs(got[:, None, 1:] - want[None, :, 1:]).max(axis=2)
    j = d.argmin(axis=1)
    ok = d[np.arange(len(got)), j] < tol
    i = np.where(ok)[0]
    assert len(set(j[i])) == len(i)
    return i, j[i]


@pytest.mark.parametrize("arch,H,W", [("TINY_ARCH", 224, 320), ("TINY_ARCH", 375, 500)])
def test_pixels_to_outputs_free_running(arch, H, W):
    """No teacher forcing: pixels -> every blob the callers read, engine vs pure oracle run.
    A 1e-6 score or 1e-4 px box difference may flip an NMS decision or swap two near-equal scores,
    after which a few RoIs differ between the two runs (legal); so RoIs are paired by coordinates,
    the unpaired remainder is bounded, and every paired RoI's outputs must agree within tolerance
    through BOTH stages.  A wrong kernel anywhere fails this."""
    from oracle import oracle as O
    w, ims, data, im_info, eng, out = _run(arch, H, W, 1)
    (boxes, masks, scores), blobs = O.im_detect(w, ims[0])
    n = int(out["roi_counts"][0].item())
    want_n = blobs["rois"].shape[0]
    assert abs(n - want_n) <= 3, (n, want_n)
    rois = out["rois"][:n].cpu().numpy()
    i, j = _match_rois(rois, blobs["rois"])
    assert len(i) >= want_n - 6, "only %d of %d RoIs found in the oracle's list" % (len(i), want_n)
    # rank agreement: paired RoIs appear in the same relative order, except where two proposals
    # whose ORACLE scores lie within the engine's score tolerance swap places (random-init RPN
    # scores of neighbouring proposals are ~1e-5 apart); the swaps are few
    inv = np.where(np.diff(j) <= 0)[0]
    assert len(inv) <= 12, inv
    if len(inv):
        _, mid = O.proposal_layer_forward(blobs["rpn_cls_prob_reshape"], blobs["rpn_bbox_pred"],
                                          np.asarray(im_info)[:1], return_intermediate=True)
        sc = mid["sorted_scores"][mid["nms_keep"]]          # score of every oracle RoI, descending
        assert len(sc) == want_n
        gap = np.abs(sc[j[inv]] - sc[j[inv + 1]])
        assert gap.max() < 0.2 * TOL, (inv, gap)
    for name, tol in (("mask_proposal", 5 * TOL), ("seg_cls_prob", 5 * TOL), ("cls_prob", 5 * TOL)):
        g = out[name][:n].cpu().numpy()[i]
        assert np.abs(g - blobs[name][j]).max() < tol, name
    bb = out["bbox_pred"][:n].cpu().numpy()[i]
    assert np.abs(bb - blobs["bbox_pred"][j]).max() < 5 * TOL * max(1.0, np.abs(blobs["bbox_pred"]).max())
    # stage 2: rois_ext and the _ext outputs of the paired rows (argmax class may only differ
    # where the top-2 seg_cls_prob are within tolerance of each other)
    ext = out["rois_ext"][:n].cpu().numpy()[i]
    want_ext = blobs["rois_ext"][j]
    p = np.sort(blobs["seg_cls_prob"][j], axis=1)
    decided = (p[:, -1] - p[:, -2]) > 20 * TOL
    assert decided.sum() >= 0.8 * len(j)
    assert np.abs(ext[decided, 1:] - want_ext[decided, 1:]).max() < 0.05
    for name in ("mask_proposal_ext", "seg_cls_prob_ext"):
        g = out[name][:n].cpu().numpy()[i][decided]
        assert np.abs(g - blobs[name][j][decided]).max() < 10 * TOL, name


def test_full_vgg16_600x1000_stagewise_and_roi_count():
    """BASELINE.json configs[0]/[1] shape: one synthetic 600x1000 image, VGG-16 random init,
    300 RoIs per stage."""
    w, ims, data, im_info, eng, out = _run("FULL_ARCH", 600, 1000, 1)
    n = _check_stagewise(w, data, im_info, eng, out, 0)
    assert n == 300
    # equal fp32 scores do occur among 21546 anchors (about 10 pairs in the top 6000 here); their
    # order is fixed by the documented rule (score desc, index asc), which the teacher-forced
    # proposal check above has just verified index for index.


def test_full_vgg16_600x1000_batch8_stagewise():
    """BASELINE.json configs[1] exactly: FULL_ARCH, batch 8, 600x1000 -- the benchmarked
    configuration (tile / wave / split-K choices depend on the batch).  Stage-wise parity of three
    of the eight images (first, middle, last: the CPU oracle needs ~10 s per image)."""
    w, ims, data, im_info, eng, out = _run("FULL_ARCH", 600, 1000, 8)
    counts = out["roi_counts"].cpu().numpy()
    assert counts.shape == (8,) and np.all(counts == 300)
    for img in (0, 3, 7):
        assert _check_stagewise(w, data, im_info, eng, out, img) == 300


def test_caffe_net_shim_and_detect_tail():
    import mnc_b200.lib as L
    L.install()
    import caffe
    from mnc_b200 import weights as Wt
    from oracle import oracle as O
    from transform.bbox_transform import clip_boxes
    caffe.set_mode_gpu()
    caffe.set_device(0)
    w = Wt.make_weights(Wt.TINY_ARCH)
    net = caffe.Net(None, w, caffe.TEST)
    im = O.synthetic_image(0, 224, 320)
    blob, im_info = O.prep_blob(im)
    net.blobs["data"].reshape(*blob.shape)
    net.blobs["im_info"].reshape(*im_info.shape)
    net.forward(data=blob.astype(np.float32, copy=False), im_info=im_info)
    rois = net.blobs["rois"].data.copy()
    n = rois.shape[0]
    assert rois.shape[1] == 5 and np.all(rois[:, 0] == 0)
    assert net.blobs["mask_proposal"].data.shape == (n, 1, 21, 21)
    assert net.blobs["seg_cls_prob"].data.shape == (n, 21)
    assert net.blobs["rois_ext"].data.shape == (n, 5)
    assert net.blobs["mask_proposal_ext"].data.shape == (n, 1, 21, 21)
    assert net.blobs["conv5_3"].data.shape[1] == 64
    with pytest.raises(Exception):
        net.forward(data=blob)  # missing im_info (pycaffe.py:96-97)
    with pytest.raises(RuntimeError):
        caffe.set_mode_cpu()
    b, _ = clip_boxes(rois[:, 1:5] / 1.0, im.shape)
    assert b.shape == (n, 4)
