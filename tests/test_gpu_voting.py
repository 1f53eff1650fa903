"""gpu_mask_voting / mv drop-ins vs the oracle."""
import numpy as np
import pytest
import torch

from tests import util
from tests.test_ref_pin import _voting_inputs

pytestmark = pytest.mark.gpu


def _install():
    import mnc_b200.lib as L
    L.install()


def test_bbox_overlaps_float64_exact():
    _install()
    from utils.cython_bbox import bbox_overlaps
    from oracle import oracle as O
    b = util.random_boxes(600, 1).astype(np.float64)
    q = util.random_boxes(7, 2).astype(np.float64)
    assert np.array_equal(bbox_overlaps(b, q), O.bbox_overlaps(b, q))
    with pytest.raises(ValueError):
        bbox_overlaps(b.astype(np.float32), q)


@pytest.mark.parametrize("nb,H,W,seed", [(120, 150, 200, 21), (600, 600, 1000, 22)])
def test_mv_host_matches_oracle(nb, H, W, seed):
    _install()
    from nms.mv import mv
    from oracle import oracle as O
    boxes, masks, scores = _voting_inputs(nb, H, W, seed)
    inds, start, weights, _, _ = O.mask_voting_candidates(boxes, scores, 21, 100)
    rm, rb = mv(boxes, masks, inds, start, weights, H, W)
    rm_o, rb_o, agg = O.mv(boxes, masks, inds, start, weights, H, W, return_agg=True)
    assert rm.shape == (len(start), 1, 21, 21) and rb.dtype == np.int32
    # int boxes match exactly: on these seeded inputs no aggregated pixel that bounds a result box
    # sits within float noise of the 0.4 threshold (checked: the nearest is reported on failure)
    assert np.array_equal(rb, rb_o), "nearest aggregate to 0.4: %.3e" % np.abs(agg - 0.4).min()
    assert util.rel_err(rm, rm_o) < 1e-4


def test_mv_empty_and_tiny():
    _install()
    from nms.mv import mv
    from oracle import oracle as O
    boxes = np.array([[10, 10, 50, 40], [12, 8, 48, 44]], dtype=np.float32)
    masks = np.full((2, 1, 21, 21), 0.1, dtype=np.float32)   # never exceeds 0.4 -> default bbox
    inds = np.array([0, 1], dtype=np.int32)
    start = np.array([2], dtype=np.int32)
    w = np.array([0.5, 0.5], dtype=np.float32)
    rm, rb = mv(boxes, masks, inds, start, w, 101, 77)
    rm_o, rb_o = O.mv(boxes, masks, inds, start, w, 101, 77)
    assert np.array_equal(rb, rb_o) and list(rb[0]) == [77 // 2, 101 // 2, 77 // 2, 101 // 2]
    assert util.rel_err(rm, rm_o) < 1e-5
    rm0, rb0 = mv(boxes, masks, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), 101, 77)
    assert rm0.shape == (0, 1, 21, 21) and rb0.shape == (0, 4)


@pytest.mark.parametrize("nb,H,W,seed", [(600, 600, 1000, 31), (245, 224, 320, 32)])
def test_gpu_mask_voting_matches_oracle(nb, H, W, seed):
    _install()
    from transform.mask_transform import gpu_mask_voting
    from mnc_b200 import ops
    from oracle import oracle as O
    boxes, masks, scores = _voting_inputs(nb, H, W, seed)
    # --- candidate lists: bit-exact
    r = ops.mask_voting(torch.from_numpy(boxes).cuda()[None], torch.from_numpy(masks).cuda()[None],
                        torch.from_numpy(scores).cuda()[None],
                        torch.tensor([[H, W]], dtype=torch.int32).cuda())
    inds, start, weights, cscores, class_bar = O.mask_voting_candidates(boxes, scores, 21, 100)
    k = int(r["n_res"][0].item())
    assert k == len(start)
    assert np.array_equal(r["class_bar"][0].cpu().numpy(), np.asarray(class_bar))
    assert np.array_equal(r["res_score"][0, :k].cpu().numpy(), cscores)
    beg = r["cand_begin"][0, :k].cpu().numpy()
    end = r["cand_end"][0, :k].cpu().numpy()
    ci = r["cand_inds"][0].cpu().numpy().ravel()
    cw = r["cand_weights"][0].cpu().numpy().ravel()
    got_inds = np.concatenate([ci[b:e] for b, e in zip(beg, end)])
    got_w = np.concatenate([cw[b:e] for b, e in zip(beg, end)])
    assert np.array_equal(np.cumsum(end - beg), start)
    assert np.array_equal(got_inds, inds)
    assert np.array_equal(got_w, weights)
    # --- public function: same structure and values as the oracle's
    lm, lb = gpu_mask_voting(masks, boxes, scores, 21, 100, W, H)
    lm_o, lb_o = O.gpu_mask_voting(masks, boxes, scores, 21, 100, W, H)
    assert len(lm) == len(lb) == 20
    for c in range(20):
        assert lb[c].shape == lb_o[c].shape and lm[c].shape == lm_o[c].shape
        if lb[c].shape[0]:
            assert np.array_equal(lb[c][:, 4], lb_o[c][:, 4])
            assert np.array_equal(lb[c][:, :4], lb_o[c][:, :4])
            assert util.rel_err(lm[c], lm_o[c]) < 1e-3


@pytest.mark.parametrize("unit_range", [True, False])
def test_mv_two_pass_equals_full_sweep(unit_range):
    """mnc_mv_device's coarse pass + exact border pass (and, for masks in [0,1], the search region
    cut down to the columns / rows whose covering weight can reach 0.4) gives the same tight boxes
    and voted masks as one full sweep of the candidates' union region -- batched, blob-shaped and
    noise masks, one image with nothing above the threshold."""
    from mnc_b200 import ops
    B, nb, H, W = 3, 300, 375, 500
    rng = np.random.default_rng(77)
    boxes = np.zeros((B, nb, 4), np.float32)
    masks = np.zeros((B, nb, 1, 21, 21), np.float32)
    scores = np.zeros((B, nb, 21), np.float32)
    yy, xx = np.mgrid[0:21, 0:21]
    for b in range(B):
        bx, mk, sc = _voting_inputs(nb, H, W, 90 + b)
        # clustered boxes (jittered copies of 12 objects) so that voting lists are long
        src = rng.integers(0, 12, nb)
        bx = bx[src] + rng.normal(0, 4, (nb, 4)).astype(np.float32)
        bx[:, 0::2] = np.clip(np.sort(bx[:, 0::2], axis=1), 0, W - 1)
        bx[:, 1::2] = np.clip(np.sort(bx[:, 1::2], axis=1), 0, H - 1)
        if b == 0:      # blob masks: high in a centred ellipse, low outside
            r = rng.uniform(5, 9, nb)[:, None, None]
            mk = (1.0 / (1.0 + np.exp(((xx - 10) ** 2 + (yy - 10) ** 2) ** 0.5 - r)))[:, None].astype(np.float32)
        elif b == 2:    # nothing ever exceeds the threshold
            mk = np.full_like(mk, 0.2)
        if not unit_range:
            mk = mk * 1.3 - 0.15                     # values outside [0,1]: full-sum predicate
        boxes[b], masks[b], scores[b] = bx, mk, sc
    args = (torch.from_numpy(boxes).cuda(), torch.from_numpy(masks).cuda(),
            torch.from_numpy(scores).cuda(),
            torch.tensor([[H, W]] * B, dtype=torch.int32).cuda())
    prev = ops.mv_set_two_pass(True)
    try:
        r2 = ops.mask_voting(*args)
        ops.mv_set_two_pass(False)
        r1 = ops.mask_voting(*args)
    finally:
        ops.mv_set_two_pass(prev)
    n = r1["n_res"].cpu().numpy()
    assert np.array_equal(n, r2["n_res"].cpu().numpy()) and n.min() > 5
    assert torch.equal(r1["result_box"], r2["result_box"])
    assert torch.equal(r1["result_mask"], r2["result_mask"])
    rb = r1["result_box"][0, :n[0]].cpu().numpy()
    assert (rb[:, 2] > rb[:, 0]).any()                # blob image: real (non-default) boxes
