"""Pin the oracle AND the CUDA path against the reference's own kernels: lib/nms/nms_kernel.cu and
lib/nms/mv_kernel.cu compiled unmodified into oracle/_ref/libmnc_ref.so (oracle/Makefile `ref`).
The .so is built in the build container and travels with the snapshot; /root/reference itself is
not needed at run time."""
import ctypes
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                      "libmnc_ref.so")


def _ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libmnc_ref.so not built (needs /root/reference at build time)")
    return ctypes.CDLL(REF_SO)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("n,thresh,seed", [(600, 0.3, 1), (6000, 0.7, 2), (10000, 0.7, 10)])
def test_nms_three_way(n, thresh, seed):
    """reference `_nms` == oracle orc_nms == mnc_nms_host, float boxes, no margin nudging."""
    from oracle import oracle as O
    from mnc_b200._lib import lib, check
    ref = _ref()
    boxes = util.random_boxes(n, seed=seed)
    scores = util.tie_free_scores(n, seed=seed + 1)
    dets = np.hstack([boxes, scores[:, None]]).astype(np.float32)
    sorted_dets = np.ascontiguousarray(dets[O.order_desc(scores)])
    keep_ref = np.zeros(n, dtype=np.int32)
    num_ref = ctypes.c_int(0)
    ref._Z4_nmsPiS_PKfiifi(_p(keep_ref), ctypes.byref(num_ref), _p(sorted_dets), n, 5,
                           ctypes.c_float(thresh), 0)
    keep_ref = keep_ref[:num_ref.value]
    keep = np.zeros(n, dtype=np.int32)
    num = ctypes.c_int(0)
    check(lib.mnc_nms_host(_p(keep), ctypes.byref(num), _p(sorted_dets), n, 5,
                           ctypes.c_float(thresh), 0), "mnc_nms_host")
    assert np.array_equal(keep[:num.value], keep_ref), "CUDA path differs from reference _nms"
    assert np.array_equal(O.nms_sorted(sorted_dets, thresh), keep_ref), "oracle differs from reference _nms"


def _voting_inputs(nb, H, W, seed):
    rng = np.random.default_rng(seed)
    boxes = util.random_boxes(nb, seed=seed, width=W, height=H, smin=12, smax=min(H, W) * 0.8)
    masks = (1.0 / (1.0 + np.exp(-rng.normal(0, 2, size=(nb, 1, 21, 21))))).astype(np.float32)
    logits = rng.normal(0, 1, size=(nb, 21))
    scores = (np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(np.float32)
    return boxes, masks, scores


def test_mv_three_way():
    """reference `_mv` vs oracle orc_mv vs mnc_mv_host on gpu_mask_voting's candidate lists
    (small image so the reference's nb*H*W render buffer stays small)."""
    from oracle import oracle as O
    from mnc_b200._lib import lib, check
    ref = _ref()
    nb, H, W = 120, 150, 200
    boxes, masks, scores = _voting_inputs(nb, H, W, seed=21)
    inds, start, weights, cscores, _ = O.mask_voting_candidates(boxes, scores, 21, 100)
    k = len(start)
    assert k > 10
    rm_ref = np.zeros((k, 1, 21, 21), dtype=np.float32)
    rb_ref = np.zeros((k, 4), dtype=np.int32)
    ref._Z3_mvPKfS0_iPKiS2_S0_iiiiiiPfPii(_p(boxes), _p(masks), nb, _p(inds), _p(start), _p(weights),
                                          len(inds), H, W, 4, 21, k, _p(rm_ref), _p(rb_ref), 0)
    rm_o, rb_o = O.mv(boxes, masks, inds, start, weights, H, W)
    rm = np.zeros_like(rm_ref)
    rb = np.zeros_like(rb_ref)
    check(lib.mnc_mv_host(_p(boxes), _p(masks), nb, _p(inds), _p(start), _p(weights), len(inds), H,
                          W, 4, 21, k, _p(rm), _p(rb), 0), "mnc_mv_host")
    # boxes: int-exact unless an aggregate value sits within an ulp of 0.4 (FMA contraction);
    # allow no mismatch on this seeded input
    assert np.array_equal(rb_o, rb_ref), "oracle boxes differ from reference _mv"
    assert np.array_equal(rb, rb_ref), "CUDA boxes differ from reference _mv"
    # mask values: the reference binary contracts a*b+c into FMA, the C oracle does not
    assert util.rel_err(rm_o, rm_ref) < 1e-4
    assert util.rel_err(rm, rm_ref) < 1e-4
