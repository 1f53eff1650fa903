"""A/B of the launch shape of mask voting's two passes (coarse stride, CTAs per result in each pass)
on the bench's own detections (batch 8, FULL_ARCH, synthetic images), and of the planes-per-lane
choice of the 14x14 row-walk ROIWarping kernel.  Results must not depend on the shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mnc_b200 import ops, weights as Wt
from mnc_b200._lib import lib
from mnc_b200.engine import MNCEngine

H, W, B = 600, 1000, 8
eng = MNCEngine(Wt.make_weights(Wt.FULL_ARCH))
u8 = np.stack([np.random.default_rng(1234 + i).integers(0, 256, size=(H, W, 3), dtype=np.uint8) for i in range(B)])
data = ops.prep_images(torch.from_numpy(u8).cuda(), 1.0)
im_info = torch.tensor([[H, W, 1.0]] * B, dtype=torch.float32, device="cuda")
hw = torch.tensor([[H, W]] * B, dtype=torch.float32, device="cuda")
sc = torch.ones(B, dtype=torch.float32, device="cuda")
for _ in range(2):
    boxes, masks, scores, valid, o = eng.detect(data, im_info, hw, sc)
torch.cuda.synchronize()
boxes, masks, scores, valid = boxes.clone(), masks.clone(), scores.clone(), valid.clone()
hw_i = torch.tensor([[H, W]] * B, dtype=torch.int32, device="cuda")
res = {}
ref = None
shapes = [(4, 4, 24), (4, 2, 24), (4, 1, 24), (4, 4, 12), (4, 2, 12), (4, 2, 8), (8, 2, 24), (8, 1, 12),
          (2, 8, 24), (2, 8, 12), (3, 4, 16), (6, 2, 16)]
for stride, c1, c2 in shapes:
    assert lib.mnc_mv_set_shape(stride, c1, c2) == 0
    ms = bench.median_ms(lambda: ops.mask_voting(boxes, masks, scores, hw_i, box_valid=valid), iters=10)
    r = ops.mask_voting(boxes, masks, scores, hw_i, box_valid=valid)
    got = (r["result_box"].clone(), r["result_mask"].clone())
    if ref is None:
        ref = got
    same = bool(torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]))
    res["mv_stride%d_c%d_c%d" % (stride, c1, c2)] = {"voting_ms": round(ms, 4), "identical": same}
    print(stride, c1, c2, round(ms, 4), same, flush=True)
lib.mnc_mv_set_shape(6, 2, 16)
ops.mv_set_two_pass(False)
res["mv_full_sweep"] = {"voting_ms": round(bench.median_ms(lambda: ops.mask_voting(boxes, masks, scores, hw_i, box_valid=valid), iters=10), 4)}
ops.mv_set_two_pass(True)
print(res["mv_full_sweep"], flush=True)

# 14x14 row walk: planes per lane
hbm = bench._peaks()[0]["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cpu").manual_seed(7)
feat = torch.randn(1, 512, 38, 63, generator=g).clamp_min(0).cuda()
R = 2000
rng = np.random.default_rng(8)
x1, y1 = rng.uniform(0, 999, R), rng.uniform(0, 599, R)
w, h = rng.uniform(16, 600, R), rng.uniform(16, 600, R)
rois = torch.from_numpy(np.stack([np.zeros(R), x1, y1, np.clip(x1 + w, 0, 999), np.clip(y1 + h, 0, 599)], 1).astype(np.float32)).cuda()
out = torch.empty(R, 512, 14, 14, device="cuda")
alg = R * 512 * 196 * 4 + 512 * 38 * 63 * 4 + R * 20
ref = None
for planes, th, cpc in [(4, 128, 32), (8, 128, 64), (8, 64, 32), (8, 256, 128), (4, 256, 64)]:
    assert lib.mnc_roi_warp_set_walk_planes14(planes) == 0 and lib.mnc_roi_warp_set_walk_shape(th, cpc) == 0
    out.zero_()
    ms = bench.median_ms(lambda: ops.roi_warp_nchw(feat, rois, 14, 14, out=out), flush=flush)
    if ref is None:
        ref = out.clone()
    res["roi14_planes%d_t%d_c%d" % (planes, th, cpc)] = {"ms": round(ms, 4), "frac_of_hbm": round(alg / ms / 1e6 / hbm, 3),
                                                        "identical": bool(torch.equal(out, ref))}
    print("roi14", planes, th, cpc, res["roi14_planes%d_t%d_c%d" % (planes, th, cpc)], flush=True)
lib.mnc_roi_warp_set_walk_planes14(4)
lib.mnc_roi_warp_set_walk_shape(128, 32)
json.dump(res, open("gpurun_out/r02e_mv_shape_roi14_ab.json", "w"), indent=1)
