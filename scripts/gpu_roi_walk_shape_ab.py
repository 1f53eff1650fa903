"""A/B of the launch shape of the row-walk ROIWarping kernel (BASELINE.json configs[3]: 2000 RoIs
on a 512x38x63 map): threads per CTA x channels per CTA.  All shapes must give identical bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mnc_b200 import ops
from mnc_b200._lib import lib

hbm = bench._peaks()[0]["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cpu").manual_seed(7)
feat = torch.randn(1, 512, 38, 63, generator=g).clamp_min(0).cuda()
R = 2000
rng = np.random.default_rng(8)
x1, y1 = rng.uniform(0, 999, R), rng.uniform(0, 599, R)
w, h = rng.uniform(16, 600, R), rng.uniform(16, 600, R)
rois = np.stack([np.zeros(R), x1, y1, np.clip(x1 + w, 0, 999), np.clip(y1 + h, 0, 599)], 1).astype(np.float32)
t = torch.from_numpy(rois).cuda()
shapes = [(256, 32), (128, 32), (128, 64), (256, 64), (64, 16), (64, 32), (96, 24), (192, 48), (128, 128), (256, 128)]
res = {}
for P in (28, 14):
    out = torch.empty(R, 512, P, P, device="cuda")
    ref = None
    alg = R * 512 * P * P * 4 + 512 * 38 * 63 * 4 + R * 20
    for th, cpc in shapes:
        assert lib.mnc_roi_warp_set_walk_shape(th, cpc) == 0
        out.zero_()
        ms = bench.median_ms(lambda: ops.roi_warp_nchw(feat, t, P, P, out=out), flush=flush)
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
        res["P%d_t%d_c%d" % (P, th, cpc)] = {"ms": round(ms, 4), "frac_of_hbm": round(alg / ms / 1e6 / hbm, 3), "identical": same}
        print(P, th, cpc, round(ms, 4), round(alg / ms / 1e6 / hbm, 3), same, flush=True)
    del out, ref
lib.mnc_roi_warp_set_walk_shape(128, 32)
json.dump(res, open("gpurun_out/r02c_roi_walk_shape_ab.json", "w"), indent=1)
