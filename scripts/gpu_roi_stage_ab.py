"""A/B of the ROIWarping layer kernel (BASELINE.json configs[3]): shared-memory staged window vs
the per-tap gather kernel, 2000 RoIs on a 512x38x63 map, plus the MNC-sized 300-RoI case."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mnc_b200 import ops
from mnc_b200._lib import lib

hbm = bench._peaks()[0]["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cpu").manual_seed(7)
feat = torch.randn(1, 512, 38, 63, generator=g).clamp_min(0).cuda()
for R in (2000, 300):
    rng = np.random.default_rng(8)
    x1, y1 = rng.uniform(0, 999, R), rng.uniform(0, 599, R)
    w, h = rng.uniform(16, 600, R), rng.uniform(16, 600, R)
    rois = np.stack([np.zeros(R), x1, y1, np.clip(x1 + w, 0, 999), np.clip(y1 + h, 0, 599)], 1).astype(np.float32)
    t = torch.from_numpy(rois).cuda()
    for P in (28, 14):
        out = torch.empty(R, 512, P, P, device="cuda")
        res = {}
        outs = {}
        for stage in (2, 1, 0):
            lib.mnc_roi_warp_set_stage(stage)
            ms = bench.median_ms(lambda: ops.roi_warp_nchw(feat, t, P, P, out=out), flush=flush)
            alg = R * 512 * P * P * 4 + 512 * 38 * 63 * 4 + R * 20
            res[{0: "gather", 1: "stage", 2: "rowwalk"}[stage]] = (round(ms, 4), round(alg / ms / 1e6 / hbm, 3))
            outs[stage] = out.clone()
        lib.mnc_roi_warp_set_stage(1)
        print("R=%d P=%d" % (R, P), res, "identical:", bool(torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])), flush=True)
        del out, outs
