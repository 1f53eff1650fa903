"""A/B of the fused engine RoI kernel (ROIWarping + 2x2 pools into the FC operand buffers): row walk
vs per-cell gathers, batch-8 sizes (2400 RoIs on a 8x38x63x512 NHWC map), tri-plane outputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mnc_b200 import ops, dense
from mnc_b200._lib import lib

torch.manual_seed(0)
B, H, W, C, R = 8, 38, 63, 512, 2400
feat = torch.relu(torch.randn(B, H, W, C, device="cuda"))
rng = np.random.default_rng(3)
x1, y1 = rng.uniform(0, 900, R), rng.uniform(0, 500, R)
w, h = np.exp(rng.uniform(np.log(32), np.log(500), R)), np.exp(rng.uniform(np.log(32), np.log(400), R))
rois = torch.from_numpy(np.stack([np.repeat(np.arange(B), R // B), x1, y1, np.clip(x1 + w, 0, 999), np.clip(y1 + h, 0, 599)], 1).astype(np.float32)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for sub in (2, 1):
    res, outs = {}, {}
    for rows in (2, 1, 0):
        lib.mnc_roi_warp_set_rows(rows)
        o14, o7 = dense.tri_alloc((R, 14, 14, C), "cuda"), dense.tri_alloc((R, 7, 7, C), "cuda")
        ms = bench.median_ms(lambda: ops.roi_warp_tri(feat, C, H, W, rois, sub, o14, o7, 9), flush=flush)
        res["rows%d" % rows if rows else "gather"] = round(ms, 4)
        outs[rows] = (o14.float(), o7.float())
    lib.mnc_roi_warp_set_rows(0)
    d14 = (outs[0][0] - outs[1][0]).abs().max().item()
    d7 = (outs[0][1] - outs[1][1]).abs().max().item()
    print("sub=%d" % sub, res, "max |diff| 14x14 %.3g 7x7 %.3g (max value %.3g)" % (d14, d7, outs[0][0].max().item()), flush=True)
