"""The four forms of the proposal NMS at the benchmarked size (8 images x 6000 candidates -> 300,
RPN-like clustered boxes and the engine's own proposals) -- run under ncu for per-kernel times, or
alone for CUDA-event medians."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np, torch
import bench
from mnc_b200 import ops
from tests.test_gpu_nms import _clustered_boxes

P, n = 8, 6000
boxes = np.stack([_clustered_boxes(n, 7 + p) for p in range(P)])
tb = torch.from_numpy(boxes).cuda()
res = {}
ref = None
for mode in (0, 1, 2, 3):
    ops.nms_set_lazy(mode)
    for batch in (8, 1):
        x = tb[:batch].contiguous()
        ms = bench.median_ms(lambda: ops.nms_sorted(x, None, 0.7, 300))
        res["mode%d_batch%d_ms" % (mode, batch)] = round(ms, 4)
    keep, num = ops.nms_sorted(tb, None, 0.7, 300)
    got = (keep.cpu().numpy(), num.cpu().numpy())
    if ref is None:
        ref = got
    res["mode%d_equal_to_matrix" % mode] = bool(np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]))
ops.nms_set_lazy(True)
print(json.dumps(res))
json.dump(res, open("gpurun_out/r02e_nms_modes.json", "w"), indent=1)
