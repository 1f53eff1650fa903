"""One or two launches of every non-tensor kernel of the path at its benchmarked size, for an
`ncu --set full` capture (profiles/r02_ncu_misc_*): the engine's RoI / mask / proposal kernels on a
batch-8 step, the configs[3] layer kernels (2000 RoIs) and the configs[4] NMS / voting kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mnc_b200 import ops, weights as Wt
from mnc_b200.engine import MNCEngine
from tests import util
from tests.test_ref_pin import _voting_inputs

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
torch.cuda.nvtx.range_push("profile")
boxes, masks, scores, valid, o = eng.detect(data, im_info, hw, sc)           # one eager step
hw_i = torch.tensor([[H, W]] * B, dtype=torch.int32, device="cuda")
ops.mask_voting(boxes, masks, scores, hw_i, box_valid=valid)                    # batched voting
# configs[3]
g = torch.Generator(device="cpu").manual_seed(7)
feat = torch.randn(1, 512, 38, 63, generator=g).clamp_min(0).cuda()
rng = np.random.default_rng(8)
x1, y1 = rng.uniform(0, 999, 2000), rng.uniform(0, 599, 2000)
w, h = rng.uniform(16, 600, 2000), rng.uniform(16, 600, 2000)
rois = torch.from_numpy(np.stack([np.zeros(2000), x1, y1, np.clip(x1 + w, 0, 999), np.clip(y1 + h, 0, 599)], 1).astype(np.float32)).cuda()
for P in (28, 14):
    out = torch.empty(2000, 512, P, P, device="cuda")
    ops.roi_warp_nchw(feat, rois, P, P, out=out)
    del out
f14 = torch.randn(2000, 512, 14, 14, device="cuda")
m14 = torch.rand(2000, 1, 14, 14, device="cuda")
ops.mask_pool_nchw(f14, m14, out=torch.empty_like(f14))
# configs[4]
bx = util.random_boxes(10000, seed=10)
scr = util.tie_free_scores(10000, seed=11)
order = np.argsort(-scr, kind="stable")
ops.nms_sorted(torch.from_numpy(bx[order]).cuda()[None].contiguous(), None, 0.7, 300)
vb, vm, vs = _voting_inputs(600, 600, 1000, 11)
ops.mask_voting(*(torch.from_numpy(a).cuda()[None] for a in (vb, vm, vs)),
                torch.tensor([[600, 1000]], dtype=torch.int32, device="cuda"))
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
