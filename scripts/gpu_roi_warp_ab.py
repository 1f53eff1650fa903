"""A/B of the two fused RoI-warp kernels on the engine's own proposals (batch 8, 600x1000):
column-walking (separable, register-cached taps) vs 4-tap gather per sample."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mnc_b200 import weights as Wt, ops, dense
from mnc_b200._lib import lib
from mnc_b200.engine import MNCEngine

B, H, W = 8, 600, 1000
eng = MNCEngine(Wt.make_weights(Wt.FULL_ARCH))
imgs = []
for i in range(B):
    rng = np.random.default_rng(1234 + i)
    im = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8).astype(np.float32)
    im -= np.array([[[102.9801, 115.9465, 122.7717]]], dtype=np.float32)
    imgs.append(im.transpose(2, 0, 1))
data = torch.from_numpy(np.stack(imgs)).cuda()
info = torch.tensor([[H, W, 1.0]] * B, dtype=torch.float32).cuda()
out = eng.forward(data, info)
torch.cuda.synchronize()
c5f = eng._buf["conv5_f32"][:B * 38 * 63 * 512].view(B, 38, 63, 512)
R = B * 300
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
res = {}
for name, rois, sub in (("stage1 (28x28 -> 14, 7)", out["rois"], 2), ("stage2 (14x14 -> 7)", out["rois_ext"], 1)):
    w = (rois[:, 3] - rois[:, 1]) / 16.0
    print("%s: RoI width in feature px: mean %.1f  median %.1f  max %.1f" % (
        name, w.mean().item(), w.median().item(), w.max().item()))
    outs = {}
    for walk in (0, 1):
        lib.mnc_roi_warp_set_walk(walk)
        f14 = torch.zeros((2, R, 14, 14, 512), dtype=torch.bfloat16, device="cuda")
        b7 = torch.zeros((2, R, 7, 7, 512), dtype=torch.bfloat16, device="cuda")
        ts = []
        for it in range(6):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ops.roi_warp_split(c5f, 512, 38, 63, rois, sub, f14, b7)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        outs[walk] = (dense.merge(f14), dense.merge(b7))
        print("   walk=%d: %.3f ms (median of 5 after warm-up)" % (walk, float(np.median(ts[1:]))))
    d14 = (outs[0][0] - outs[1][0]).abs().max().item() / max(outs[0][0].abs().max().item(), 1e-30)
    d7 = (outs[0][1] - outs[1][1]).abs().max().item() / max(outs[0][1].abs().max().item(), 1e-30)
    print("   max |difference| between the two kernels, relative to max |value|: %.2e (14x14), %.2e (7x7)" % (d14, d7))
lib.mnc_roi_warp_set_walk(1)
