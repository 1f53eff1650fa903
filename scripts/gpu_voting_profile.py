"""Batched gpu_mask_voting on real engine outputs (batch 8, 600x1000): one warm-up + one profiled call
inside an NVTX range (for `ncu --nvtx --nvtx-include "vote/"`), plus CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mnc_b200 import weights as Wt, ops
from mnc_b200.engine import MNCEngine

B, H, W = 8, 600, 1000
w = Wt.make_weights(Wt.FULL_ARCH)
eng = MNCEngine(w)
imgs = []
for i in range(B):
    rng = np.random.default_rng(1234 + i)
    im = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8).astype(np.float32)
    im -= np.array([[[102.9801, 115.9465, 122.7717]]], dtype=np.float32)
    imgs.append(im.transpose(2, 0, 1))
data = torch.from_numpy(np.stack(imgs)).cuda()
info = torch.tensor([[H, W, 1.0]] * B, dtype=torch.float32).cuda()
hw = torch.tensor([[H, W]] * B, dtype=torch.float32).cuda()
sc = torch.ones(B).cuda()
boxes, masks, scores, valid, _ = eng.detect(data, info, hw, sc)
hwi = torch.tensor([[H, W]] * B, dtype=torch.int32).cuda()
for _ in range(2):
    r = ops.mask_voting(boxes, masks, scores, hwi, box_valid=valid)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
torch.cuda.nvtx.range_push("vote")
e0.record()
r = ops.mask_voting(boxes, masks, scores, hwi, box_valid=valid)
e1.record()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("voting batch 8: %.3f ms; results per image %s; candidates per result (img 0): %s" % (
    e0.elapsed_time(e1), r["n_res"].cpu().tolist(),
    (r["cand_end"][0] - r["cand_begin"][0])[:int(r["n_res"][0])].cpu().tolist()[:20]))

# ---- where does forward+voting time go?  (device time vs host launch time)
import time
def timed(fn, n=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, t_host * 1e3 / n

def fwd():
    return eng.detect(data, info, hw, sc)

def vote():
    return ops.mask_voting(boxes, masks, scores, hwi, box_valid=valid)

def both():
    b_, m_, s_, v_, _ = eng.detect(data, info, hw, sc)
    return ops.mask_voting(b_, m_, s_, hwi, box_valid=v_)

for name, fn in (("forward", fwd), ("voting x10 back-to-back", vote), ("forward+voting", both)):
    fn()
    dev_ms, host_ms = timed(fn)
    print("%-28s device %.3f ms/iter   host launch %.3f ms/iter" % (name, dev_ms, host_ms))
