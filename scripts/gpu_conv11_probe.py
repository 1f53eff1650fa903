"""conv1_1: tensor-core form vs fp32 FMA form (batch 8, 600x1000, L2 flushed between launches).
(The stage-elimination numbers in profiles/README.md finding 10 came from two temporary switches in
mnc_conv1_1_tc -- epilogue drains and discards / producers skip the gathers -- since removed.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mnc_b200 import dense

B, H, W = 8, 600, 1000
data = (torch.rand(B, 3, H, W, device="cuda") * 255 - 115).contiguous()
w = torch.randn(64, 3, 3, 3, device="cuda") * 0.02
b = torch.randn(64, device="cuda")
out = torch.zeros(2, B, H, W, 64, device="cuda", dtype=torch.bfloat16)
wt = dense.conv1_1_weight_to_tc(w)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def t(fn, n=6):
    ts = []
    for it in range(n + 2):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


print("conv1_1: tensor cores %.3f ms   fp32 FMA %.3f ms" % (
    t(lambda: dense.conv1_1_tc(data, wt, b, out)), t(lambda: dense.conv1_1(data, w, b, out))))
