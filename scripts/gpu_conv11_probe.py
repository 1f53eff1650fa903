"""conv1_1: tensor-core form vs fp32 FMA form, and which stage of the TC kernel bounds it
(MNC_C11_MODE=3: epilogue drains TMEM and discards; MNC_C11_NOLOAD=1: producers skip the gathers)."""
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


print("MODE=%s NOLOAD=%s : tc %.3f ms   simt %.3f ms" % (
    os.environ.get("MNC_C11_MODE", "0"), os.environ.get("MNC_C11_NOLOAD", "0"),
    t(lambda: dense.conv1_1_tc(data, wt, b, out)), t(lambda: dense.conv1_1(data, w, b, out))))
