"""Why does conv1_2 take ~1.15 ms inside the step but 0.86 ms in a back-to-back loop?
Variants: input statistics (N(0,1) vs post-ReLU), L2 flushed between launches, fresh vs reused
output buffer, per-launch events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mnc_b200 import dense

dev = "cuda"
B, H, W, C = 8, 600, 1000, 64
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ws = dense.conv_weight_to_split(torch.randn(C, C, 3, 3, device=dev) / (9 * C) ** 0.5)
bias = torch.zeros(C, device=dev)


def bench(x, pool, do_flush, iters=8):
    Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if pool else (H, W)
    out = torch.zeros(2, B, Ho, Wo, C, device=dev, dtype=torch.bfloat16)
    ts = []
    for it in range(iters + 2):
        if do_flush:
            flush.fill_(it & 255)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        dense.igemm(x, B, H, W, C, ws, C, 9, bias=bias, relu=True, out=out, pool=pool)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


x_n = dense.split(torch.randn(B, H, W, C, device=dev))
x_r = dense.split(torch.relu(torch.randn(B, H, W, C, device=dev)))
x_z = dense.split(torch.zeros(B, H, W, C, device=dev))
x_s = dense.split(torch.relu(torch.randn(B, H, W, C, device=dev)) * 30.0)
for name, x in (("N(0,1)", x_n), ("relu(N(0,1))", x_r), ("zeros", x_z), ("30*relu", x_s)):
    for pool in (False, True):
        for fl in (False, True):
            print("input %-13s pool=%d flush=%d : %.3f ms" % (name, pool, fl, bench(x, pool, fl)), flush=True)
