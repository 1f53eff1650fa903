"""Diagnostic: how much of a conv launch is epilogue?  out_mode 3 drains TMEM and discards."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnc_b200 import dense
from mnc_b200._lib import lib, ptr, cur_stream, check, c_int, c_ll

dev = "cuda"


def run(B, H, W, Cin, Cout, bn, mode, iters=10):
    xs = dense.split(torch.randn(B, H, W, Cin, device=dev))
    ws = dense.conv_weight_to_split(torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5)
    out = torch.zeros(2, B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(Cout, device=dev)

    def call():
        check(lib.mnc_igemm_tc(ptr(xs[0]), ptr(xs[1]), c_int(B), c_int(H), c_int(W), c_int(Cin),
                               ptr(ws[0]), ptr(ws[1]), c_int(Cout), c_int(9), ptr(bias), c_int(1),
                               c_int(mode), ptr(out[0]), ptr(out[1]), c_ll(Cout), c_int(0), c_int(1),
                               c_ll(0), c_int(bn), c_int(0), cur_stream()), "igemm")
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * H * W * Cout * 9 * Cin
    print("conv B=%d %dx%d %d->%d bn=%d mode=%d: %.3f ms %.1f TF/s" % (B, H, W, Cin, Cout, bn, mode, ms, fl / ms / 1e9), flush=True)


for shape in [(8, 600, 1000, 64, 64, 64), (8, 300, 500, 64, 128, 128), (8, 300, 500, 128, 128, 128),
              (8, 150, 250, 256, 256, 256), (8, 38, 63, 512, 512, 256)]:
    for mode in (0, 2, 3):
        run(*shape, mode)
