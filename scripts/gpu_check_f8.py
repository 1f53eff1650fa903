"""Bring-up / numerics check of precision mode 1 (fp16 main product + two FP8 correction products,
csrc/igemm_tc.cu) on the layer shapes of the path, against fp64, next to the split-bf16 mode."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mnc_b200 import dense


def err(got, ref):
    return ((got.double() - ref).abs().max() / ref.abs().max()).item()


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def fc(M, N, K, bn=0, amp=1.0):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.relu(torch.randn(M, K, device="cuda", generator=g)) * amp
    w = torch.randn(N, K, device="cuda", generator=g) * (2.0 / K) ** 0.5
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    ref = x.double() @ w.double().T + b.double()
    out0 = torch.empty(M, N, device="cuda")
    xs, ws = dense.split(x), dense.split(w)
    f0 = lambda: dense.igemm2(xs.view(2, 1, 1, M, K), 1, 1, M, K, ws, N, 1, bias=b, out_f32=out0, bn=bn)
    f0()
    e0 = err(out0, ref)
    xt, wt = dense.tri_from_f32(x), dense.tri_from_f32(w, weight=True)
    out1 = torch.empty(M, N, device="cuda")
    amax = torch.zeros(1, dtype=torch.int32, device="cuda")
    f1 = lambda: dense.igemm2(xt.view(1, 1, M, K), 1, 1, M, K, wt, N, 1, bias=b, out_f32=out1, bn=bn, amax=amax)
    f1()
    torch.cuda.synchronize()
    e1 = err(out1, ref)
    am = amax.view(torch.float32).item()
    if N % 16:      # tri-plane outputs need 16-byte aligned one-byte rows
        print("fc   %6dx%5dx%6d bf16x3 err %.2e | f16+f8 err %.2e (fp32 out only)" % (M, N, K, e0, e1), flush=True)
        return e1
    # tri-plane output + ReLU, read back
    ot = dense.tri_alloc((M, N), "cuda")
    oe = dense.exp_for(float(torch.relu(ref).max()))
    dense.igemm2(xt.view(1, 1, M, K), 1, 1, M, K, wt, N, 1, bias=b, relu=True, out=ot, bn=bn, out_exp=oe)
    torch.cuda.synchronize()
    e2 = err(ot.float(), torch.relu(ref))
    want = dense.tri_from_f32(torch.relu(out1), exp=oe)
    same = (torch.equal(ot.h, want.h), torch.equal(ot.l, want.l), torch.equal(ot.c, want.c))
    t0, t1 = timeit(f0), timeit(f1)
    fl = 2.0 * M * N * K / 1e9
    print("fc   %6dx%5dx%6d bn=%3d  bf16x3 err %.2e %7.3f ms %6.0f TF/s | f16+f8 err %.2e %7.3f ms %6.0f TF/s"
          " | tri-out err %.2e planes==torch %s amax %.3f (true %.3f)"
          % (M, N, K, bn, e0, t0, fl / t0, e1, t1, fl / t1, e2, same, am, float(ref.abs().max())), flush=True)
    return e1


def conv(B, H, W, cin, cout, pool=False):
    g = torch.Generator(device="cuda").manual_seed(H + W + cin)
    x = torch.relu(torch.randn(B, H, W, cin, device="cuda", generator=g))
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1)
    ref = torch.relu(ref)
    if pool:
        ref = torch.nn.functional.max_pool2d(ref, 2, 2, ceil_mode=True)
    ref = ref.permute(0, 2, 3, 1)
    xt, wt = dense.tri_from_f32(x), dense.conv_weight_to_tri(w)
    Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if pool else (H, W)
    ot = dense.tri_alloc((B, Ho, Wo, cout), "cuda")
    oe = dense.exp_for(float(ref.max()))
    f1 = lambda: dense.igemm2(xt, B, H, W, cin, wt, cout, 9, bias=b, relu=True, out=ot, pool=pool, out_exp=oe)
    f1()
    torch.cuda.synchronize()
    e1 = err(ot.float(), ref)
    xs, ws = dense.split(x), dense.conv_weight_to_split(w)
    o0 = torch.empty(2, B, Ho, Wo, cout, dtype=torch.bfloat16, device="cuda")
    halo_was = cout <= 128
    f0 = lambda: dense.igemm(xs, B, H, W, cin, ws, cout, 9, bias=b, relu=True, out=o0, pool=pool)
    f0()
    e0 = err(dense.merge(o0), ref)
    # split-bf16 input -> tri-plane output (what conv2_2 hands to conv3_1)
    ot2 = dense.tri_alloc((B, Ho, Wo, cout), "cuda")
    dense.igemm2(xs, B, H, W, cin, ws, cout, 9, bias=b, relu=True, out=ot2, pool=pool, out_exp=oe)
    torch.cuda.synchronize()
    e2 = err(ot2.float(), ref)
    t0, t1 = timeit(f0), timeit(f1)
    fl = 2.0 * B * H * W * cout * 9 * cin / 1e9
    print("conv %dx%dx%dx%d->%d pool=%d bf16x3 err %.2e %7.3f ms %6.0f TF/s | f16+f8 err %.2e %7.3f ms %6.0f TF/s"
          " | bf16-in tri-out err %.2e" % (B, H, W, cin, cout, pool, e0, t0, fl / t0, e1, t1, fl / t1, e2), flush=True)
    return e1


if __name__ == "__main__":
    worst = 0.0
    worst = max(worst, fc(256, 256, 512))
    worst = max(worst, fc(300, 441, 256))
    worst = max(worst, fc(2400, 4096, 4096, bn=192))
    worst = max(worst, fc(2400, 4096, 25088, bn=192))
    worst = max(worst, fc(2400, 4096, 25088, bn=256))
    worst = max(worst, fc(2400, 126, 8192))
    worst = max(worst, fc(19152, 54, 512))
    worst = max(worst, fc(600, 256, 100352))
    worst = max(worst, fc(2400, 4096, 4096, bn=192, amp=300.0))
    worst = max(worst, conv(1, 38, 63, 512, 512))
    worst = max(worst, conv(8, 38, 63, 512, 512))
    worst = max(worst, conv(8, 75, 125, 512, 512, pool=True))
    worst = max(worst, conv(8, 75, 125, 256, 512))
    worst = max(worst, conv(2, 150, 250, 256, 256, pool=True))
    worst = max(worst, conv(2, 300, 500, 128, 128, pool=True))
    worst = max(worst, conv(1, 300, 500, 64, 128))
    print("worst f16+f8 per-layer error %.2e (gate 2e-5)" % worst)
