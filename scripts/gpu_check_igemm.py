"""Bring-up check of the tensor-core implicit GEMM against torch fp64 on the GPU (run via gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnc_b200 import dense

torch.manual_seed(0)
dev = "cuda"


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def check_gemm(M, K, N, bn=0, split_k=1, relu=False, impl="tc"):
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    xs, ws = dense.split(x), dense.split(w)
    ref = dense.merge(xs).double() @ dense.merge(ws).double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    if split_k > 1:
        part = torch.zeros(split_k, M, N, device=dev)
        dense.igemm(xs.view(2, 1, 1, M, K), 1, 1, M, K, ws, N, 1, out_f32=part, split_k=split_k,
                    split_stride=M * N, bn=bn)
        out = torch.empty(M, N, device=dev)
        dense.splitk_reduce(part, split_k, M * N, M, N, bias=b, relu=relu, out_f32=out)
    else:
        out = torch.full((M, N), float("nan"), device=dev)
        dense.igemm(xs.view(2, 1, 1, M, K), 1, 1, M, K, ws, N, 1, bias=b, relu=relu, out_f32=out,
                    bn=bn, impl=impl)
    torch.cuda.synchronize()
    e = relerr(out, ref)
    print("gemm M=%d K=%d N=%d bn=%d split=%d impl=%s relerr=%.3e nan=%d" % (
        M, K, N, bn, split_k, impl, e, int(torch.isnan(out).sum())), flush=True)
    return e


def check_conv(B, H, W, Cin, Cout, bn=0, impl="tc", split_out=True):
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=dev)
    xs = dense.split(x.permute(0, 2, 3, 1).contiguous())
    ws = dense.conv_weight_to_split(w)
    xr = dense.merge(xs).permute(0, 3, 1, 2).double()
    wr = dense.merge(ws).view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).double()
    ref = torch.nn.functional.conv2d(xr, wr, b.double(), padding=1).clamp_min(0)
    ref = ref.permute(0, 2, 3, 1)
    if split_out:
        out = torch.zeros(2, B, H, W, Cout, device=dev, dtype=torch.bfloat16)
        dense.igemm(xs, B, H, W, Cin, ws, Cout, 9, bias=b, relu=True, out=out, bn=bn, impl=impl)
        got = dense.merge(out)
    else:
        got = torch.full((B, H, W, Cout), float("nan"), device=dev)
        dense.igemm(xs, B, H, W, Cin, ws, Cout, 9, bias=b, relu=True, out_f32=got, bn=bn, impl=impl)
    torch.cuda.synchronize()
    e = relerr(got, ref)
    print("conv B=%d %dx%d Cin=%d Cout=%d bn=%d impl=%s split_out=%d relerr=%.3e" % (
        B, H, W, Cin, Cout, bn, impl, split_out, e), flush=True)
    return e


def bench_conv(B, H, W, Cin, Cout, bn=0, iters=10):
    xs = dense.split(torch.randn(B, H, W, Cin, device=dev))
    ws = dense.conv_weight_to_split(torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5)
    out = torch.zeros(2, B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        dense.igemm(xs, B, H, W, Cin, ws, Cout, 9, relu=True, out=out, bn=bn)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        dense.igemm(xs, B, H, W, Cin, ws, Cout, 9, relu=True, out=out, bn=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * H * W * Cout * 9 * Cin
    print("bench conv B=%d %dx%d %d->%d bn=%d: %.3f ms  %.1f TFLOP/s algorithmic (x3 tensor)" % (
        B, H, W, Cin, Cout, bn, ms, fl / ms / 1e9), flush=True)


def bench_gemm(M, K, N, bn=0, iters=10):
    xs = dense.split(torch.randn(M, K, device=dev))
    ws = dense.split(torch.randn(N, K, device=dev))
    out = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        dense.igemm(xs.view(2, 1, 1, M, K), 1, 1, M, K, ws, N, 1, relu=True, out=out, bn=bn)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        dense.igemm(xs.view(2, 1, 1, M, K), 1, 1, M, K, ws, N, 1, relu=True, out=out, bn=bn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("bench gemm M=%d K=%d N=%d bn=%d: %.3f ms  %.1f TFLOP/s algorithmic" % (
        M, K, N, bn, ms, 2.0 * M * K * N / ms / 1e9), flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    # SIMT cross-check first (plain CUDA, should just work)
    check_gemm(300, 256, 441, impl="simt")
    check_conv(1, 38, 63, 64, 64, impl="simt")
    # tensor-core path, smallest first
    check_gemm(128, 64, 64, bn=64)
    check_gemm(128, 256, 64, bn=64)
    check_gemm(300, 256, 441, bn=64)
    check_gemm(300, 512, 128, bn=128)
    check_gemm(2400, 4096, 4096, bn=256)
    check_gemm(300, 4096, 256, bn=256, split_k=4, relu=True)
    check_conv(1, 8, 16, 64, 64, bn=64)
    check_conv(1, 38, 63, 64, 64, bn=64)
    check_conv(2, 38, 63, 512, 512, bn=256)
    check_conv(1, 75, 125, 256, 512, bn=128)
    check_conv(1, 38, 63, 128, 54, bn=64, split_out=False)
    for bn in (128, 256):
        bench_conv(8, 38, 63, 512, 512, bn=bn)
        bench_conv(8, 75, 125, 512, 512, bn=bn)
    bench_conv(8, 150, 250, 256, 256, bn=256)
    bench_conv(8, 300, 500, 128, 128, bn=128)
    bench_conv(8, 600, 1000, 64, 64, bn=64)
    for bn in (128, 256):
        bench_gemm(2400, 25088, 4096, bn=bn)
    bench_gemm(2400, 4096, 4096, bn=256)
