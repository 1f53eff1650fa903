"""Error model for DESIGN.md section 11 item 1 (CPU, no GPU needed): how accurate is

    D = A_hi.B_hi (bf16 x bf16, fp32 accumulate)  +  corrections A_lo.B_hi + A_hi.B_lo

when the two correction products are computed from block-scaled FP8 (E4M3 values, one power-of-two
scale per 32 K-elements, as `kind::mxf8f6f4` does) instead of bf16?  Compared with the current
3 x bf16 scheme and with dropping the corrections, against an fp64 reference, on post-ReLU
activations x He-initialised weights of the layer shapes of the path."""
import numpy as np
import torch

torch.manual_seed(0)


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def fp16(x):
    return x.to(torch.float16).to(torch.float32)


def mx_e4m3(x, block=32):
    """Per-row blocks of `block` K-elements share a power-of-two scale; values stored as E4M3."""
    M, K = x.shape
    xb = x.reshape(M, K // block, block)
    amax = xb.abs().amax(dim=2, keepdim=True)
    # E8M0 scale: the block maximum lands in [128, 256), inside E4M3's range (max normal 448)
    scale = torch.where(amax > 0, torch.exp2(torch.floor(torch.log2(amax.clamp_min(1e-30))) - 7),
                        torch.ones_like(amax))
    q = (xb / scale).to(torch.float8_e4m3fn).to(torch.float32) * scale
    return q.reshape(M, K)


def run(M, N, K):
    a = torch.relu(torch.randn(M, K))                       # post-ReLU activations
    w = torch.randn(N, K) * (2.0 / K) ** 0.5                # He init
    ref = a.double() @ w.double().T
    a_hi, w_hi = bf16(a), bf16(w)
    a_lo, w_lo = bf16(a - a_hi), bf16(w - w_hi)
    main = a_hi @ w_hi.T
    corr_bf16 = a_lo @ w_hi.T + a_hi @ w_lo.T
    corr_fp8 = mx_e4m3(a_lo) @ mx_e4m3(w_hi).T + mx_e4m3(a_hi) @ mx_e4m3(w_lo).T
    # variant: fp16 main product (11-bit significands: residuals are 2^-12, eight times smaller)
    a_h16, w_h16 = fp16(a), fp16(w)
    a_l16, w_l16 = a - a_h16, w - w_h16
    main16 = a_h16 @ w_h16.T
    corr16_fp8 = mx_e4m3(a_l16) @ mx_e4m3(w_h16).T + mx_e4m3(a_h16) @ mx_e4m3(w_l16).T
    den = ref.abs().max()

    def err(x):
        return ((x.double() - ref).abs().max() / den).item()
    return err(main), err(main + corr_bf16), err(main + corr_fp8), err(main16), err(main16 + corr16_fp8)


print("%-28s %11s %11s %13s %11s %13s" % ("layer (M x N x K)", "bf16 only", "3 x bf16", "bf16+2xfp8",
                                          "fp16 only", "fp16+2xfp8"))
for name, (M, N, K) in (("conv1_2 tile", (2048, 64, 576)), ("conv3_2 tile", (2048, 256, 2304)),
                        ("conv5_x tile", (1024, 512, 4608)), ("fc7", (600, 1024, 4096)),
                        ("fc6", (300, 512, 25088)), ("fc6_maskest", (300, 256, 100352))):
    e = run(M, N, K)
    print("%-28s %11.2e %11.2e %13.2e %11.2e %13.2e" % (("%s %dx%dx%d" % (name, M, N, K),) + e))
print("(max |error| / max |reference|; the path's bar is 1e-3 end to end over ~20 layers)")
