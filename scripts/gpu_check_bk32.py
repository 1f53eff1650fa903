"""BLOCK_K = 32 (SWIZZLE_64B) and BN = 192 variants of the implicit GEMM: correctness vs fp64, then
per-shape timing against the BLOCK_K = 64 configurations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnc_b200 import dense
from scripts.gpu_check_igemm import check_gemm, check_conv, bench_conv, bench_gemm

if __name__ == "__main__":
    dense.set_block_k(0)
    check_gemm(300, 512, 441, bn=192)
    check_gemm(2400, 4096, 4096, bn=256)
    check_gemm(2400, 4096, 4096, bn=192)
    check_gemm(300, 4096, 256, bn=256, split_k=4, relu=True)
    check_conv(2, 38, 63, 512, 512, bn=256)
    check_conv(1, 75, 125, 256, 512, bn=192)
    check_conv(1, 19, 33, 64, 200, bn=192)
    for bk, bns in ((64, (128, 256)), (0, (192, 256))):
        dense.set_block_k(bk)
        print("---- block_k setting", bk, flush=True)
        for bn in bns:
            bench_gemm(2400, 25088, 4096, bn=bn)
            bench_gemm(2400, 4096, 4096, bn=bn)
            bench_conv(8, 150, 250, 256, 256, bn=bn)
            bench_conv(8, 75, 125, 512, 512, bn=bn)
            bench_conv(8, 38, 63, 512, 512, bn=bn)
