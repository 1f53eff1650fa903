"""Halo-kernel launches (precision mode 1) at the conv1_2 / conv2_2 shapes of the benchmark, CTA-pair
and single-CTA forms, for an `ncu --set full --import-source on -k regex:conv_halo` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnc_b200 import dense


def conv(B, H, W, cin, cout, pool):
    x = torch.relu(torch.randn(B, H, W, cin, device="cuda"))
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
    xt, wt = dense.tri_from_f32(x), dense.conv_weight_to_tri(w)
    Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if pool else (H, W)
    ot = dense.tri_alloc((B, Ho, Wo, cout), "cuda")
    for pair in (1, 0):
        dense.set_halo_pair(pair)
        dense.igemm2(xt, B, H, W, cin, wt, cout, 9, relu=True, out=ot, out_exp=8, pool=pool)
    dense.set_halo_pair(1)


conv(8, 600, 1000, 64, 64, True)      # conv1_2
conv(8, 300, 500, 128, 128, True)     # conv2_2
torch.cuda.synchronize()
