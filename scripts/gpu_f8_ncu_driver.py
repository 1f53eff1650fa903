"""One launch per precision mode of two representative shapes, for an ncu --set full capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnc_b200 import dense

def fc(M, N, K, bn):
    x = torch.relu(torch.randn(M, K, device="cuda")); w = torch.randn(N, K, device="cuda") * (2.0 / K) ** 0.5
    out = torch.empty(M, N, device="cuda")
    xs, ws = dense.split(x), dense.split(w)
    xt, wt = dense.tri_from_f32(x), dense.tri_from_f32(w, weight=True)
    ot = dense.tri_alloc((M, N), "cuda"); os_ = torch.empty(2, M, N, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        dense.igemm2(xs.view(2, 1, 1, M, K), 1, 1, M, K, ws, N, 1, relu=True, out=os_, bn=bn)
        dense.igemm2(xt.view(1, 1, M, K), 1, 1, M, K, wt, N, 1, relu=True, out=ot, bn=bn, out_exp=8)
def conv(B, H, W, cin, cout):
    x = torch.relu(torch.randn(B, H, W, cin, device="cuda")); w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
    xs, ws = dense.split(x), dense.conv_weight_to_split(w)
    xt, wt = dense.tri_from_f32(x), dense.conv_weight_to_tri(w)
    ot = dense.tri_alloc((B, H, W, cout), "cuda"); os_ = torch.empty(2, B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        dense.igemm2(xs, B, H, W, cin, ws, cout, 9, relu=True, out=os_)
        dense.igemm2(xt, B, H, W, cin, wt, cout, 9, relu=True, out=ot, out_exp=8)
fc(2400, 4096, 25088, 192)
conv(8, 75, 125, 512, 512)
torch.cuda.synchronize()
