import torch, sys, os
sys.path.insert(0, "/root/repo")
import bench
out = torch.empty(2000, 512, 28, 28, device="cuda")
src = torch.empty_like(out)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = bench.median_ms(lambda: out.fill_(1.0), flush=flush)
print("fill 3.2GB: %.3f ms -> %.0f GB/s write-only" % (ms, out.numel()*4/ms/1e6))
ms = bench.median_ms(lambda: out.copy_(src), flush=flush)
print("copy 3.2GB: %.3f ms -> %.0f GB/s r+w" % (ms, 2*out.numel()*4/ms/1e6))
ms = bench.median_ms(lambda: out.zero_(), flush=flush)
print("zero (memset) 3.2GB: %.3f ms -> %.0f GB/s" % (ms, out.numel()*4/ms/1e6))
