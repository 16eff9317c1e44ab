"""One launch set of the tcgen05 projection kernels at the 8 M x 128 x 128 shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepviewagg_b200 import ops
M, K, N = 8_000_000, 128, 128
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); g = torch.randn(M, N, device="cuda")
for _ in range(2):
    y = ops._tc_gemm(x, w, 0, N)
    gx = ops._tc_gemm(g, w, 1, K)
    gw = ops._tc_gemm(g, x, 2, K)
torch.cuda.synchronize()
print("done")
