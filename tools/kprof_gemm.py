import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import torch
from dip_b200 import ops
B, T, D = 16, 4000, 128
dev = torch.device("cuda")
x = torch.randn(B * T, D, device=dev); W = torch.randn(D, D, device=dev) * 0.1; bias = torch.randn(D, device=dev)
for _ in range(4):
    c = ops.gemm_tc(x, W, bias, res=x)["C"]
torch.cuda.synchronize()
print("ok")
