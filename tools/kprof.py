"""Tiny driver for ncu: runs each tensor-core attention kernel a few times at B=4, T=4000."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import torch
from dip_b200 import ops
B, T, D = int(os.environ.get("KB_B", 4)), int(os.environ.get("KB_T", 4000)), 128
dev = torch.device("cuda")
torch.manual_seed(0)
qkv = torch.randn(B * T, 3 * D, device=dev); d_o = torch.randn(B * T, D, device=dev)
mask = torch.zeros(T, dtype=torch.bool, device=dev)
hi, lo = ops.split_bf16(qkv)
for _ in range(3):
    o, lse = ops.attn_fwd_tc(hi, lo, B, T, mask)
    dqkv = ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T, mask)
torch.cuda.synchronize()
print("ok", float(o.abs().mean()), float(dqkv.abs().mean()))
# GEMMs: fp32 SIMT vs tcgen05 on the decoder-sized activation
x = torch.randn(B * T, D, device=dev); W = torch.randn(D, D, device=dev) * 0.1; bias = torch.randn(D, device=dev)
for _ in range(3):
    a = ops.gemm_nt(x, W, bias)
    c = ops.gemm_tc(x, W, bias)["C"]
torch.cuda.synchronize()
print("gemm ok", float((a - c).abs().max()))
