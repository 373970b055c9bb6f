"""Event trace of CTA (0,0) of the tcgen05 backward kernels (KB_KERNEL=ks: key-stationary dK/dV; default dq) (development aid; needs tools/libdip_trace.so)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import torch, numpy as np
from dip_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "libdip_trace.so")
from dip_b200 import ops
lib = _lib.load()
B, T, D = int(os.environ.get("KB_B", 37)), int(os.environ.get("KB_T", 4000)), 128
dev = torch.device("cuda")
qkv = torch.randn(B * T, 3 * D, device=dev); d_o = torch.randn(B * T, D, device=dev)
hi, lo = ops.split_bf16(qkv)
o, lse = ops.attn_fwd_tc(hi, lo, B, T)
which = os.environ.get("KB_KERNEL", "dq")
for _ in range(2): ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T, q_begin=2000, do_begin=2000)
buf = torch.zeros(9 * 16384, dtype=torch.int64, device=dev)
fn = lib.dip_debug_set_trace; fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
fn(buf.data_ptr())
ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T, q_begin=2000, do_begin=2000)
torch.cuda.synchronize(); fn(None)
a = buf.cpu().numpy().astype(np.uint64)
names_ks = {20: "mma:S begin", 22: "mma:S issued", 24: "mma:w_full", 25: "mma:dP issued", 23: "mma:dK issued", 28: "mma:dV issued", 30: "row:iter", 31: "row:g1_full", 32: "row:P stored", 33: "row:g2_full", 34: "row:dS ready+ds_free", 35: "row:dS stored", 40: "prod:U begin", 41: "prod:U issue", 42: "prod:W begin", 43: "prod:W issue"}
names = {20: "mma:begin", 21: "mma:v_full", 22: "mma:dP issued", 24: "mma:S begin", 25: "mma:S issued", 26: "mma:acc begin", 27: "mma:ds_full", 28: "mma:acc issued",
         30: "row:iter", 31: "row:dP full", 32: "row:dP loaded", 33: "row:S full", 34: "row:dS ready+ds_free", 35: "row:dS stored",
         40: "prod:U begin", 41: "prod:U issue", 42: "prod:W begin", 43: "prod:W issue"}
# the dK/dV kernel logs into regions 0-3, the dQ kernel into regions 5-8
lo_r = 0 if which == "ks" else 5
a = a[lo_r * 16384:(lo_r + 4) * 16384]
evs = []
for v in a:
    v = int(v)
    if v: evs.append((v & 0xffffffff, (v >> 32) & 0xffff, v >> 48))
t0 = min(e[0] for e in evs)
evs = sorted((((c - t0) & 0xffffffff), t, e) for c, t, e in evs)
if which == 'ks': names = names_ks
lo_t = int(os.environ.get("KB_TILE", 30))
for c, t, e in evs:
    if lo_t <= t <= lo_t + 1 and e in names: print("%8d  tile %3d  %s" % (c, t, names[e]))
per = [c for c, t, e in evs if e == 35]
print("mean cycles per tile:", (per[-1] - per[5]) / (len(per) - 6))
