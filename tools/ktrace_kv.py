"""Event timeline of CTA (0,0) of the key-stationary attention backward (dK/dV) kernel, a few steady-state tiles (needs tools/build_trace.sh)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import torch, numpy as np
from dip_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "libdip_trace.so")
from dip_b200 import ops
lib = _lib.load()
B, T, D = 37, int(os.environ.get("KB_T", 2000)), 128
dev = torch.device("cuda")
qkv = torch.randn(B * T, 3 * D, device=dev)
d_o = torch.randn(B * T, D, device=dev)
hi, lo = ops.split_bf16(qkv)
o, lse = ops.attn_fwd_tc(hi, lo, B, T)
fn = lib.dip_debug_set_trace; fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
call = lambda: ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T)
for _ in range(2): call()
buf = torch.zeros(5 * 16384, dtype=torch.int64, device=dev)
fn(buf.data_ptr()); call(); torch.cuda.synchronize(); fn(None)
a = buf.cpu().numpy().astype(np.uint64)
names = {20: "mma: S^T wait Q", 22: "mma: S^T issued", 24: "mma: dP^T dO there", 25: "mma: dP^T issued", 28: "mma: dV acc issued (P there)", 23: "mma: dK acc issued (dS there)",
         30: "row: iter", 31: "row: S^T seen", 32: "row: P stored + dS flushed", 33: "row: dP^T seen", 34: "row: dS computed", 35: "row: dS stored",
         40: "prod: Q wait slot", 41: "prod: Q slot free", 42: "prod: Q issued / dO wait", 43: "prod: dO slot free"}
evs = []
for v in a:
    v = int(v)
    if v == 0: continue
    evs.append((v & 0xffffffff, (v >> 32) & 0xffff, v >> 48))
t0 = min(e[0] for e in evs)
evs = sorted((((c - t0) & 0xffffffff), t, e) for c, t, e in evs)
lo_t, hi_t = int(os.environ.get("TILE0", 12)), int(os.environ.get("TILE1", 14))
for c, t, e in evs:
    if lo_t <= t <= hi_t and e >= 20: print("%8d  tile %3d  %s" % (c, t, names.get(e, e)))
