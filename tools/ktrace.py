"""Event trace of CTA (0,0) of the tcgen05 attention forward (development aid; needs a -DDIP_TRACE build:
   nvcc ... -DDIP_TRACE, see tools/build_trace.sh)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import torch, numpy as np
from dip_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "libdip_trace.so")
from dip_b200 import ops
lib = _lib.load()
B, T, D = int(os.environ.get("KB_B", 16)), int(os.environ.get("KB_T", 4000)), 128
dev = torch.device("cuda")
qkv = torch.randn(B * T, 3 * D, device=dev)
hi, lo = ops.split_bf16(qkv)
for _ in range(2): ops.attn_fwd_tc(hi, lo, B, T)
buf = torch.zeros(5 * 16384, dtype=torch.int64, device=dev)
fn = lib.dip_debug_set_trace; fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
fn(buf.data_ptr())
ops.attn_fwd_tc(hi, lo, B, T)
torch.cuda.synchronize()
fn(None)
a = buf.cpu().numpy().astype(np.uint64)
names = {20: "mma:S begin", 21: "mma:k_full", 22: "mma:s_free", 23: "mma:S issued", 24: "mma:PV begin", 25: "mma:v_full", 26: "mma:p_full", 27: "mma:pv_free",
         28: "mma:PV issued", 30: "row:iter", 31: "row:s_full", 32: "row:S loaded", 33: "row:max xchg", 34: "row:exps+split", 35: "row:P stored", 36: "row:pv_full",
         37: "row:PV folded", 40: "prod:K begin", 41: "prod:K issue", 42: "prod:V begin", 43: "prod:V issue"}
evs = []
for v in a:
    v = int(v)
    if v == 0: continue
    evs.append((v & 0xffffffff, (v >> 32) & 0xffff, v >> 48))
t0 = min(e[0] for e in evs)
evs = sorted((((c - t0) & 0xffffffff), t, e) for c, t, e in evs)
for c, t, e in evs:
    if 30 <= t <= 32: print("%8d  tile %3d  %s" % (c, t, names.get(e, e)))
per = [c for c, t, e in evs if e == 35]
print("mean cycles per tile:", (per[-1] - per[5]) / (len(per) - 6))
