"""Event trace + timing of the tcgen05 Linear kernel (development aid; needs tools/libdip_trace.so from tools/build_trace.sh)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import torch, numpy as np
from dip_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "libdip_trace.so")
from dip_b200 import ops
lib = _lib.load()
B, T, D = int(os.environ.get("KB_B", 37)), 4000, 128
dev = torch.device("cuda")
x = torch.randn(B * T, D, device=dev); W = torch.randn(D, D, device=dev) * 0.1; bias = torch.randn(D, device=dev)
W3 = torch.randn(3 * D, D, device=dev) * 0.1; b3 = torch.randn(3 * D, device=dev)
g, be = torch.ones(D, device=dev), torch.zeros(D, device=dev)
out = torch.empty(B * T, D, device=dev); pre = torch.empty(B * T, D, device=dev)
x3 = torch.randn(B * T, 3 * D, device=dev); W384 = torch.randn(D, 3 * D, device=dev) * 0.1
cases = {
    "plain": lambda: ops.gemm_tc(x, W, bias, out=out),
    "residual": lambda: ops.gemm_tc(x, W, bias, out=out, res=x),
    "aux": lambda: ops.gemm_tc(x, W, None, out=out, act="mul_quickgelu_grad", aux=x),
    "ln_qkv_planes": lambda: ops.gemm_tc(x, W3, b3, ln=(g, be), want_stats=True, want_split=True, want_c=False),
    "planes_out": lambda: ops.gemm_tc(x, W, None, out=out, want_split=True),
    "ln_fc_gelu_preact": lambda: ops.gemm_tc(x, W, bias, out=out, ln=(g, be), want_stats=True, act="quickgelu", preact=pre),
    "k384": lambda: ops.gemm_tc(x3, W384, None, out=out),
}
names = {1: "ld:begin", 2: "ld:stage free", 3: "ld:stored", 10: "mma:begin", 11: "mma:a_full", 12: "mma:issued", 20: "ep:begin", 21: "ep:c_full", 22: "ep:chunk", 23: "ep:tmem loaded", 24: "ep:staged", 25: "ep:batch stored"}
fn = lib.dip_debug_set_trace_gemm; fn.restype = C.c_int; fn.argtypes = [C.c_void_p]
kn = lib.dip_debug_set_gemm_knobs; kn.restype = C.c_int; kn.argtypes = [C.c_int]
knobs = int(os.environ.get("KNOBS", 0)); kn(knobs); print("knobs", knobs)
for name, f in cases.items():
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print("%-16s %8.1f us per launch (includes split of W on the host side wrapper)" % (name, e0.elapsed_time(e1) * 100))
    buf = torch.zeros(3 * 16384, dtype=torch.int64, device=dev)
    fn(buf.data_ptr()); f(); torch.cuda.synchronize(); fn(None)
    a = buf.cpu().numpy().astype(np.uint64)
    evs = []
    for v in a:
        v = int(v)
        if v: evs.append((v & 0xffffffff, (v >> 32) & 0xffff, v >> 48))
    t0 = min(e[0] for e in evs)
    evs = sorted((((c - t0) & 0xffffffff), t, e) for c, t, e in evs)
    tiles = sorted(set(t for _, t, _ in evs))
    show = [3, 4, 3 * 148, 4 * 148]
    for c, t, e in evs:
        if t in show: print("   %8d  tile %5d  %s" % (c, t, names.get(e, e)))
    st = [c for c, t, e in evs if e == 3]
    print("   cycles per tile (loader):", (st[-1] - st[1]) / max(len(st) - 2, 1), "tiles", len(tiles))
