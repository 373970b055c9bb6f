"""Where the per-CTA fixed cost of the attention kernels goes: clock64 events of thread 0 of CTA (0,0) (a row thread) at kernel entry (1),
after the set-up barrier (2), first score tile seen (31), last tile done -> epilogue entered (3), epilogue done (4), after the final
barrier (5).  Needs the -DDIP_TRACE build (tools/build_trace.sh)."""
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


def run(name, call):
    for _ in range(2): call()
    buf = torch.zeros(5 * 16384, dtype=torch.int64, device=dev)
    fn(buf.data_ptr())
    call()
    torch.cuda.synchronize()
    fn(None)
    a = buf.cpu().numpy().astype(np.uint64)[:16384]          # region 0 = thread 0
    evs = [(int(v) & 0xffffffff, (int(v) >> 32) & 0xffff, int(v) >> 48) for v in a if int(v) != 0]
    t0 = [c for c, t, e in evs if e == 1][0]
    rel = lambda c: (c - t0) & 0xffffffff
    first = {}
    last = {}
    for c, t, e in evs:
        first.setdefault(e, rel(c)); last[e] = rel(c)
    n_tiles = len([1 for c, t, e in evs if e == 31])
    print(name, "T", T, "tiles", n_tiles, {"setup_done": first.get(2), "first_scores_seen": first.get(31), "first_P_stored/arrive": first.get(35), "loop_done": first.get(3),
                 "epilogue_done": first.get(4), "after_final_barrier": first.get(5)})
    steady = (last.get(31) - first.get(31)) / max(n_tiles - 1, 1)
    print("   cycles per tile (steady)", round(steady), " prologue", first.get(31), " epilogue+exit", first.get(5) - first.get(3),
          " fixed share of CTA", round((first.get(31) + first.get(5) - first.get(3)) / first.get(5), 3))


run("fwd", lambda: ops.attn_fwd_tc(hi, lo, B, T))
run("bwd dkdv", lambda: ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T))
