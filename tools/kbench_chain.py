"""Micro-benchmark of the row-chain kernels (chain_tc.cu) against the separate Linear launches they replace (development aid).
   python tools/kbench_chain.py [B T win]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import ctypes as C
import torch
from dip_b200 import ops, _lib
from dip_b200._lib import Rows, ptr, check
from dip_b200.engine import _stream

D = 128


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, T, win = (int(a) for a in (sys.argv[1:4] + ["37", "4000", "0"][len(sys.argv) - 1:]))
    dev = torch.device("cuda")
    lib = _lib.load()
    g = lambda *s: torch.randn(*s, device=dev)
    w = {"attn.out_proj.weight": g(D, D) * 0.1, "attn.out_proj.bias": g(D), "ln_2.weight": 1 + 0.1 * g(D), "ln_2.bias": 0.1 * g(D),
         "mlp.c_fc.weight": g(D, D) * 0.1, "mlp.c_fc.bias": g(D), "mlp.c_proj.weight": g(D, D) * 0.1, "mlp.c_proj.bias": g(D)}
    pl = {k: ops.split_bf16(w[k]) for k in ("attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight")}
    plT = {k: ops.split_bf16(w[k].t().contiguous()) for k in ("attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight")}
    R = B * T
    attn, xin, dxout = g(R, D), g(R, D), g(R, D)
    xout, xmid, u = torch.empty(R, D, device=dev), torch.empty(R, D, device=dev), torch.empty(R, D, device=dev)
    mean2, rstd2 = torch.empty(R, device=dev), torch.empty(R, device=dev)
    dxm, dao = torch.empty(R, D, device=dev), torch.empty(R, D, device=dev)
    hi, lo = torch.empty(R, D, dtype=torch.bfloat16, device=dev), torch.empty(R, D, dtype=torch.bfloat16, device=dev)
    o, f, p = "attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight"

    def fwd():
        check(lib.dip_block_chain_fwd_tc(ptr(xout), ptr(xmid), ptr(mean2), ptr(rstd2), ptr(u), ptr(attn), Rows(None, ptr(xin), T, 0), ptr(pl[o][0]), ptr(pl[o][1]),
                                         ptr(w["attn.out_proj.bias"]), ptr(w["ln_2.weight"]), ptr(w["ln_2.bias"]), ptr(pl[f][0]), ptr(pl[f][1]), ptr(w["mlp.c_fc.bias"]),
                                         ptr(pl[p][0]), ptr(pl[p][1]), ptr(w["mlp.c_proj.bias"]), B, T, win, _stream()))

    def bwd():
        check(lib.dip_block_chain_bwd_tc(ptr(dxm), ptr(dao), ptr(hi), ptr(lo), ptr(dxout), ptr(u), ptr(xmid), ptr(mean2), ptr(rstd2), ptr(w["ln_2.weight"]),
                                         ptr(plT[p][0]), ptr(plT[p][1]), ptr(plT[f][0]), ptr(plT[f][1]), ptr(plT[o][0]), ptr(plT[o][1]), B, T, win, _stream()))

    live = B * (T - win)
    tiles = (live + 127) // 128
    res = {"B": B, "T": T, "win": win, "tiles": tiles}
    for name, fn, passes in (("fwd", fwd, 5), ("bwd", bwd, 6)):     # fp32 row passes that touch HBM (bwd: dxout, u, xmid in; dxm, dao out + 2 bf16 planes)
        t = timeit(fn)
        res[name + "_ms"] = t
        res[name + "_GBs"] = passes * live * D * 4 / t / 1e6
        res[name + "_us_per_tile_per_sm"] = t * 1e3 / ((tiles + 147) // 148)
    # the separate launches (one plain Linear with residual as a yardstick)
    x = g(R, D); out = torch.empty(R, D, device=dev)
    t = timeit(lambda: ops.gemm_tc(x, w[o], w["attn.out_proj.bias"], out=out, res=xin, res_T=T))
    res["one_linear_res_ms"] = t
    print(json.dumps(res))


if __name__ == "__main__":
    main()
