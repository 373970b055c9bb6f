"""Micro-benchmarks of single kernels with CUDA events (not part of the judged bench; development aid)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import torch
from dip_b200 import ops

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    B, T, D = int(os.environ.get("KB_B", 16)), int(os.environ.get("KB_T", 4000)), 128
    dev = torch.device("cuda")
    qkv = torch.randn(B * T, 3 * D, device=dev)
    d_o = torch.randn(B * T, D, device=dev)
    hi, lo = ops.split_bf16(qkv)
    o, lse = ops.attn_fwd(qkv, B, T)
    res = {"B": B, "T": T}
    fl = 4.0 * T * T * D * B
    t = timeit(lambda: ops.attn_fwd(qkv, B, T)); res["fwd_simt_ms"] = t; res["fwd_simt_tflops"] = fl / t / 1e9
    t = timeit(lambda: ops.attn_fwd_tc(hi, lo, B, T)); res["fwd_tc_ms"] = t; res["fwd_tc_tflops"] = fl / t / 1e9
    t = timeit(lambda: ops.attn_bwd(qkv, o, d_o, lse, B, T)); res["bwd_simt_ms"] = t; res["bwd_simt_tflops"] = 2 * fl / t / 1e9
    if hasattr(ops, "attn_bwd_tc"):
        t = timeit(lambda: ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T)); res["bwd_tc_ms"] = t; res["bwd_tc_tflops"] = 2 * fl / t / 1e9
    x = torch.randn(B * T, D, device=dev); W = torch.randn(D, D, device=dev) * 0.1; bias = torch.randn(D, device=dev)
    t = timeit(lambda: ops.gemm_nt(x, W, bias)); res["gemm128_simt_ms"] = t; res["gemm128_simt_tflops"] = 2.0 * B * T * D * D / t / 1e9
    res["gemm128_simt_gbs"] = 2.0 * B * T * D * 4 / t / 1e6
    import ctypes as C
    from dip_b200 import _lib
    from dip_b200._lib import GemmArgs, Rows, ptr, check
    from dip_b200.engine import _stream
    lib = _lib.load()
    def tc_gemm(A, Wm, out, K, N, res=None, ln=None, split=None):
        whi, wlo = ops.split_bf16(Wm)
        def run():
            g = GemmArgs(); g.A, g.lda, g.bias, g.C, g.ldc = ptr(A), K, ptr(bias[:N] if N <= D else None), ptr(out), N
            g.M, g.N, g.K, g.act = A.shape[0], N, K, 0
            g.res = Rows(None, ptr(res), 0, 0)
            if split is not None: g.split_hi, g.split_lo = ptr(split[0]), ptr(split[1])
            check(lib.dip_gemm_tc(C.byref(g), ptr(whi), ptr(wlo), None, ptr(ln[0]) if ln else None, ptr(ln[1]) if ln else None, None, None, _stream()))
        return run
    out = torch.empty(B * T, D, device=dev)
    t = timeit(tc_gemm(x, W, out, D, D)); res["gemm128_tc_ms"] = t; res["gemm128_tc_gbs"] = 2.0 * B * T * D * 4 / t / 1e6
    t = timeit(tc_gemm(x, W, out, D, D, res=x)); res["gemm128_tc_res_ms"] = t; res["gemm128_tc_res_gbs"] = 3.0 * B * T * D * 4 / t / 1e6
    W3 = torch.randn(3 * D, D, device=dev) * 0.1
    hi3 = torch.empty(B * T, 3 * D, dtype=torch.bfloat16, device=dev); lo3 = torch.empty_like(hi3)
    gam = torch.ones(D, device=dev); bet = torch.zeros(D, device=dev)
    t = timeit(tc_gemm(x, W3, None, D, 3 * D, ln=(gam, bet), split=(hi3, lo3))); res["qkv_tc_ln_planes_ms"] = t; res["qkv_tc_gbs"] = 4.0 * B * T * D * 4 / t / 1e6
    x3 = torch.randn(B * T, 3 * D, device=dev); W4 = torch.randn(D, 3 * D, device=dev) * 0.1
    t = timeit(tc_gemm(x3, W4, out, 3 * D, D)); res["gemm_k384_tc_ms"] = t; res["gemm_k384_tc_gbs"] = 4.0 * B * T * D * 4 / t / 1e6
    t = timeit(lambda: ops.gemm_nt(x3, W4)); res["gemm_k384_simt_ms"] = t
    print(json.dumps(res))

if __name__ == "__main__":
    main()
