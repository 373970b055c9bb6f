"""Per-CTA fixed cost of the attention kernels: time(T) / waves = n_tiles(T) * t_tile + X for T = 1000, 2000, 4000 at B = 37 (grids of
exactly 2 / 4 / 8 waves of 148 CTAs), least-squares fit of the per-tile time and the per-CTA overhead X (development aid)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-integer-programming_b200"))
import numpy as np
import torch
from dip_b200 import ops

D = 128


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3       # us


def main():
    B = 37
    dev = torch.device("cuda")
    rows = {"fwd": [], "bwd_all": []}
    for T in (1000, 2000, 4000):
        qkv = torch.randn(B * T, 3 * D, device=dev)
        d_o = torch.randn(B * T, D, device=dev)
        hi, lo = ops.split_bf16(qkv)
        o, lse = ops.attn_fwd_tc(hi, lo, B, T)
        waves = (-(-T // 128)) * B / 148.0
        rows["fwd"].append((T, waves, timeit(lambda: ops.attn_fwd_tc(hi, lo, B, T))))
        rows["bwd_all"].append((T, waves, timeit(lambda: ops.attn_bwd_tc(hi, lo, o, d_o, lse, B, T))))
    out = {}
    for k, r in rows.items():
        nt = np.array([T / 64.0 for T, _, _ in r])
        per_wave = np.array([t / w for _, w, t in r])
        A = np.stack([nt, np.ones_like(nt)], 1)
        (t_tile, X), *_ = np.linalg.lstsq(A, per_wave, rcond=None)
        out[k] = {"us": [round(t, 1) for _, _, t in r], "us_per_wave": [round(x, 1) for x in per_wave], "us_per_64_tile": round(float(t_tile), 3),
                  "fixed_us_per_cta": round(float(X), 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
