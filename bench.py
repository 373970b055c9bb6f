#!/usr/bin/env python
"""bench.py -- guided-DDIM sampling throughput of the B200-native path (and of the CPU reference port).

A "step" is one pass of the hot path over one wave: B samples of a Set-Cover instance (BASELINE.json
configs[1] shapes: 1000 x 2000, nnz 100 000, D 128, denoiser 1 layer, decoder 2 layers) through
S = 100 guided DDIM steps (gs 1e5, lambda 0.9) + final decode + round + integer feasibility check.
The full configs[1] job (256 instances x 1000 samples) is 16 000 such waves at B = 16; the bench times
K of them per GPU (weak scaling: every rank owns its own instance and its own waves).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference --steps 3 --warmup 1      # CPU port of the reference, host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "diffusion-integer-programming_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

D = 128
METRIC = "guided-DDIM samples/sec (Set Cover, S=100; every sample decoded and integer-checked; feasible samples/sec reported beside)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--wave", type=int, default=37, help="samples per wave (bench step)")
    ap.add_argument("--ddim-steps", type=int, default=100)
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--cols", type=int, default=2000)
    ap.add_argument("--density", type=float, default=0.05)
    ap.add_argument("--gs", type=float, default=1e5)
    ap.add_argument("--lam", type=float, default=0.9)
    ap.add_argument("--cpu-samples", type=int, default=4, help="samples in the bounded CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=2, help="DDIM steps in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def step_flops(L, n_den=1, n_dec=2):
    """Algorithmic FLOPs per sample-step."""
    T1, T2 = L + 1, 2 * L
    den = 6 * T1 * D * D + n_den * (12 * T1 * D * D + 4 * T1 * T1 * D)
    dec_f = n_dec * (12 * T2 * D * D + 4 * T2 * T2 * D)
    dec_b = n_dec * (12 * T2 * D * D + 8 * T2 * T2 * D)
    return den, dec_f, dec_b


# --------------------------------------------------------------------------------------------------
# CPU reference port (oracle) -- the only place bench.py executes oracle/
# --------------------------------------------------------------------------------------------------
def cpu_port_rate(args, inst, mip, kpm, Wd, Wdec, n_rep=1):
    """samples/s of the reference's algorithm on the host cores: a bounded sample (cpu_samples samples x
    cpu_steps guided DDIM steps + one final decode), scaled to S steps (per-step cost is step-independent)."""
    from oracle import restate
    from dip_b200 import synth
    tw = lambda w: {k: torch.from_numpy(v) for k, v in w.items()}
    Wd_t, Wdec_t = tw(Wd), tw(Wdec)
    Bc, Sc, S = args.cpu_samples, args.cpu_steps, args.ddim_steps
    L = mip.shape[0]
    x = torch.from_numpy(synth.initial_noise(0, range(Bc), L))
    mipB, kpmB = torch.from_numpy(mip)[None].repeat(Bc, 1, 1), torch.from_numpy(kpm)[None].repeat(Bc, 1)
    A = (torch.from_numpy(inst.rows), torch.from_numpy(inst.cols), torch.from_numpy(inst.a_val), inst.M)
    b, c = torch.from_numpy(inst.b), torch.from_numpy(inst.c)
    tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], S)
    best = None
    for _ in range(n_rep):
        t0 = time.perf_counter()
        xx = x
        for k in range(Sc):
            # dec_requires_grad=True: like the reference, autograd also builds the (unused) decoder weight gradients
            xx = restate.guided_ddim_step(Wd_t, Wdec_t, xx, S - 1 - k, restate.ddim_step_coeffs(tab, k), mipB, kpmB, A, b, c, args.gs, args.lam,
                                          dec_requires_grad=True)["x_prev"]
        t1 = time.perf_counter()
        restate.final_decode(Wdec_t, mipB, xx, kpmB, A, b)
        t2 = time.perf_counter()
        total = (t1 - t0) / Sc * S + (t2 - t1)
        best = total if best is None else min(best, total)
    return Bc / best, best


# --------------------------------------------------------------------------------------------------
def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    def reader():
        for line in p.stdout:
            out.append(line.strip())
    th = threading.Thread(target=reader, daemon=True)
    th.start()
    stop.wait()
    p.terminate()
    th.join(timeout=2)


def summarize_clocks(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [s.strip() for s in ln.split(",")]
        if len(f) < 6:
            continue
        try:
            sm.append(float(f[0])); mx = max(mx, float(f[1]))
        except ValueError:
            continue
        for nm, v in zip(names, f[2:6]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.backends.cuda.matmul.allow_tf32 = False

    from dip_b200 import synth
    L = args.cols
    S = args.ddim_steps
    workload = ("Set Cover %dx%d nnz=%d (ecole SetCoverGenerator shapes), guided DDIM S=%d gs=%g lambda=%g, D=128, denoiser 1 layer, decoder 2 layers, "
                "random-init weights; step = one wave of %d samples (bounded slice of configs[1]: 256 instances x 1000 samples)"
                % (args.rows, args.cols, int(args.rows * args.cols * args.density), S, args.gs, args.lam, args.wave))
    Wd, Wdec = synth.denoiser_weights(1, seed=1), synth.decoder_weights(2, seed=2)

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return 0
        inst = synth.set_cover(args.rows, args.cols, args.density, seed=0)
        mip = synth.mip_features(L, inst.n, seed=10)
        kpm = np.zeros(L, dtype=bool)
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        for _ in range(args.warmup):
            cpu_port_rate(args, inst, mip, kpm, Wd, Wdec)
        rates, secs = [], []
        for _ in range(args.steps):
            r, s = cpu_port_rate(args, inst, mip, kpm, Wd, Wdec)
            rates.append(r); secs.append(s)
        value = args.cpu_samples * len(secs) / sum(secs)
        sample = "%d samples x %d guided DDIM steps (+1 final decode) per step, scaled x%d/%d to S=%d" % (
            args.cpu_samples, args.cpu_steps, S, args.cpu_steps, S)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU port of the reference algorithm (oracle/restate.py: same torch ops incl. SDPA + autograd with decoder weight grads); "
                    "/root/reference itself is Python needing absent torch_sparse/torch_geometric and does not exist on the GPU box",
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from dip_b200 import _lib, dist as ddist
    from dip_b200.engine import Engine, randn
    import ctypes as C
    lib = _lib.load()

    instance_id = rank                     # weak scaling: every rank owns its own instance and waves
    inst = synth.set_cover(args.rows, args.cols, args.density, seed=instance_id)
    kpm = np.zeros(L, dtype=bool)
    to_dev = lambda w: {k: torch.from_numpy(v).to(dev) for k, v in w.items()}
    eng = Engine()
    eng.set_denoiser(to_dev(Wd)); eng.set_decoder(to_dev(Wdec))

    # CMSP encoder (K3) on the synthetic bipartite graph -> conditions
    from dip_b200 import encoder as enc, ops
    Wenc = {k[len("mip_encoder.gnn_model."):]: v for k, v in to_dev(synth.encoder_weights(seed=3)).items()}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    graph = (T(inst.constraint_features), T(inst.edge_index), T(inst.edge_attr), T(inst.variable_features))
    int_idx = T(inst.int_indices)

    def encode_and_set():
        v = enc.gnn_forward(Wenc, *graph)
        mip = ops.gather_pad(v, int_idx, L)
        eng.set_instance(mip, T(kpm), inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c, a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)
        return mip

    mip_dev = encode_and_set()
    from dip_b200 import schedule
    steps = schedule.ddim_coeffs(S)

    B, K, W = args.wave, args.steps, args.warmup
    eng.workspace(B)

    def noise(wave_idx):
        return randn((B, L, D), seed=1000 + instance_id, subseq=wave_idx, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # warm-up
    for w in range(W):
        eng.sample_guided_ddim(noise(10_000 + w), steps, args.gs, args.lam)
    xs = [noise(k) for k in range(K)]
    rec = ddist.make_records(world, inst.n)

    stop, clk_lines = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, clk_lines, local_rank), daemon=True)
    th.start()
    time.sleep(0.3)

    # ---- timed region 1: inputs resident in HBM ----
    barrier()
    launches0 = lib.dip_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    encode_and_set()
    outs = []
    for k in range(K):
        outs.append(eng.sample_guided_ddim(xs[k], steps, args.gs, args.lam))
    for o in outs:
        ddist.update_record(rec, instance_id, o["xbits"].cpu().numpy(), o["viol_int"].cpu().numpy(), o["obj_int"].cpu().numpy())
    merged = ddist.all_gather_records(rec, dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.dip_launch_count() - launches0

    # ---- kernel durations: one more wave of the same shape with the in-library CUDA-event timer on (an event pair around every
    # launch, on the launch stream).  Kept out of the pass above because the event pairs cost ~4 % of throughput.
    barrier()
    lib.dip_prof_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    eng.sample_guided_ddim(noise(30_000), steps, args.gs, args.lam)
    p1.record()
    torch.cuda.synchronize()
    lib.dip_prof_enable(0)
    ms_prof_wave = p0.elapsed_time(p1)
    ncls = len(_lib.PROF_CLASSES)
    pms = (C.c_double * ncls)(); pcnt = (C.c_uint64 * ncls)()
    _lib.check(lib.dip_prof_collect(pms, pcnt, ncls))
    prof = {nm: {"ms": pms[i], "launches": int(pcnt[i])} for i, nm in enumerate(_lib.PROF_CLASSES)}

    # ---- timed region 2: end to end through the public API, host buffers, copies inside ----
    host_x = [torch.empty((B, L, D), dtype=torch.float32).pin_memory() for _ in range(K)]
    for k in range(K):
        host_x[k].copy_(noise(20_000 + k).cpu())
    def host_bufs():
        return {"xbits": torch.empty((B, inst.n), dtype=torch.uint8).pin_memory(), "penalty": torch.empty(B, dtype=torch.float32).pin_memory(),
                "viol_int": torch.empty(B, dtype=torch.int64).pin_memory(), "obj_int": torch.empty(B, dtype=torch.int64).pin_memory()}
    host_out = [host_bufs(), host_bufs()]                 # double-buffered: wave k is read back while wave k+1 is already queued
    done = [torch.cuda.Event(), torch.cuda.Event()]
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    n_feas_e2e = 0

    def collect(k):
        done[k & 1].synchronize()
        return int((host_out[k & 1]["viol_int"] == 0).sum())

    for k in range(K):
        xd = host_x[k].to(dev, non_blocking=True)                         # H2D of this wave's inputs (pinned)
        o = eng.sample_guided_ddim(xd, steps, args.gs, args.lam)
        for nm in host_out[k & 1]:
            host_out[k & 1][nm].copy_(o[nm], non_blocking=True)           # D2H of this wave's result
        done[k & 1].record()
        if k >= 1:
            n_feas_e2e += collect(k - 1)                                  # host reads wave k-1 while the GPU runs wave k
    n_feas_e2e += collect(K - 1)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    stop.set()
    th.join(timeout=3)

    t_ms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    feas_local = sum(int((o["viol_int"] == 0).sum()) for o in outs)
    cnt = torch.tensor([feas_local, n_feas_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])
    total_samples = world * K * B
    value = total_samples / (ms / 1e3)
    e2e_value = total_samples / (ms_e2e / 1e3)

    if rank == 0:
        pk = peaks()
        den, dec_f, dec_b = step_flops(L)
        T2 = 2 * L
        # dominant kernel: key-stationary attention backward (dK, dV, dP products): 6*Tq*Tk*D algorithmic FLOPs per sample-layer.
        # With the row windows each decoder layer's launch pairs 2L keys with the L live queries (last layer) or L keys with 2L
        # queries (layer 0): Tq*Tk = 2L*L for both -- the dead half the reference computes is not counted as achieved work.
        dk = prof["attn_bwd_dkdv"]
        flops_per_launch = 6.0 * T2 * L * D * B
        avg_ms = dk["ms"] / max(dk["launches"], 1)
        achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        upd, gd = prof["update"], prof["guidance"]
        upd_gbs = (16.0 * L * D * B) / (upd["ms"] / max(upd["launches"], 1) * 1e-3) / 1e9 if upd["ms"] > 0 else 0.0
        total_prof_ms = sum(v["ms"] for v in prof.values())
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "wave": B, "ddim_steps": S, "instance_per_rank": True,
                       "l2": "working set per wave (%.1f GB) exceeds the 126 MB L2; no flush needed" % (eng.workspace(B)[1] / 1e9),
                       "precision_mode": "attention and Linear layers on tcgen05 tensor cores with bf16x2-split fp32 operands (hi*hi + hi*lo + lo*hi, fp32 TMEM accumulation; one-step latent error ~1e-5 rel vs the fp32 reference); guidance, DDIM update, rounding and the integer feasibility check in fp32/int64 SIMT"},
            "feasible_per_s": float(cnt[0]) / (ms / 1e3), "feasible_ratio": float(cnt[0]) / total_samples,
            "tflops_effective": (den + dec_f + dec_b) * S * total_samples / (ms / 1e3) / 1e12,
            "tflops_effective_note": "reference-equivalent FLOPs (SURVEY 8d: what the reference executes per sample-step, dead tokens included) / time",
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": B * L * D * 4,
                    "d2h_bytes_per_step": B * inst.n + B * 4 + B * 16, "feasible_ratio": float(cnt[1]) / total_samples},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "attn_bwd_kv_tc_kernel (key-stationary decoder attention backward: dP, dV, dK; tcgen05)", "bound": "tensor", "achieved": achieved,
                         "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                         # DRAM bytes per launch from the committed ncu --set full capture (profiles/r01_ncu_dkdv_summary.txt: mean of the two
                         # decoder layers' launches at wave 37, SC shapes); algorithmic bytes of the launch are ~400 MB, i.e. no DRAM re-reads
                         "traffic": 314.7e6 if (B == 37 and L == 2000) else None,
                         "ceiling_note": "fp32-grade products cost 3 bf16 MMA passes and the launch recomputes S (8 executed per 6 credited products): "
                                         "attainable ceiling = peak * (1/3) * (6/8) = 0.25 of the bf16 peak",
                         "peak_source": pk["src"] + ", sustained bf16", "launch_ms": avg_ms,
                         "share_of_step": dk["ms"] / total_prof_ms if total_prof_ms > 0 else None},
            "kernel_time_share": {k: (v["ms"] / total_prof_ms if total_prof_ms > 0 else 0.0) for k, v in prof.items()},
            "kernel_timing": {"how": "one extra wave with a CUDA-event pair around every launch (dip_prof_*), after the timed region",
                              "wave_ms_with_events": ms_prof_wave, "sum_of_kernels_ms": total_prof_ms,
                              "gpu_busy_frac_of_timed_wave": total_prof_ms / (ms / K) if ms > 0 else None},
            "hbm_kernels": {"ddim_update_gbs": upd_gbs, "ddim_update_frac": upd_gbs / pk["hbm_gbs"], "peak_gbs": pk["hbm_gbs"]},
            "clocks": summarize_clocks(clk_lines),
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            mip_host = mip_dev.cpu().numpy()
            cpu_port_rate(args, inst, mip_host, kpm, Wd, Wdec)              # warm-up
            rate, secs = cpu_port_rate(args, inst, mip_host, kpm, Wd, Wdec, n_rep=2)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": "%d samples x %d guided DDIM steps + final decode, scaled to S=%d (%.1f s of CPU work per repetition)"
                                              % (args.cpu_samples, args.cpu_steps, S, secs * args.cpu_steps / S)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
