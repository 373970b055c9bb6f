#!/usr/bin/env python
"""bench.py -- guided-DDIM sampling throughput of the B200-native path, next to the reference's own CPU path.

A "step" is one pass of the hot path over one WAVE of samples: S = 100 guided DDIM steps + final decode + round + integer
feasibility / objective check (sample.py:151-174).  Workloads (`--config`, BASELINE.json `configs`):

  sc          configs[1] slice (default): Set Cover 1000 x 2000, nnz 1e5 (ecole SetCoverGenerator shapes), gs 1e5, lambda 0.9,
              D 128, denoiser 1 layer, decoder 2 layers, random-init weights; wave = 148 samples (one per SM; 37 for cf)
  sc1         configs[0]: ONE instance x 100 samples (strong scaling: the 100 samples are split over the ranks)
  is          configs[2]: Independent Set 1500 nodes / 6360 rows with the two checkpoints the reference ships (2 denoiser layers)
  ca, cf      configs[3]: Combinatorial-Auction shapes (1500 bids / 775 rows) and Capacitated Facility Location (5050 / 5151,
              decoder sequence 10 100), instance encoder + sampler end to end
  partialsol  configs[4]: sample_partialsol.py flow -- 15 samples per instance, many instances, waves that MIX instances,
              device-side selection of the coverage subset, bit-packed partial solutions handed to the host

  python bench.py [--config sc] --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference --steps 1 --warmup 1     # the unmodified reference on the host cores (bounded sample)

Numbers in the JSON line: `value` = samples/s with inputs resident in HBM (one library call per wave); `e2e` = the same through the
reference-facing API (model.diffusion.DDIMSampler.constraint_guided_sample + model.decoder.MipConditionalDecorder, host tensors,
H2D / D2H inside the timed region) with the engine's one-call number beside it; `roofline` = the dominant kernel, `roofline_all`
every kernel class, each with its algorithmic work per launch, live CUDA-event duration and DRAM traffic per launch read from the
committed ncu capture (profiles/r02_ncu_kernels.json); `cpu_baseline` = the reference's CPU path on this box's host cores.
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
METRIC = "guided-DDIM samples/sec (S=100; every sample decoded and integer-checked; feasible samples/sec reported beside)"
GOLDEN = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    # wave 148 = one sample per SM: measured 80.5 samples/s against 78.6 at wave 37 on the same box (fewer launch / CTA boundaries per
    # sample; the run is power-capped either way, at a lower clock with the larger wave)
    "sc": dict(family="SC", gs=1e5, lam=0.9, n_den=1, ckpt=False, wave=148, spi=None,
               what="configs[1] slice (256 instances x 1000 samples is %(waves)d such waves): Set Cover %(M)dx%(n)d nnz=%(nnz)d"),
    "sc1": dict(family="SC", gs=1e5, lam=0.9, n_den=1, ckpt=False, wave=148, spi=100,
                what="configs[0]: ONE Set Cover instance %(M)dx%(n)d nnz=%(nnz)d x 100 samples, samples split over the ranks (strong scaling)"),
    "is": dict(family="IS", gs=2e4, lam=0.5, n_den=2, ckpt=True, wave=148, spi=None,
               what="configs[2]: Independent Set n=%(n)d M=%(M)d nnz=%(nnz)d, the reference's shipped checkpoints (2 denoiser layers, trained)"),
    "ca": dict(family="CA", gs=2e4, lam=0.7, n_den=1, ckpt=False, wave=148, spi=None,
               what="configs[3]: Combinatorial-Auction shapes n=%(n)d M=%(M)d nnz=%(nnz)d, encoder + sampler"),
    "cf": dict(family="CF", gs=1e3, lam=0.7, n_den=1, ckpt=False, wave=37, spi=None,
               what="configs[3]: Capacitated Facility Location n=%(n)d M=%(M)d nnz=%(nnz)d (decoder sequence 10100), encoder + sampler"),
    "partialsol": dict(family="SC", gs=1e5, lam=0.9, n_den=1, ckpt=False, wave=148, spi=15,
                       what="configs[4]: sample_partialsol.py flow, Set Cover %(M)dx%(n)d nnz=%(nnz)d, 15 samples per instance, waves mixing "
                            "instances, coverage 0.2 partial solutions selected and bit-packed on the device, handed to the host"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="sc", choices=sorted(CONFIGS))
    ap.add_argument("--wave", type=int, default=0, help="samples per wave (0: the config's default: 148 = one per SM; 37 = 148 SMs / 4 for cf, whose decoder sequence is 10 100)")
    ap.add_argument("--ddim-steps", type=int, default=100)
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--cols", type=int, default=2000)
    ap.add_argument("--density", type=float, default=0.05)
    ap.add_argument("--cpu-samples", type=int, default=8, help="samples in the bounded CPU sample (BASELINE.md section 3: B = 8)")
    ap.add_argument("--cpu-steps", type=int, default=5, help="DDIM steps in the bounded CPU sample (BASELINE.md section 3: S' = 5)")
    ap.add_argument("--cpu-full", action="store_true", help="reference arm: time ALL S steps of the bounded batch (no scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api-e2e", action="store_true", help="skip the DDIMSampler-API end-to-end region")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """DRAM bytes per launch of every profiled kernel, from the committed ncu capture (tools/ncu_summarize.py)."""
    for name in ("r02_ncu_kernels.json",):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(path):
            return json.load(open(path)), "profiles/" + name
    return {}, None


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def make_instance(args, family, seed):
    from dip_b200 import synth
    if family == "SC":
        return synth.set_cover(args.rows, args.cols, args.density, seed=seed)
    if family == "IS":
        return synth.independent_set(1500, 4860, 1500, seed=seed)
    if family == "CA":
        return synth.combinatorial_auction(n_items=775, n_bids=1500, seed=seed)
    if family == "CF":
        return synth.facility_location(100, 50, seed=seed)
    raise ValueError(family)


def load_weights(cfg):
    """(denoiser state_dict, decoder state_dict, encoder state_dict) as numpy / torch CPU tensors."""
    from dip_b200 import synth
    if cfg["ckpt"]:
        t = torch.load(os.path.join(GOLDEN, "ddpm_independent_set.pth"), map_location="cpu", weights_only=True)
        d = torch.load(os.path.join(GOLDEN, "decoder_independent_set.pth"), map_location="cpu", weights_only=True)
        Wd = {k: v.numpy() for k, v in t.items() if k.startswith("predict_model.")}
        Wdec = {k: v.numpy() for k, v in d.items()}
    else:
        Wd, Wdec = synth.denoiser_weights(cfg["n_den"], seed=1), synth.decoder_weights(2, seed=2)
    return Wd, Wdec, synth.encoder_weights(seed=3)


def step_flops(L, n_den=1, n_dec=2):
    """Algorithmic FLOPs per sample-step: what the reference executes, dead tokens included."""
    T1, T2 = L + 1, 2 * L
    den = 6 * T1 * D * D + n_den * (12 * T1 * D * D + 4 * T1 * T1 * D)
    dec_f = n_dec * (12 * T2 * D * D + 4 * T2 * T2 * D)
    dec_b = n_dec * (12 * T2 * D * D + 8 * T2 * T2 * D)
    return den, dec_f, dec_b


def class_work(L, n, B, n_den, n_dec, S, inst):
    """Algorithmic work of one wave (S guided steps + final decode) per kernel class, counted on the LIVE rows (the row windows skip
    the dead tokens the reference computes and throws away): {class: (bound, amount, unit)}.  Attention: 4 Tq Tk D forward,
    6 Tq Tk D for the dK/dV/dP launch, 2 Tq Tk D for dQ (recomputation not credited).  Row-wise kernels: bytes a fused
    implementation must move per row (input row + output row, + saved activations where the backward needs them)."""
    T1, T2 = L + 1, 2 * L
    row = 4.0 * D
    fwd = n_den * 4.0 * T1 * T1 * D + 4.0 * T2 * T2 * D * max(n_dec - 1, 0) + 4.0 * L * T2 * D      # last layer: only the kept L queries
    dec_fwd_final = 4.0 * T2 * T2 * D * max(n_dec - 1, 0) + 4.0 * L * T2 * D
    pair = T2 * L * n_dec * 1.0                                                                     # live (query, key) pairs per sample, both layers
    # Linear + LayerNorm chains, bytes per sample-step: denoiser (mlp_xt 2 GEMMs, per block in-proj / out-proj / fc / proj), decoder fwd and
    # input-gradient bwd (the same four Linears + two LayerNorms per block, mirrored); per row and Linear: read 512 B + write 512 B
    den_rows = T1 * (2 + 4 * n_den)
    dec_rows_f = (T2 + L) * 4.0 if n_dec == 2 else T2 * 4.0 * n_dec                                   # layer 0: 2L rows, last layer: L live rows
    dec_rows_b = dec_rows_f
    gemm_bytes = (den_rows + dec_rows_f + dec_rows_b) * 2 * row
    return {
        "attn_fwd": ("tensor", B * (S * fwd + dec_fwd_final), "FLOP"),
        "attn_bwd_dkdv": ("tensor", B * S * 6.0 * pair * D, "FLOP"),
        "attn_bwd_dq": ("tensor", B * S * 2.0 * pair * D, "FLOP"),
        "attn_bwd_fused": ("tensor", B * S * 8.0 * pair * D, "FLOP"),
        "gemm": ("hbm", B * (S + 0.35) * gemm_bytes, "B"),
        "update": ("hbm", B * S * 16.0 * L * D, "B"),
        "guidance": ("hbm", B * (S + 1) * 8.0 * n + (S + 1) * (8.0 * inst.nnz * 2 + 8.0 * inst.M + 4.0 * n), "B"),
    }


# --------------------------------------------------------------------------------------------------
# the reference's CPU path (the unmodified modules staged by oracle/stage_ref.py; else the oracle port) -- the only place
# bench.py executes oracle/
# --------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, args, cfg, inst, mip, kpm, Wd, Wdec):
        from oracle import ref_loader
        self.args, self.cfg, self.inst = args, cfg, inst
        self.kind = "reference" if ref_loader.available() else "port"
        self.Bc, self.Sc, self.S = args.cpu_samples, (args.ddim_steps if args.cpu_full else args.cpu_steps), args.ddim_steps
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        self.mipB, self.kpmB = t(mip)[None].repeat(self.Bc, 1, 1), t(kpm)[None].repeat(self.Bc, 1)
        self.b, self.c = t(inst.b), t(inst.c)
        from dip_b200 import synth
        self.x = t(synth.initial_noise(0, range(self.Bc), mip.shape[0]))
        if self.kind == "reference":
            R = ref_loader.load()
            n_den = cfg["n_den"]
            tr = R.DDPMTrainer(D, 1, n_den, "cpu")
            tr.load_state_dict({k: t(v) for k, v in Wd.items()}, strict=False)
            dec = R.MipConditionalDecorder(D, 1, 2, use_select_net="select_net.0.weight" in Wdec)
            dec.load_state_dict({k: t(v) for k, v in Wdec.items()})
            tr.eval(); dec.eval()
            self.R, self.tr, self.dec = R, tr, dec
            self.A = torch.sparse_coo_tensor(t(np.stack([inst.rows, inst.cols])), t(inst.a_val), size=(inst.M, inst.n))
        else:
            self.Wd, self.Wdec = {k: t(v) for k, v in Wd.items()}, {k: t(v) for k, v in Wdec.items()}
            self.A = (t(inst.rows), t(inst.cols), t(inst.a_val), inst.M)

    def sample(self):
        """(seconds of the guided steps, seconds of final decode + round + penalty) for Bc samples x Sc steps."""
        if self.kind == "reference":
            R = self.R
            s = R.DDIMSampler(self.tr, self.dec, gradient_scale=self.cfg["gs"], obj_guided_coef=self.cfg["lam"], device="cpu")
            s.make_schedule(self.S)
            s.predict_model, s.batch_size, s.key_padding_mask, s.conditions = self.tr.predict_model, self.Bc, self.kpmB, self.mipB
            x = self.x.clone()
            t0 = time.perf_counter()
            for k in range(self.Sc):                                   # the body of consraint_guided_ddim_sampling (diffusion.py:588-591)
                x, _ = s.consraint_guided_p_sample_ddim(x, self.S - 1 - k, k, self.A, self.b, self.c)
            t1 = time.perf_counter()
            import torch_sparse
            with torch.no_grad():                                      # sample.py:162-164, sample_partialsol.py:167-168
                pred, _ = self.dec(self.mipB, x, self.kpmB)
                xr = torch.round(pred).view(self.Bc, -1, 1)
                Ac = self.A.coalesce()
                torch.max(torch_sparse.spmm(Ac.indices(), Ac.values(), self.A.size()[0], self.A.size()[1], xr).squeeze(-1) - self.b,
                          torch.tensor(0)).sum(axis=-1)
            t2 = time.perf_counter()
            for p in self.dec.parameters():
                p.grad = None
        else:
            from oracle import restate
            tab = restate.ddim_tables(restate.trainer_buffers()["alphas_cumprod"], self.S)
            x = self.x
            t0 = time.perf_counter()
            for k in range(self.Sc):
                x = restate.guided_ddim_step(self.Wd, self.Wdec, x, self.S - 1 - k, restate.ddim_step_coeffs(tab, k), self.mipB, self.kpmB, self.A,
                                             self.b, self.c, self.cfg["gs"], self.cfg["lam"], dec_requires_grad=True)["x_prev"]
            t1 = time.perf_counter()
            restate.final_decode(self.Wdec, self.mipB, x, self.kpmB, self.A, self.b)
            t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def rate(self, n_rep=1):
        """samples/s scaled to S steps (per-step cost does not depend on the step index), measured seconds of the bounded sample."""
        best = None
        for _ in range(n_rep):
            ts, td = self.sample()
            total, measured = ts / self.Sc * self.S + td, ts + td
            if best is None or total < best[0]:
                best = (total, measured)
        return self.Bc / best[0], best[1]

    def describe(self, measured_s):
        return ("%d samples x %d of the %d guided DDIM steps + final decode + round + penalty through %s, measured %.1f s, scaled x%d/%d "
                "(per-step cost is step-independent)" % (self.Bc, self.Sc, self.S,
                "the unmodified reference modules (oracle/_ref)" if self.kind == "reference" else "the oracle port (oracle/restate.py)",
                measured_s, self.S, self.Sc))


# --------------------------------------------------------------------------------------------------
def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
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
    sm, mx, pw, reasons = [], 0, [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [s.strip() for s in ln.split(",")]
        if len(f) < 6:
            continue
        try:
            sm.append(float(f[0])); mx = max(mx, float(f[1]))
            if len(f) > 6:
                pw.append(float(f[6]))
        except ValueError:
            continue
        for nm, v in zip(names, f[2:6]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
            "power_w_median": float(np.median(pw)) if pw else None}


# --------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = CONFIGS[args.config]
    S = args.ddim_steps
    B = args.wave or cfg["wave"]
    from dip_b200 import synth
    Wd, Wdec, Wenc_np = load_weights(cfg)
    inst0 = make_instance(args, cfg["family"], seed=0)
    L = inst0.n
    workload = (cfg["what"] % {"M": inst0.M, "n": inst0.n, "nnz": inst0.nnz, "waves": 256 * (-(-1000 // B))}) + (
        "; guided DDIM S=%d gs=%g lambda=%g, D=128, denoiser %d layer(s), decoder 2 layers; step = one wave of %d samples"
        % (S, cfg["gs"], cfg["lam"], cfg["n_den"], B))
    config = {"workload": workload, "name": args.config, "wave": B, "ddim_steps": S}

    # ------------------------------------------------------------------ reference arm: the reference's own CPU path on the host cores
    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        mip = synth.mip_features(L, inst0.n, seed=10)
        kpm = np.zeros(L, dtype=bool)
        arm = CpuArm(args, cfg, inst0, mip, kpm, Wd, Wdec)
        for _ in range(args.warmup):
            arm.sample()
        rates, secs = [], []
        for _ in range(args.steps):
            r, m = arm.rate()
            rates.append(r); secs.append(m)
        value = float(np.mean(rates))
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": arm.kind, "sample": arm.describe(float(np.mean(secs)))},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "ms_per_step is the MEASURED wall time of one bounded sample (what this run spent per step); value extrapolates it to "
                    "S steps.  kind 'reference' = the unmodified /root/reference model/*.py (byte copies staged by oracle/stage_ref.py, "
                    "third-party torch_sparse / torch_geometric served by oracle/shims.py), all host cores, fp32.",
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
    import ctypes as C
    from dip_b200 import _lib, dist as ddist, handoff, schedule, ops
    from dip_b200.engine import Engine, randn
    from model.cmsp import CMSP
    from model.decoder import MipConditionalDecorder
    from model.diffusion import DDIMSampler, DDPMTrainer
    lib = _lib.load()
    K, W = args.steps, args.warmup
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tdev = lambda w: {k: T(v) for k, v in w.items()}
    steps = schedule.ddim_coeffs(S)

    # the reference-facing model objects (sample.py:93-117) and the engine they share
    trainer = DDPMTrainer(D, 1, cfg["n_den"], dev)
    trainer.load_state_dict(tdev(Wd), strict=False)
    decoder = MipConditionalDecorder(D, 1, 2, use_select_net="select_net.0.weight" in Wdec).to(dev)
    decoder.load_state_dict(tdev(Wdec))
    cmsp = CMSP(emb_num=3, emb_dim=D, n_heads=1, n_layers=1, padding_len=L).to(dev)
    cmsp.load_state_dict(tdev(Wenc_np), strict=False)
    trainer.eval(); decoder.eval(); cmsp.eval()
    sampler = DDIMSampler(trainer, decoder, gradient_scale=cfg["gs"], obj_guided_coef=cfg["lam"], device=dev)
    sampler.keep_intermediates = True            # reference behaviour (S+1 latents returned)
    eng = Engine()
    eng.set_denoiser(tdev(Wd)); eng.set_decoder(tdev(Wdec))

    # ---- work items of this rank
    spi = cfg["spi"]
    if args.config == "sc1":
        lo, hi = (spi * rank) // world, (spi * (rank + 1)) // world               # strong scaling: one instance, samples split
        inst_ids = [0]
        waves = [(0, s, min(s + B, hi)) for s in range(lo, hi, B)]                # (instance, sample_begin, sample_end)
        K = len(waves)
        total_samples_all = spi
    elif args.config == "partialsol":
        n_inst_rank = max(1, (K * B + spi - 1) // spi)                            # enough instances to fill K waves of B samples
        inst_ids = [rank * 10_000 + i for i in range(n_inst_rank)]
        waves = None
        total_samples_all = world * K * B
    else:
        inst_ids = [rank]                                                          # weak scaling: every rank its own instance and waves
        waves = [(rank, k * B, (k + 1) * B) for k in range(K)]
        total_samples_all = world * K * B

    insts = {i: (inst0 if i == 0 else make_instance(args, cfg["family"], seed=i)) for i in inst_ids}
    kpm = np.zeros(L, dtype=bool)

    class Bigraph:      # the fields CMSP.get_features reads (dataset.py:12-34), host side
        def __init__(self, inst):
            self.host = {"constraint_features": inst.constraint_features, "edge_index": inst.edge_index, "edge_attr": inst.edge_attr,
                         "variable_features": inst.variable_features, "int_indices": inst.int_indices}
            self.n_int_vars = inst.n_int_vars

        def to(self, device):
            for k, v in self.host.items():
                setattr(self, k, torch.from_numpy(np.ascontiguousarray(v)).to(device, non_blocking=True))
            return self

    def encode(inst):
        mip, _, mask = cmsp.get_features(Bigraph(inst).to(dev))
        return mip[0].contiguous(), mask[0].contiguous()

    def prepare(inst):
        mip, mask = encode(inst)
        return eng.prepare_instance(mip, mask, inst.rows, inst.cols, inst.a_val, inst.M, inst.b, inst.c,
                                    a_int=inst.a_int, b_int=inst.b_int, c_int=inst.c_int)

    def noise(inst_id, lo, hi):
        return randn((hi - lo, L, D), seed=1000 + inst_id, subseq=lo, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- the pass over this rank's waves, inputs resident (one library call per wave)
    def run_waves(record=None):
        outs = []
        if args.config == "partialsol":
            # 15 samples per instance, waves of B samples mixing instances (a wave may start / end in the middle of an instance)
            preps = {i: prepare(insts[i]) for i in inst_ids}
            flat = [(i, s) for i in inst_ids for s in range(spi)][:K * B]
            for k in range(K):
                chunk = flat[k * B:(k + 1) * B]
                ids = sorted(set(i for i, _ in chunk))
                eng.use_instances([preps[i] for i in ids], [ids.index(i) for i, _ in chunk])
                x = torch.cat([noise(i, s, s + 1) for i, s in chunk])
                o = eng.sample_guided_ddim(x, steps, cfg["gs"], cfg["lam"])
                # hand-off (sample_partialsol.py:171-199): coverage subset + values, bit-packed, per instance group
                o["partial"] = [handoff.select_gpu(o["xbits"][[j for j, (ii, _) in enumerate(chunk) if ii == i]][:, :insts[i].n].contiguous(), 0.2,
                                                   seed=7, instance_id=i, sample_base=min(s for ii, s in chunk if ii == i), reference_quirk=False)
                                for i in ids]
                o["chunk"] = chunk
                outs.append(o)
            return outs
        cur = None
        for (i, lo, hi) in waves:
            if cur != i:
                eng.use_instance(prepare(insts[i]))                               # encoder (K3) + CSR build: once per instance
                cur = i
            o = eng.sample_guided_ddim(noise(i, lo, hi), steps, cfg["gs"], cfg["lam"])
            o["item"] = (i, lo, hi)
            outs.append(o)
        return outs

    for w in range(W):
        run_waves()
    rec = ddist.make_records(max(world, 1) if args.config != "partialsol" else len(inst_ids), inst0.n)
    stop, clk_lines = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, clk_lines, local_rank), daemon=True)
    th.start()
    time.sleep(0.3)

    # ---- timed region 1: inputs resident in HBM ----
    barrier()
    launches0 = lib.dip_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = run_waves()
    n_feas = 0
    packed_bytes = 0
    for o in outs:
        vi = o["viol_int"].cpu().numpy()
        n_feas += int((vi == 0).sum())
        if args.config == "partialsol":
            for (vals, mask) in o["partial"]:
                packed_bytes += vals.cpu().numpy().nbytes + mask.cpu().numpy().nbytes
        else:
            ddist.update_record(rec, o["item"][0] if args.config != "sc1" else 0, o["xbits"].cpu().numpy(), vi, o["obj_int"].cpu().numpy())
    if args.config != "partialsol":
        ddist.all_gather_records(rec, dev)                                         # the path's single collective
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.dip_launch_count() - launches0
    my_samples = sum(int(o["viol_int"].numel()) for o in outs)

    # ---- kernel durations: one more wave of the same shape with the in-library CUDA-event timer on (an event pair around every launch,
    # on the launch stream).  Kept out of the pass above because the event pairs cost ~4 % of throughput.
    barrier()
    Bp = int(outs[0]["viol_int"].numel())
    xprof = noise(inst_ids[0], 500_000, 500_000 + Bp)
    if args.config == "partialsol":
        eng.use_instance(prepare(insts[inst_ids[0]]))
    lib.dip_prof_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    eng.sample_guided_ddim(xprof, steps, cfg["gs"], cfg["lam"])
    p1.record()
    torch.cuda.synchronize()
    lib.dip_prof_enable(0)
    ms_prof_wave = p0.elapsed_time(p1)
    ncls = len(_lib.PROF_CLASSES)
    pms = (C.c_double * ncls)(); pcnt = (C.c_uint64 * ncls)()
    _lib.check(lib.dip_prof_collect(pms, pcnt, ncls))
    prof = {nm: {"ms": pms[i], "launches": int(pcnt[i])} for i, nm in enumerate(_lib.PROF_CLASSES)}

    # ---- timed region 2: end to end through the reference-facing API, host buffers, copies inside ----
    ms_e2e = ms_e2e_engine = None
    n_feas_e2e = 0
    h2d = d2h = 0
    if not args.no_api_e2e and args.config not in ("partialsol",):
        inst_e = insts[inst_ids[0]]
        Kw = len(waves)
        host_x = [torch.empty((hi - lo, L, D), dtype=torch.float32).pin_memory() for (_, lo, hi) in waves]
        for hx, (i, lo, hi) in zip(host_x, waves):
            hx.copy_(noise(i, 700_000 + lo, 700_000 + hi).cpu())
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        hostA = {"idx": pin(np.stack([inst_e.rows, inst_e.cols])), "val": pin(inst_e.a_val), "b": pin(inst_e.b), "c": pin(inst_e.c)}

        def host_bufs(bb):
            return {"xbits": torch.empty((bb, inst_e.n), dtype=torch.uint8).pin_memory(), "penalty": torch.empty(bb, dtype=torch.float32).pin_memory(),
                    "viol_int": torch.empty(bb, dtype=torch.int64).pin_memory(), "obj_int": torch.empty(bb, dtype=torch.int64).pin_memory()}
        host_out = [host_bufs(B), host_bufs(B)]                  # double-buffered: wave k is read back while wave k+1 is queued
        done = [torch.cuda.Event(), torch.cuda.Event()]

        def collect(k, bb):
            done[k & 1].synchronize()
            return int((host_out[k & 1]["viol_int"][:bb] == 0).sum())

        def api_pass():
            """sample.py:136-168 per instance / wave, with this repo's drop-in classes."""
            nf = 0
            # instance -> device (sample.py:139-148) and its embedding (sample.py:150): once per instance
            bigraph = Bigraph(inst_e).to(dev)
            A = torch.sparse_coo_tensor(hostA["idx"].to(dev, non_blocking=True), hostA["val"].to(dev, non_blocking=True), size=(inst_e.M, inst_e.n))
            b_d, c_d = hostA["b"].to(dev, non_blocking=True), hostA["c"].to(dev, non_blocking=True)
            mip_features, _, key_padding_mask = cmsp.get_features(bigraph)
            for k, (i, lo, hi) in enumerate(waves):
                bb = hi - lo
                cond = mip_features.repeat((bb, 1, 1))                                        # sample.py:151-152
                kp = key_padding_mask.repeat((bb, 1))
                sampler.initial_noise = host_x[k].to(dev, non_blocking=True)                   # H2D of this wave's noise (pinned); the reference's hook :584
                x, _ = sampler.constraint_guided_sample(cond, kp, A, b_d, c_d, S=S)           # sample.py:157
                pred, _ = decoder(cond, x, kp)                                                # sample.py:162-163
                prep = sampler.last_instance
                p_full = torch.zeros((bb, L), dtype=torch.float32, device=dev)
                p_full[~kp] = pred
                csr = dict(prep["keep"], nnz=prep["info"]["nnz"])
                r = ops.round_feasibility(p_full, prep["keep"]["pos"], csr, prep["keep"]["b"], inst_e.M)              # round + penalty, sample.py:164-168
                for nm in ("xbits", "penalty"):
                    host_out[k & 1][nm][:bb].copy_(r[nm], non_blocking=True)                  # D2H of this wave's result
                host_out[k & 1]["viol_int"][:bb].copy_((r["penalty"] > 1e-5).to(torch.int64), non_blocking=True)     # sample.py:174
                done[k & 1].record()
                if k >= 1:
                    nf += collect(k - 1, waves[k - 1][2] - waves[k - 1][1])
            nf += collect(Kw - 1, waves[-1][2] - waves[-1][1])
            return nf

        def engine_pass():
            nf = 0
            eng.use_instance(prepare(inst_e))
            for k, (i, lo, hi) in enumerate(waves):
                bb = hi - lo
                o = eng.sample_guided_ddim(host_x[k].to(dev, non_blocking=True), steps, cfg["gs"], cfg["lam"])
                for nm in host_out[k & 1]:
                    host_out[k & 1][nm][:bb].copy_(o[nm], non_blocking=True)
                done[k & 1].record()
                if k >= 1:
                    nf += collect(k - 1, waves[k - 1][2] - waves[k - 1][1])
            nf += collect(Kw - 1, waves[-1][2] - waves[-1][1])
            return nf

        api_pass() if W > 0 else None                         # warm the API path (allocator, caches) once
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        n_feas_e2e = api_pass()
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        engine_pass()
        g1.record()
        barrier()
        ms_e2e_engine = g0.elapsed_time(g1)
        inst_bytes = sum(v.nbytes for v in Bigraph(inst_e).host.values()) + sum(v.numel() * v.element_size() for v in hostA.values())
        h2d = int(np.mean([hx.numel() * 4 for hx in host_x])) + inst_bytes // max(Kw, 1)
        d2h = B * inst_e.n + B * 4 + B * 8
    stop.set()
    th.join(timeout=3)

    vals = [ms, ms_e2e or 0.0, ms_e2e_engine or 0.0]
    t_ms = torch.tensor(vals, dtype=torch.float64, device=dev)
    cnt = torch.tensor([n_feas, n_feas_e2e, my_samples], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
    ms, ms_e2e_max, ms_e2e_eng_max = float(t_ms[0]), float(t_ms[1]), float(t_ms[2])
    total_samples = int(cnt[2])
    value = total_samples / (ms / 1e3)

    if rank == 0:
        pk = peaks()
        den, dec_f, dec_b = step_flops(L, cfg["n_den"])
        work = class_work(L, inst0.n, Bp, cfg["n_den"], 2, S, inst0)
        traffic, traffic_src = ncu_traffic()
        total_prof_ms = sum(v["ms"] for v in prof.values())
        roof_all = {}
        for cls, (bound, amount, unit) in work.items():
            pv = prof.get(cls)
            if not pv or pv["launches"] == 0 or pv["ms"] <= 0:
                continue
            rate = amount / (pv["ms"] * 1e-3)
            peak = pk["tflops"] * 1e12 if bound == "tensor" else pk["hbm_gbs"] * 1e9
            tr = dict(traffic.get(cls, {}))
            cap = traffic.get("_capture", {"config": "sc", "wave": 37})
            # DRAM bytes per launch were captured at cap["wave"] samples on the Set-Cover shapes: per-launch traffic of every kernel here is
            # proportional to the wave; other shapes have no capture
            if tr.get("dram_bytes_per_launch") is not None:
                tr["dram_bytes_per_launch"] = (tr["dram_bytes_per_launch"] * Bp / cap["wave"]) if cfg["family"] == "SC" else None
            roof_all[cls] = {"bound": bound, "achieved": rate / (1e12 if bound == "tensor" else 1e9), "peak": peak / (1e12 if bound == "tensor" else 1e9),
                             "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": rate / peak, "launches": pv["launches"],
                             "launch_ms": pv["ms"] / pv["launches"], "share_of_step": pv["ms"] / total_prof_ms,
                             "traffic": tr.get("dram_bytes_per_launch"), "traffic_kernel": tr.get("kernel")}
        tensor_classes = [c for c in roof_all if roof_all[c]["bound"] == "tensor"]
        dom = max(tensor_classes, key=lambda c: roof_all[c]["share_of_step"]) if tensor_classes else None
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / max(K, 1),
            "higher_is_better": True, "scaling": "strong" if args.config == "sc1" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, instance_per_rank=args.config != "sc1",
                           l2="working set per wave (%.1f GB) exceeds the 126 MB L2; no flush needed" % (eng.workspace(Bp)[1] / 1e9),
                           precision_mode="attention and Linear layers on tcgen05 tensor cores with bf16x2-split fp32 operands (hi*hi + hi*lo + lo*hi, "
                                          "fp32 TMEM accumulation; one-step latent error ~1e-5 rel vs the fp32 reference); guidance, DDIM update, rounding "
                                          "and the integer feasibility check in fp32/int64 SIMT"),
            "feasible_per_s": float(cnt[0]) / (ms / 1e3), "feasible_ratio": float(cnt[0]) / max(total_samples, 1),
            "tflops_effective": (den + dec_f + dec_b) * S * total_samples / (ms / 1e3) / 1e12,
            "tflops_effective_note": "reference-equivalent FLOPs (SURVEY 8d: what the reference executes per sample-step, dead tokens included) / time",
            "gpu_launches": int(launches),
            "kernel_time_share": {k: (v["ms"] / total_prof_ms if total_prof_ms > 0 else 0.0) for k, v in prof.items()},
            "kernel_timing": {"how": "one extra wave with a CUDA-event pair around every launch (dip_prof_*), after the timed region",
                              "wave_ms_with_events": ms_prof_wave, "sum_of_kernels_ms": total_prof_ms,
                              "gpu_busy_frac_of_timed_wave": total_prof_ms / (ms / max(K, 1)) if ms > 0 else None},
            "roofline_all": roof_all,
            "roofline_traffic_source": traffic_src,
            "clocks": summarize_clocks(clk_lines),
        }
        if dom:
            r = roof_all[dom]
            line["roofline"] = {"kernel": dom, "bound": r["bound"], "achieved": r["achieved"], "peak": r["peak"], "unit": r["unit"], "frac": r["frac"],
                                "traffic": r["traffic"], "launch_ms": r["launch_ms"], "share_of_step": r["share_of_step"],
                                "peak_source": pk["src"] + ", sustained bf16",
                                "ceiling_note": "fp32-grade products cost 3 bf16 MMA passes: attainable ceiling = 1/3 of the bf16 peak times (credited / executed products)"}
        if ms_e2e is not None:
            e2e_value = total_samples / (ms_e2e_max / 1e3)
            line["e2e"] = {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                           "feasible_ratio": float(cnt[1]) / max(total_samples, 1),
                           "path": "model.cmsp.CMSP.get_features -> model.diffusion.DDIMSampler.constraint_guided_sample -> model.decoder.MipConditionalDecorder "
                                   "-> round + penalty (sample.py:136-168), pinned host buffers, H2D of instance + noise and D2H of assignments inside the timed region",
                           "engine_one_call": {"value": total_samples / (ms_e2e_eng_max / 1e3), "unit": "samples/s",
                                               "path": "Engine.sample_guided_ddim (one library call per wave), same host buffers"}}
        else:
            line["e2e"] = {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(packed_bytes // max(K, 1)),
                           "path": "partial solutions (bit-packed values + selection masks) copied to the host inside the timed region; inputs generated on the device"}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            mip_host = encode(inst0)[0].cpu().numpy()
            arm = CpuArm(args, cfg, inst0, mip_host, kpm, Wd, Wdec)
            arm.sample()                                                           # warm-up
            rate, measured = arm.rate(n_rep=2)
            line["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": cores, "kind": arm.kind, "sample": arm.describe(measured)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
