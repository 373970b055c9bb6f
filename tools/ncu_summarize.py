"""Turn `ncu --page raw --csv` exports (tools/ncu_capture.sh) into the two files the repo keeps under profiles/:

    python tools/ncu_summarize.py r02 gpurun_out/r02b_attn_raw.csv gpurun_out/r02b_gemm_raw.csv gpurun_out/r02b_hbm_raw.csv

  profiles/<tag>_ncu_launches.json   one record per profiled launch: kernel, grid, duration, DRAM bytes read / written, DRAM and
                                     tensor-pipe utilisation, L2 hit rate, registers, achieved occupancy
  profiles/<tag>_ncu_kernels.json    per kernel class (the classes of dip_prof_*): mean DRAM bytes per launch -- what bench.py reports
                                     as roofline.traffic -- with the kernel name and the number of launches averaged
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
CLASS = [("attn_fwd_tc_kernel", "attn_fwd"), ("attn_bwd_kv_tc_kernel", "attn_bwd_dkdv"), ("attn_bwd_dq_tc_kernel", "attn_bwd_dq"), ("attn_bwd_dq_stream_tc_kernel", "attn_bwd_dq_stream"),
         ("attn_bwd_fused_tc_kernel", "attn_bwd_fused"), ("chain_", "chain"), ("gemm_tc_kernel", "gemm"), ("gemm_nt_kernel", "gemm_simt"),
         ("ln_", "layernorm"), ("ddim_step_kernel", "update"), ("guidance_kernel", "guidance"), ("round_feas_kernel", "round_feas"),
         ("conv_reduce_kernel", "gnn_conv"), ("spmm_csr_kernel", "gnn_spmm"), ("head_", "head"), ("delta_kernel", "delta")]
WANT = {
    "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "duration_us": "gpu__time_duration.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread", "sm_mhz": "sm__cycles_elapsed.avg.per_second",
    "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_busy_cycles": "sm__cycles_active.avg", "sm_elapsed_cycles": "sm__cycles_elapsed.avg",
}


def classify(name):
    for key, cls in CLASS:
        if key in name:
            return cls
    return "other"


def read(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        rec = {"kernel": r[col["Kernel Name"]].split("(")[0], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k, m in WANT.items():
            i = col.get(m)
            if i is None or r[i] == "":
                rec[k] = None
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if k in ("dram_read", "dram_write", "duration_us"):
                v *= UNIT.get(u, 1.0)
            if k == "sm_mhz":
                v *= {"Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6}.get(u, 1.0)
            rec[k] = v
        rec["class"] = classify(rec["kernel"])
        out.append(rec)
    return out


def main():
    tag, paths = sys.argv[1], sys.argv[2:]
    launches = []
    for p in paths:
        launches += read(p)
    per = {}
    for r in launches:
        if r["dram_read"] is None:
            continue
        per.setdefault(r["class"], []).append(r)
    kernels = {}
    for cls, rs in per.items():
        kernels[cls] = {"kernel": rs[0]["kernel"], "launches_profiled": len(rs),
                        "dram_bytes_per_launch": sum(r["dram_read"] + r["dram_write"] for r in rs) / len(rs),
                        "duration_us_under_ncu": sum(r["duration_us"] for r in rs) / len(rs),
                        "tensor_pct_elapsed": sum((r["tensor_pct_elapsed"] or 0) for r in rs) / len(rs),
                        "dram_pct": sum((r["dram_pct"] or 0) for r in rs) / len(rs)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(launches, open(os.path.join(ROOT, "profiles", tag + "_ncu_launches.json"), "w"), indent=0)
    # the shapes the capture ran on (tools/ncu_capture.sh -> tools/prof_step.py): bench.py scales the per-launch bytes to its own wave
    kernels["_capture"] = {"config": "sc", "wave": int(os.environ.get("WAVE", 148)), "command": "python tools/prof_step.py --steps 1 --wave <wave>"}
    json.dump(kernels, open(os.path.join(ROOT, "profiles", tag + "_ncu_kernels.json"), "w"), indent=1)
    for cls, k in sorted(kernels.items()):
        if cls.startswith("_"):
            continue
        print("%-16s %-28s n=%3d  dram %8.1f MB/launch  %8.1f us  tensor %5.1f%%  dram %5.1f%%" % (
            cls, k["kernel"][:28], k["launches_profiled"], k["dram_bytes_per_launch"] / 1e6, k["duration_us_under_ncu"], k["tensor_pct_elapsed"], k["dram_pct"]))


if __name__ == "__main__":
    main()
