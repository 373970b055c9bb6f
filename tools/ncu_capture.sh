#!/bin/bash
# Run on the GPU box (under gpurun): launch list + `ncu --set full` of every kernel class of one wave.  gpurun brings back at most
# 64 MiB, so the reports are turned into raw-metric CSVs on the box and only the (small) attention report is kept.
#   bash tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_launches.csv, <tag>_{attn,gemm,hbm}_raw.csv, <tag>_attn.ncu-rep
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/prof_step.py --steps 1 --wave ${WAVE:-148}"
$CMD > $OUT/${TAG}_prof_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 $OUT/${TAG}_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
capture() {   # name, kernel regex, count, extra flags
    ncu --set full --clock-control none $4 -k regex:"$2" -c $3 -o /tmp/${TAG}_$1 -f $CMD > $OUT/${TAG}_ncu_$1.log 2>&1
    echo "$1 rc=$?"
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
    ls -la /tmp/${TAG}_$1.ncu-rep
}
# attention kernels of one step (1 denoiser fwd, 2 decoder fwd, 2 + 2 bwd) -- report kept for the source page
capture attn 'attn_' 7 "--import-source on"
[ $(stat -c %s /tmp/${TAG}_attn.ncu-rep) -lt 30000000 ] && cp /tmp/${TAG}_attn.ncu-rep $OUT/
# GEMM / LayerNorm / row-chain kernels of the step
capture gemm 'gemm_tc|ln_|layernorm|delta_kernel|chain_' 40 ""
# HBM-bound kernels: K1 update, K2 guidance + rounding, K3 encoder, head, time token
capture hbm 'ddim_step|guidance|round_feas|conv_reduce|spmm_csr|head_|time_token|gather_pad|linear_smallk|gemm_nt' 60 ""
du -sh $OUT
