#!/bin/bash
# Run on the GPU box (under gpurun): launch list + `ncu --set full` of every kernel class of one wave.
#   bash tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{attn,gemm,hbm}.ncu-rep
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/prof_step.py --steps 2 --wave 37"
$CMD > $OUT/${TAG}_prof_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 $OUT/${TAG}_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# attention kernels: every launch of the 2 steps + final decode (3 fwd, 2+2 bwd per step)
ncu --set full --clock-control none --import-source on -k regex:'attn_' -c 16 -o $OUT/${TAG}_attn -f $CMD > $OUT/${TAG}_ncu_attn.log 2>&1
echo "attn rc=$?"
# GEMM / LayerNorm kernels of the first step
ncu --set full --clock-control none -k regex:'gemm_tc|ln_|layernorm|delta_kernel' -c 48 -o $OUT/${TAG}_gemm -f $CMD > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "gemm rc=$?"
# HBM-bound kernels: K1 update, K2 guidance + rounding, K3 encoder, head, time token
ncu --set full --clock-control none -k regex:'ddim_step|guidance|round_feas|conv_reduce|spmm_csr|head_|time_token|gather_pad|linear_smallk|gemm_nt' -c 60 -o $OUT/${TAG}_hbm -f $CMD > $OUT/${TAG}_ncu_hbm.log 2>&1
echo "hbm rc=$?"
ls -la $OUT/${TAG}_*
