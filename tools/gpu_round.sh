#!/bin/bash
# One gpurun call: GPU tests, the bench line of every config, ncu capture.  Usage: bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest.log | tail -n 3; grep -E "^(FAILED|ERROR)" $OUT/${TAG}_pytest.log | head -n 20
python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_sc.json 2> $OUT/${TAG}_bench_sc.err; echo "bench sc rc=$?"; tail -c 300 $OUT/${TAG}_bench_sc.err
python -c "
import json,sys
l=json.loads(open('$OUT/${TAG}_bench_sc.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'],'engine',l['e2e'].get('engine_one_call'),'cpu',l.get('cpu_baseline',{}).get('value'),l.get('cpu_baseline',{}).get('kind'))
print('share',{k:round(v,3) for k,v in l['kernel_time_share'].items()})
print('roof',{k:(round(v['frac'],3),round(v['launch_ms'],3)) for k,v in l['roofline_all'].items()})
print('clocks',l['clocks'])
"
if [ "${2:-}" != "quick" ]; then
for c in is ca cf partialsol sc1; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; echo "bench $c rc=$?"; tail -c 300 $OUT/${TAG}_bench_$c.err
  python -c "
import json
l=json.loads(open('$OUT/${TAG}_bench_$c.json').read().strip().splitlines()[-1])
print('$c value',l['value'],'feasible_ratio',l['feasible_ratio'],'e2e',l['e2e']['value'],'ms_per_step',l['ms_per_step'])
"
done
python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"; tail -c 400 $OUT/${TAG}_bench_ref.json
if [ "${2:-}" != "noncu" ]; then bash tools/ncu_capture.sh $TAG; fi
fi
nproc > $OUT/${TAG}_nproc.txt
du -sh $OUT
