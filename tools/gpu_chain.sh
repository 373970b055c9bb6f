#!/bin/bash
# chain kernel bring-up: its test under a short timeout, then the whole GPU suite, then A/B bench lines
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k block_chain > $OUT/${TAG}_chain_pytest.log 2>&1; rc=$?; echo "chain pytest rc=$rc"; tail -n 25 $OUT/${TAG}_chain_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest.log | tail -n 3; grep -E "^(FAILED|ERROR)" $OUT/${TAG}_pytest.log | head -n 20
for v in 0 1; do
DIP_NO_CHAIN=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_sc_nochain$v.json 2> $OUT/${TAG}_bench_sc_nochain$v.err; echo "bench nochain=$v rc=$?"; tail -c 300 $OUT/${TAG}_bench_sc_nochain$v.err
python -c "
import json,sys
l=json.loads(open('$OUT/${TAG}_bench_sc_nochain$v.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'],'engine',l['e2e'].get('engine_one_call'))
print('share',{k:round(v,3) for k,v in l['kernel_time_share'].items()})
print('roof',{k:(round(v['frac'],3),round(v['launch_ms'],3)) for k,v in l['roofline_all'].items()})
print('clocks',l['clocks'])
"
done
