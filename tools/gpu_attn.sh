#!/bin/bash
# attention-kernel iteration: the tensor-core tests under a timeout, then a quick bench line
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 4 $OUT/${TAG}_pytest.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR)|Error|assert" $OUT/${TAG}_pytest.log | head -n 20; exit 0; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_sc.json 2> $OUT/${TAG}_bench_sc.err; echo "bench rc=$?"
python -c "
import json,sys
l=json.loads(open('$OUT/${TAG}_bench_sc.json').read().strip().splitlines()[-1])
print('value',l['value'],'e2e',l['e2e']['value'])
print('share',{k:round(v,3) for k,v in l['kernel_time_share'].items()})
print('roof',{k:(round(v['frac'],3),round(v['launch_ms'],3)) for k,v in l['roofline_all'].items()})
print('clocks',l['clocks'])
"
