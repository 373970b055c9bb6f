#!/bin/bash
# A/B of an environment switch on the bench line: bash tools/gpu_ab.sh VAR tag
VAR=$1; TAG=${2:-ab}
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_parity2.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -n 3 $OUT/${TAG}_pytest.log
if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR)|Error|assert" $OUT/${TAG}_pytest.log | head -n 20; fi
for v in 1 0 1 0; do
env $VAR=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-api-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$VAR=$v value',round(l['value'],2),'mhz',l['clocks']['sm_mhz'],'W',l['clocks']['power_w_median'],'fwd',round(l['roofline_all']['attn_fwd']['launch_ms'],4),'dkdv',round(l['roofline_all']['attn_bwd_dkdv']['launch_ms'],4),'share_fwd',round(l['kernel_time_share']['attn_fwd'],3))
"
done
