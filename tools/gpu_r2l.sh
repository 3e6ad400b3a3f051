#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_wce.py tests/test_minc.py tests/test_gpu_edge_cases.py tests/test_gpu_methods.py tests/test_deliverability.py tests/test_gpu_newton.py tests/test_gpu_fullsize.py tests/test_gpu_fused.py -x -q 2>&1 | tail -4
for c in 2 4; do
for l in 1 0; do
WB_JAC_LANES=$l timeout -k 10 300 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().split('\n') if l.startswith('{')][-1])
print('config $c lanes=$l', round(d['value'],3), 'steps/s', d['config']['us_per_ksp_iteration'], 'us/it', {k:v['ms'] for k,v in d['phases_ms'].items()})"
done; done
