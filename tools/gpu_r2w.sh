#!/bin/bash
# BiCGStab in the persistent kernel: parity tests, then the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 300 python -m pytest tests/test_gpu_fused.py -q -x -k bcgs 2>&1 | tail -3 | cut -c1-300
timeout -k 10 200 python bench.py --ksp bcgs --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2w_bcgs_fused1.json 2> gpurun_out/r2w_bcgs_fused1.err
grep -a '^{' gpurun_out/r2w_bcgs_fused1.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('fused bcgs', round(d['value'],3), 'steps/s', c['ksp_iterations_per_step'], 'its', c['us_per_ksp_iteration'], 'us/it', 'launches', d['gpu_launches'], 'reason', c['ksp_reason'], c['ksp_rnorm'], d.get('ksp_breakdown_us_per_iteration'))
" || tail -3 gpurun_out/r2w_bcgs_fused1.err
