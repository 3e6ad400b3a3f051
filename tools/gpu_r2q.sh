#!/bin/bash
# end-of-round check (1 GPU): whole GPU test suite + smoke, the default bench line, and the additive Schwarz lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest_gpu.log
tail -4 gpurun_out/r2q_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 10 600 python bench.py > gpurun_out/r2q_bench_c2.json 2> gpurun_out/r2q_bench_c2.err; echo "bench rc=$?"
timeout -k 10 600 python bench.py --ksp bcgs --pc asm --no-cpu-baseline > gpurun_out/r2q_bench_c2_bcgs_asm.json 2> gpurun_out/r2q_bench_c2_bcgs_asm.err; echo "bcgs+asm rc=$?"
timeout -k 10 600 python bench.py --pc asm --no-cpu-baseline > gpurun_out/r2q_bench_c2_gmres_asm.json 2> gpurun_out/r2q_bench_c2_gmres_asm.err; echo "gmres+asm rc=$?"
for f in c2 c2_bcgs_asm c2_gmres_asm; do grep -a '^{' gpurun_out/r2q_bench_$f.json | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('$f', round(d['value'],3), 'steps/s', c['ksp_iterations_per_step'], 'its', c['us_per_ksp_iteration'], 'us/it', 'e2e', round(d['e2e']['value'],3), 'roofline', d['roofline']['frac'], 'parity', d['parity']['residual_relerr'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
"; done
