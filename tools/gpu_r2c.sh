#!/bin/bash
# round-2 GPU job C (1 GPU): persistent GMRES after the producer / SpMV changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout -k 10 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_linalg.py tests/test_minc.py -x -q > gpurun_out/r2c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_tests.log
tail -8 gpurun_out/r2c_tests.log
timeout -k 10 300 python tools/microbench.py --skip-pcs --its 200 > gpurun_out/r2c_micro.json 2> gpurun_out/r2c_micro.err; tail -c 800 gpurun_out/r2c_micro.err; cat gpurun_out/r2c_micro.json
for c in 4 5; do
timeout -k 10 400 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c$c.json 2> gpurun_out/r2c_bench_c$c.err; tail -c 600 gpurun_out/r2c_bench_c$c.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_c$c.json"))
print($c, d["value"], d["ms_per_step"], d["config"]["ksp_iterations_per_step"], d["config"]["us_per_ksp_iteration"], d.get("ksp_breakdown_us_per_iteration"), d["gpu_launches"])
PY
done
