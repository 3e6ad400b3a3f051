#!/bin/bash
# round-2 GPU job B (1 GPU): first runs of the persistent GMRES kernel -- parity tests under a timeout, then timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout -k 10 300 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2b_fused_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_fused_tests.log
tail -25 gpurun_out/r2b_fused_tests.log
timeout -k 10 300 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_newton.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2b_more_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_more_tests.log
tail -8 gpurun_out/r2b_more_tests.log
timeout -k 10 300 python tools/microbench.py --skip-pcs --its 200 > gpurun_out/r2b_micro.json 2> gpurun_out/r2b_micro.err; tail -c 800 gpurun_out/r2b_micro.err; cat gpurun_out/r2b_micro.json
timeout -k 10 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; tail -c 600 gpurun_out/r2b_bench_c2.err; head -c 2500 gpurun_out/r2b_bench_c2.json
timeout -k 10 400 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2b_bench_c4.json 2> gpurun_out/r2b_bench_c4.err; tail -c 600 gpurun_out/r2b_bench_c4.err; head -c 2500 gpurun_out/r2b_bench_c4.json
timeout -k 10 400 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; tail -c 600 gpurun_out/r2b_bench_c5.err; head -c 2500 gpurun_out/r2b_bench_c5.json
