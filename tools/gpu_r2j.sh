#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 600 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_newton.py tests/test_gpu_tracer.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py tests/test_gpu_wce.py tests/test_minc.py -x -q 2>&1 | tail -4
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -k 10 300 python tools/microbench.py --skip-pcs --its 200 2>&1 | grep -E "spmv_us|us_per_it|pc_apply" 
WB_SPMV_SELL=0 timeout -k 10 300 python tools/microbench.py --skip-pcs --its 200 2>&1 | grep -E "spmv_us|us_per_it|pc_apply"
for c in 4 5; do timeout -k 10 300 python tools/microbench.py --config $c --skip-pcs --its 200 2>&1 | grep -E "spmv_us|us_per_it|pc_apply|workload"; done
