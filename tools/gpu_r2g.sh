#!/bin/bash
# round-2 GPU job G (1 GPU): row-major ILU level records (both solvers), persistent kernel probes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout -k 10 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_linalg.py tests/test_minc.py tests/test_gpu_newton.py tests/test_gpu_wce.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -5
for d in "50 50 50" "100 100 100"; do
  timeout -k 10 200 python tools/fused_probe.py --dims $d 2>&1 | tail -1
  WB_FUSED=0 timeout -k 10 200 python tools/fused_probe.py --dims $d 2>&1 | tail -1
done
timeout -k 10 300 python tools/microbench.py --skip-pcs --its 200 2>&1 | grep -v "^ *$" | tail -40
