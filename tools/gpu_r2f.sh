#!/bin/bash
# round-2 GPU job F (1 GPU): persistent kernel with replicated state; per-GPU-sized problems
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout -k 10 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_linalg.py -x -q 2>&1 | tail -3
for d in "50 50 50" "100 50 50" "100 100 50" "100 100 100"; do
  timeout -k 10 200 python tools/fused_probe.py --dims $d 2>&1 | tail -1
done
