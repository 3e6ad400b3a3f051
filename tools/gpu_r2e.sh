#!/bin/bash
# round-2 GPU job E (1 GPU): persistent kernel with LL transport; per-GPU-sized problems (50^3 = what one of 8 GPUs holds)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout -k 10 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
for d in "50 50 50" "100 50 50" "100 100 50" "100 100 100"; do
  timeout -k 10 200 python tools/fused_probe.py --dims $d 2>&1 | tail -1
  WB_FUSED=0 timeout -k 10 200 python tools/fused_probe.py --dims $d 2>&1 | tail -1
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_gmres_fused -c 1 -o gpurun_out/r2e_fused50 \
  python tools/fused_probe.py --dims 50 50 50 --maxit 150 --reps 1 > gpurun_out/r2e_ncu.log 2>&1; tail -2 gpurun_out/r2e_ncu.log
