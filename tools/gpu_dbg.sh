#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_fused.py::test_fused_gmres_matches_oracle_and_unfused[case7]" -x -q > gpurun_out/dbg_memcheck.log 2>&1
grep -n "=========" gpurun_out/dbg_memcheck.log | head -60
