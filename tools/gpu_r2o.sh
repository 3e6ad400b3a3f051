#!/bin/bash
# source-level profile of the persistent GMRES kernel (report small enough to travel: one kernel, one launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gmres_fused' -s 1 -c 1 \
    -f -o gpurun_out/r2o_fused python bench.py --steps 1 --warmup 1 --ksp-maxit 90 --no-cpu-baseline --no-parity --spmv-launches 2 > gpurun_out/r2o_fused_bench.log 2>&1
echo "fused capture rc=$?"
ls -la gpurun_out/r2o_fused.ncu-rep
