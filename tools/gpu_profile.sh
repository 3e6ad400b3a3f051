#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for the Newton-step bench (run under gpurun, 1 GPU).
#   tools/gpu_profile.sh <tag>      writes gpurun_out/<tag>_launches.csv, <tag>_full.ncu-rep, <tag>_bench.json
# The launch list and the full capture run the bench with the Krylov solve capped (--ksp-maxit) so that the
# serialised, ~40x-replayed kernels finish in minutes; the kernels' SHARES of a Krylov iteration do not
# depend on the cap.  Numbers printed by runs under ncu are never bench values.
set -u
TAG=${1:-prof}
MAXIT=${2:-64}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --ksp-maxit $MAXIT --no-cpu-baseline --spmv-launches 5 > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k 'regex:k_bsr_spmv|k_ilu0_block_solve|k_mdot_all|k_maxpy_all|k_jacobian|k_residual|k_eos|k_ilu0_factor' -s 40 -c 14 \
    -f -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --ksp-maxit 40 --no-cpu-baseline --spmv-launches 5 > gpurun_out/${TAG}_full_bench.log 2>&1
echo "full capture rc=$?"
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
