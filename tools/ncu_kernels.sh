#!/bin/bash
# ncu --set full capture of selected kernels (microbench driver: SpMV, PC apply, short GMRES solves)
#   tools/ncu_kernels.sh <tag> <kernel regex> [skip] [count]
TAG=${1:-k}
RE=${2:-'k_bsr_spmv|k_ilu0_block_solve|k_mdot_all|k_maxpy_all'}
SKIP=${3:-20}
COUNT=${4:-24}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $COUNT -f -o gpurun_out/${TAG}_kern \
    python tools/microbench.py --its 31 > gpurun_out/${TAG}_kern.log 2>&1
echo "ncu rc=$?"
