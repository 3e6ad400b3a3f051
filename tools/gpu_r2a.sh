#!/bin/bash
# round-2 GPU job A (1 GPU): full GPU test suite, bench lines of configs 2 / 4 / 5(1 GPU), Krylov microbench with the
# global ILU(0) and a restart sweep, ncu of the assembly kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_env.txt; nproc >> gpurun_out/r2a_env.txt
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; tail -c 600 gpurun_out/r2a_bench_c2.err
timeout 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; tail -c 600 gpurun_out/r2a_bench_c4.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/r2a_bench_c5_1gpu.json 2> gpurun_out/r2a_bench_c5_1gpu.err; tail -c 600 gpurun_out/r2a_bench_c5_1gpu.err
timeout 900 python tools/microbench.py --global-ilu --restarts 30 60 100 200 > gpurun_out/r2a_micro.json 2> gpurun_out/r2a_micro.err; tail -c 600 gpurun_out/r2a_micro.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_jacobian|k_eos|k_residual|k_ilu0_factor|k_ilu_repack|k_gather_vals|k_transitions" -c 14 -o gpurun_out/r2a_k4 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --ksp-maxit 3 > gpurun_out/r2a_ncu_k4.log 2>&1; tail -3 gpurun_out/r2a_ncu_k4.log
head -c 1500 gpurun_out/r2a_bench_c2.json; echo; head -c 1200 gpurun_out/r2a_bench_c4.json; echo; cat gpurun_out/r2a_micro.json
